set -x
mkdir -p gpurun_out
MPC_B200_LIB=ab/lib_split6.so python tools/occ_scan.py > gpurun_out/r2_occ_scan.txt 2>&1
