set -x
mkdir -p gpurun_out
MPC_B200_LIB=ab/lib_split6.so python tools/ab_split.py 4096 10 > gpurun_out/r2_ab_split.txt 2>&1
MPC_B200_LIB=ab/lib_split6.so python tools/ab_split.py 2048 16 >> gpurun_out/r2_ab_split.txt 2>&1
MPC_B200_LIB=ab/lib_split6.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mpc_ -s 30 -c 16 --csv --log-file gpurun_out/r2_launches_split.csv python tools/ab_split.py 4096 10 > /dev/null 2>&1
