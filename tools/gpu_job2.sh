set -x
mkdir -p gpurun_out
python tools/ab_time.py ab/lib_base.so ab/lib_dense4.so ab/lib_dense5.so > gpurun_out/r2_ab_dense.txt 2>&1
python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2_pytest_gpu2.log
