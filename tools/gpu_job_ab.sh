mkdir -p gpurun_out
python tools/ab_time.py $AB_LIBS > gpurun_out/$AB_OUT 2>&1
