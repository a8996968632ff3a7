mkdir -p gpurun_out
for lib in ab/lib_timing_pf.so; do for ex in 153600 0; do echo "== $lib extra_smem=$ex"; MPC_B200_LIB=$lib MPC_B200_DEBUG_EXTRA_SMEM=$ex python tools/stage_timing.py; done; done > gpurun_out/r2_stage_timing_pf.txt 2>&1
python tools/ab_time.py ab/lib_base.so ab/lib_pf.so > gpurun_out/r2_ab_pf.txt 2>&1
