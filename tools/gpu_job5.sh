mkdir -p gpurun_out
for lib in ab/lib_timing_base.so ab/lib_timing_dense.so; do for ex in 153600 14336 0; do echo "== $lib extra_smem=$ex"; MPC_B200_LIB=$lib MPC_B200_DEBUG_EXTRA_SMEM=$ex python tools/stage_timing.py; done; done > gpurun_out/r2_stage_timing.txt 2>&1
