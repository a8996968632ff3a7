mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_run.py > gpurun_out/r2f_racecheck.txt 2>&1
timeout 70 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2f_memcheck.txt 2>&1
