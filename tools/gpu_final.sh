mkdir -p gpurun_out
python bench.py > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e_pytest_gpu.log
SWEEP_SIZES=256,1024,4096,16384,65536 python tools/sweep.py > gpurun_out/r2e_sweep_1gpu.json 2> gpurun_out/r2e_sweep_1gpu.err
