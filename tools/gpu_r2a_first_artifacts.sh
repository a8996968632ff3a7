set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_qp.json 2> gpurun_out/r2_bench_qp.err
python bench.py --workload tick_mixed_gaits_h10_n2048 --steps 20 --warmup 5 > gpurun_out/r2_bench_tick.json 2> gpurun_out/r2_bench_tick.err
python bench.py --workload tick_mixed_gaits_h10_n2048 --tick-cold --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_tick_cold.json 2>> gpurun_out/r2_bench_tick.err
python bench.py --workload mixed_robots_h16_n2048 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_h16.json 2> gpurun_out/r2_bench_h16.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_tick.csv python bench.py --workload tick_mixed_gaits_h10_n2048 --steps 3 --warmup 3 --no-cpu-baseline --no-side > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mpc_tick_kernel -s 7 -c 1 -o gpurun_out/r2_tick_full python bench.py --workload tick_mixed_gaits_h10_n2048 --steps 3 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/r2_ncu_tick.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mpc_solve_kernel -s 6 -c 1 -o gpurun_out/r2_h16_full python bench.py --workload mixed_robots_h16_n2048 --steps 3 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/r2_ncu_h16.log 2>&1
ls -la gpurun_out | tail -20
