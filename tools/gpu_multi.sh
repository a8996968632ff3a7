set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2d_gpus.txt
python -m pytest tests/test_dist.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2d_pytest_dist_gpu.log
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --workload tick_mixed_gaits_h10_n2048 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_config3_tick_8gpu.json 2> gpurun_out/r2d_config3.err
$TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --workload mixed_robots_h16_n2048 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_config4_h16_4gpu.json 2> gpurun_out/r2d_config4.err
for n in 8 4 2; do SWEEP_SIZES=1024,4096,16384,65536 $TR --nproc-per-node $n --master-port $((29520+n)) tools/sweep.py > gpurun_out/r2d_sweep_${n}gpu.json 2> gpurun_out/r2d_sweep_${n}gpu.err; done
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-side > gpurun_out/r2d_config2_qp_8gpu.json 2> gpurun_out/r2d_config2_8.err
ls -la gpurun_out | tail -15
