# the named multi-GPU configurations of BASELINE.json on one 8-GPU box
cd $GRAFT_REPO_ROOT
N=${1:-8}
nvidia-smi -L | wc -l
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# config 4: 65 536 replicas over 8 GPUs (8192 each) + NCCL reduction of the blocking-probability counters
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --replicas 8192 --steps 2 --warmup 3 > gpurun_out/r2_config4_bench_${N}gpu.json 2> gpurun_out/r2_config4_bench_${N}gpu.err
tail -2 gpurun_out/r2_config4_bench_${N}gpu.err; cat gpurun_out/r2_config4_bench_${N}gpu.json | cut -c1-600
# the default bench (4096 per GPU) at N GPUs
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-inlib-allreduce > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -2 gpurun_out/r2_bench_${N}gpu.err; cat gpurun_out/r2_bench_${N}gpu.json | cut -c1-300
# config 5: call_rate 150-300/h x 8 learning-rate settings as a replica grid
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/sweep_grid.py --seeds 1024 --events 100000 > gpurun_out/r2_config5_sweep_grid_${N}gpu.txt 2> gpurun_out/r2_config5_sweep_${N}gpu.err
tail -2 gpurun_out/r2_config5_sweep_${N}gpu.err; cat gpurun_out/r2_config5_sweep_grid_${N}gpu.txt
timeout 300 python -m pytest tests/test_gpu_sim_runner.py -x -q -k allreduce 2>&1 | tail -2
