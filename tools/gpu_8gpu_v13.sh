# v13 on one 8-GPU box: BASELINE config 4 (65 536 replicas, 8192 per GPU) and the default bench (4096 per GPU)
cd $GRAFT_REPO_ROOT
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --replicas 8192 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_v13_config4_bench_${N}gpu.json 2> gpurun_out/r2_v13_config4_bench_${N}gpu.err
tail -2 gpurun_out/r2_v13_config4_bench_${N}gpu.err; cut -c1-400 gpurun_out/r2_v13_config4_bench_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-inlib-allreduce > gpurun_out/r2_v13_bench_${N}gpu.json 2> gpurun_out/r2_v13_bench_${N}gpu.err
tail -2 gpurun_out/r2_v13_bench_${N}gpu.err; cut -c1-400 gpurun_out/r2_v13_bench_${N}gpu.json
