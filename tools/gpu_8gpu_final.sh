# the driver's launch line at N = 8 on the final build (bench.py prints one JSON line on stdout)
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_8gpu.json 2> gpurun_out/r2_final_bench_8gpu.err
echo rc=$? lines=$(wc -l < gpurun_out/r2_final_bench_8gpu.json); cut -c1-300 gpurun_out/r2_final_bench_8gpu.json
