# two ranks of bench.py under torchrun (the driver's launch line) after a change of the multi-rank flow
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err
echo rc=$?; tail -3 gpurun_out/r2_final_bench_2gpu.err; cat gpurun_out/r2_final_bench_2gpu.json | cut -c1-400
