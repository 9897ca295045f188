cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_bench_v10.json 2> gpurun_out/r2_bench_v10.err; tail -3 gpurun_out/r2_bench_v10.err; cat gpurun_out/r2_bench_v10.json
