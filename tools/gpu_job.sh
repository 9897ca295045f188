cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_v11_bench.json 2> gpurun_out/r2_v11_bench.err; tail -2 gpurun_out/r2_v11_bench.err; cut -c1-330 gpurun_out/r2_v11_bench.json
