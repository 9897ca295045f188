cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_order_stats.py -x -q -s 2>&1 | tail -25
timeout 600 python tools/order_stats.py > gpurun_out/r2_fast_vs_ref_order_stats.md 2> gpurun_out/r2_order_stats.err; tail -3 gpurun_out/r2_order_stats.err
