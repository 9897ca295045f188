cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_path.py tests/test_gpu_sim_runner.py -x -q 2>&1 | tail -4
for n in 4096 5920; do timeout 120 python tools/prof_run.py $n 10000 20000; done
