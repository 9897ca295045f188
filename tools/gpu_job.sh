cd $GRAFT_REPO_ROOT
bash tools/gpu_ncu_full.sh r2_v11 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_v11_bench.json 2> gpurun_out/r2_v11_bench.err || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_v11_launches_default_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check > gpurun_out/r2_v11_ncu_bench.log 2>&1
wc -l gpurun_out/r2_v11_launches_default_bench.csv
timeout 120 python tools/prof_run.py 5920 10000 20000
