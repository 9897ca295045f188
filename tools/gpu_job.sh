cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_ncu_full.sh r2_v12 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_v12_bench.json 2> gpurun_out/r2_v12_bench.err || exit 1
cut -c1-330 gpurun_out/r2_v12_bench.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_v12_launches_default_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check > gpurun_out/r2_v12_ncu_bench.log 2>&1
wc -l gpurun_out/r2_v12_launches_default_bench.csv
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_v12_bench_reference_arm.json 2>gpurun_out/r2_ref.err; cut -c1-200 gpurun_out/r2_v12_bench_reference_arm.json
for n in 740 2220 4096 5920 8192; do timeout 120 python tools/prof_run.py $n 10000 20000; done
