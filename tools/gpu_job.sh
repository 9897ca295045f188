cd $GRAFT_REPO_ROOT
# launch list of the default bench command (after it exited 0 without ncu)
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_v10_bench.json 2> gpurun_out/r2_v10_bench.err || exit 1
tail -1 gpurun_out/r2_v10_bench.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_v10_launches_default_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check > gpurun_out/r2_v10_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_v10_ncu_bench.log | cut -c1-300
wc -l gpurun_out/r2_v10_launches_default_bench.csv
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_v10_bench_reference_arm.json 2>gpurun_out/r2_ref.err; cat gpurun_out/r2_v10_bench_reference_arm.json | cut -c1-400
