# v13 validation on one B200: the whole GPU suite, the bench line, the reference arm, the ncu launch list of the bench
# command and one `ncu --set full` capture of the run kernel (each ncu pass after the same command exited 0 without it)
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_v13_bench.json 2> gpurun_out/r2_v13_bench.err; tail -2 gpurun_out/r2_v13_bench.err; cat gpurun_out/r2_v13_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_v13_bench_reference_arm.json 2> gpurun_out/r2_v13_ref.err; cat gpurun_out/r2_v13_bench_reference_arm.json
for n in 740 2220 4096 5920 8192; do timeout 120 python tools/prof_run.py $n 10000 20000; done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_v13_launches_default_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-spot-check > gpurun_out/r2_v13_ncu_launches.log 2>&1
bash tools/gpu_ncu_full.sh r2_v13
