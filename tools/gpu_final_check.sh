# final check of the round on one B200: the whole GPU suite, smoke(), the bench line and the reference arm
cd $GRAFT_REPO_ROOT
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cat gpurun_out/r2_final_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_bench_reference_arm.json 2> gpurun_out/r2_final_ref.err; cat gpurun_out/r2_final_bench_reference_arm.json
for n in 4096 8192; do timeout 120 python tools/prof_run.py $n 10000 20000; done
timeout 300 python tools/ref_run.py 4096 1000 3000 ref
# one ncu capture of the REF-order run kernel in steady state (the same command ran without ncu just above)
timeout 300 python tools/ref_run.py 4096 300 3000 ref && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run -s 1 -c 1 -f -o gpurun_out/r2_ref_final python tools/ref_run.py 4096 300 3000 ref > gpurun_out/r2_ref_final_ncu.log 2>&1
tail -2 gpurun_out/r2_ref_final_ncu.log
