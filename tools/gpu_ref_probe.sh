# the REF-order run kernel in steady state: one resident wave, two, and the bench's replica count; one ncu capture
cd $GRAFT_REPO_ROOT
nvidia-smi -L
for n in 296 592 4096; do timeout 300 python tools/ref_run.py $n 1000 3000 ref; done
timeout 300 python tools/ref_run.py 4096 300 0 ref
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run -s 1 -c 1 -f -o gpurun_out/r2_ref_probe python tools/ref_run.py 296 300 3000 ref > gpurun_out/r2_ref_probe_ncu.log 2>&1
tail -3 gpurun_out/r2_ref_probe_ncu.log
