# REF-order run kernel after a change: the GPU suite, then its steady-state throughput (one wave, the bench's batch) and phase clocks
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for n in 296 4096; do timeout 300 python tools/ref_run.py $n 1000 3000 ref; done
timeout 300 python tools/ref_run.py 4096 300 0 ref
timeout 300 python tools/ref_phase_clock.py 296 1000 3000
