# one `ncu --set full` capture of the run kernel on a warmed-up state (after the same command exited 0 without ncu)
cd $GRAFT_REPO_ROOT
NAME=${1:-r2_v10}
timeout 300 python tools/prof_run.py 4096 2000 20000 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_fast -s 1 -c 1 -f -o gpurun_out/$NAME python tools/prof_run.py 4096 2000 20000 > gpurun_out/${NAME}_ncu.log 2>&1
tail -3 gpurun_out/${NAME}_ncu.log
ls -la gpurun_out/$NAME.ncu-rep
