cd $GRAFT_REPO_ROOT
for n in 4096 5920; do timeout 120 python tools/prof_run.py $n 10000 20000; done
