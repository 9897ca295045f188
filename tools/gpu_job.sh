cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/step_cost.py > gpurun_out/r2_step_cost.json 2> gpurun_out/r2_step_cost.err; tail -3 gpurun_out/r2_step_cost.err; cat gpurun_out/r2_step_cost.json
