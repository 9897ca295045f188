cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lookahead or config2 or philox_ref or guard" 2>&1 | tail -15
