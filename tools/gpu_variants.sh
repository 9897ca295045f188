# time every development variant haskelldca_b200/_build/libdca_b200_var_*.so on the GPU box (tools/build_variants.py built them):
# a quick parity check against the oracle, then the throughput at 4096 (twice) and 8192 replicas
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v in haskelldca_b200/_build/libdca_b200_var_*.so; do
  [ -f "$v" ] || continue
  echo "== variant $v"
  DCA_LIB=$PWD/$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "philox_replicas or randomised" 2>&1 | tail -1
  for n in 4096 4096 8192; do DCA_LIB=$PWD/$v timeout 120 python tools/prof_run.py $n 10000 20000; done
done
