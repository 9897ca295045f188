# REF-order run kernel: development variants of the fold (haskelldca_b200/_build/libdca_b200_var_*.so), parity subset + timing
cd $GRAFT_REPO_ROOT
for v in haskelldca_b200/_build/libdca_b200.so haskelldca_b200/_build/libdca_b200_var_*.so; do
  [ -f "$v" ] || continue
  echo "== $v"
  DCA_LIB=$PWD/$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "philox_ref_order or config1" 2>&1 | tail -1
  for n in 296 4096; do DCA_LIB=$PWD/$v timeout 120 python tools/ref_run.py $n 1000 3000 ref; done
done
