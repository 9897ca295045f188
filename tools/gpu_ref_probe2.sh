# timing probes of the REF-order run kernel (wrong results on purpose): see the DCA_REF_PROBE_* switches
cd $GRAFT_REPO_ROOT
for v in haskelldca_b200/_build/libdca_b200_var_*.so; do
  echo "== $v"
  for n in 296 4096; do DCA_LIB=$PWD/$v timeout 120 python tools/ref_probe_run.py $n 1000 3000; done
done
