# one kernel experiment on the GPU box: quick parity, throughput at 4096 / 5920 / 8192 replicas, phase clocks;
# then the same throughput runs for every development variant haskelldca_b200/_build/libdca_b200_var_*.so
cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_path.py -x -q -k "config1 or philox_replicas or extreme or randomised or stop_rules or bench_configuration" 2>&1 | tail -4
for n in 4096 5920 8192; do timeout 120 python tools/prof_run.py $n 10000 20000; done
if [ -f haskelldca_b200/_build/libdca_b200_phase.so ]; then timeout 120 python tools/phase_clock.py 740 5000 20000; fi
for v in haskelldca_b200/_build/libdca_b200_var_*.so; do
  [ -f "$v" ] || continue
  echo "== variant $v"
  DCA_LIB=$PWD/$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "philox_replicas or extreme" 2>&1 | tail -2
  for n in 4096 5920 8192; do DCA_LIB=$PWD/$v timeout 120 python tools/prof_run.py $n 10000 20000; done
done
