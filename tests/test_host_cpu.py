"""CPU-side checks of the product's host layer: C-ABI exports, Opt parser, report strings."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from haskelldca_b200 import _ffi

    L = _ffi.load()
    hdr = open(os.path.join(ROOT, "include", "dca_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dca_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for name in declared:
        assert getattr(L, name) is not None


def test_no_device_fails_loudly_instead_of_falling_back():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from haskelldca_b200 import Opt, Replicas, _ffi, gridfuncs

    with pytest.raises(_ffi.DcaError, match="no CUDA device"):
        Replicas(Opt())
    with pytest.raises(_ffi.DcaError, match="no CUDA device"):
        gridfuncs.feature_rep(gridfuncs.mk_grid())


def test_product_does_not_import_the_oracle():
    for top in ("haskelldca_b200", "include", "host", "hs"):  # everything that ships
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".hs")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "pyoracle" not in src and "dca_oracle" not in src and "oracle/" not in src, f


def test_neighbourhood_tables_host_side():
    import pyoracle as po
    from haskelldca_b200 import gridneighs

    assert gridneighs.get_neighs(1, (3, 3), False) == [(3, 2), (4, 2), (4, 3), (3, 4), (2, 4), (2, 3)]
    for cell in [(0, 0), (3, 3), (6, 2), (4, 5)]:
        for d in (1, 2, 3, 4):
            assert gridneighs.get_neighs(d, cell, True) == po.get_neighs(d, *cell, True)


def test_opt_defaults_and_flags():  # src/Opt.hs:28-94
    from haskelldca_b200 import get_opts

    o = get_opts([], rand=1234)
    assert (o.call_dur, o.call_dur_hoff, o.call_rate, o.hoff_prob) == (3.0, 1.0, 200.0, 0.0)
    assert (o.n_events, o.log_iter) == (10000, 1000)
    assert (o.learning_rate, o.learning_rate_avg, o.learning_rate_grad, o.min_loss) == (2.52e-6, 0.06, 5e-6, 0.0)
    assert o.rng_seed == 1234 and not o.verify_reuse_constraint and not o.verify_nelig_bounds
    o = get_opts("--fixed_rng --hoff_prob 0.15 --n_events 100000 --call_rate 250 --verify_reuse_constraint --backend gpu".split(), rand=99)
    assert o.rng_seed == 0 and o.hoff_prob == 0.15 and o.n_events == 100000 and o.call_rate == 250 and o.verify_reuse_constraint
    with pytest.raises(SystemExit):
        get_opts(["--backend", "interp"])


def test_stats_reports_format():  # src/Stats.hs:83-152
    from haskelldca_b200 import _ffi, stats

    st = np.array([5520, 4913, 607, 0, 4480, 0, 540, 61], dtype=np.int64)
    new, hoff, tot = stats.stats_cumus(st)
    assert new == 607 / 5521 and hoff == 0 and tot == 607 / 5521
    out = np.zeros(3)
    _ffi.load().dca_stats_cumus(_ffi.ptr(st), _ffi.ptr(out))
    assert out.tolist() == [new, hoff, tot]
    s = stats.stats_report_period(st, 10000, 1000, np.float32(-1234.5), 1000, 333.85226)
    assert s == "Blocking probability events 9000-10000: 0.1128, cumulative 0.1099, avg. loss: -1.2, avg. reward 333.8523"
    e = stats.stats_report_end(st, 433, 10000)
    assert e.startswith("Finished 10000 events. Blocking probability 0.1099 for new calls.\n5520 new arrivals of which 607 have been rejected. 433 calls in progress")
    assert "SOME CALLS WERE LOST" not in e
    assert "SOME CALLS WERE LOST" in stats.stats_report_end(st, 432, 10000)


def test_haskell_bindings_match_the_library():
    """hs/DcaB200/*.hs cannot be compiled here (no GHC): pin what can be checked without it --
    every `foreign import`ed symbol is exported, and the dca_config offsets poked by Fused.hs
    are the C struct's (ctypes mirrors include/dca_b200.h)."""
    import ctypes as C
    import re

    from haskelldca_b200 import _ffi

    L = C.CDLL(_ffi.lib_path())
    ffi_src = open(os.path.join(ROOT, "hs", "DcaB200", "FFI.hs")).read()
    syms = set(re.findall(r'foreign import ccall safe "&?(\w+)"', ffi_src))
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), s
        assert s in _ffi.SYMBOLS, s  # and it is declared in the header mirror
    fused = open(os.path.join(ROOT, "hs", "DcaB200", "Fused.hs")).read()
    offs = {k: int(v) for k, v in re.findall(r"\b(off\w+|sizeofConfig) = (\d+)", fused)}
    want = {"offNReplicas": "n_replicas", "offSeed": "seed", "offRngMode": "rng_mode", "offOrder": "order",
            "offNEvents": "n_events", "offMinLoss": "min_loss", "offVerifyRC": "verify_reuse_constraint",
            "offVerifyNB": "verify_nelig_bounds", "offCallDur": "call_dur", "offCallDurHoff": "call_dur_hoff",
            "offCallRate": "call_rate", "offHoffProb": "hoff_prob", "offAlphaNet": "alpha_net",
            "offAlphaAvg": "alpha_avg", "offAlphaGrad": "alpha_grad"}
    for hs_name, field in want.items():
        assert offs[hs_name] == getattr(_ffi.Config, field).offset, hs_name
    assert offs["sizeofConfig"] == C.sizeof(_ffi.Config)


def test_haskell_patch_applies(tmp_path):
    """hs/patches/0001-backend-gpu.patch is a real unified diff against the reference modules: it must apply cleanly
    (`patch -p1`) to a copy of the files it names.  Needs the reference tree (absent on the GPU box: skipped there)."""
    import shutil
    import subprocess

    ref = "/root/reference"
    patch = os.path.join(ROOT, "hs", "patches", "0001-backend-gpu.patch")
    files = ["src/Base.hs", "src/Opt.hs", "src/AccUtils.hs", "src/SimRunner.hs", "dca.cabal"]
    if not all(os.path.exists(os.path.join(ref, f)) for f in files) or not shutil.which("patch"):
        pytest.skip("reference tree or patch(1) not available")
    for f in files:
        dst = tmp_path / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(ref, f), dst)
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "| GPU" in (tmp_path / "src/Base.hs").read_text()
    sim = (tmp_path / "src/SimRunner.hs").read_text()
    assert "gpuGetAction RefOrder" in sim and "runSimGpu" in sim and "runN bkend getAction" in sim
    opt = (tmp_path / "src/Opt.hs").read_text()
    assert '"gpu" -> return GPU' in opt and "n_replicas" in opt
    # the modules the patch imports exist and export what it uses
    backend = open(os.path.join(ROOT, "hs", "DcaB200", "Backend.hs")).read()
    fused = open(os.path.join(ROOT, "hs", "DcaB200", "Fused.hs")).read()
    assert "gpuGetAction" in backend and "gpuGridStepTrain" in backend and "GpuOrder(..)" in backend
    assert "module DcaB200.Fused (runSimGpu, RngSource(..)" in fused


def test_dca_create_rejects_a_run_longer_than_the_32_bit_event_ids_allow():
    """Argument checks come before the device is touched, so this runs without a GPU: n_events > DCA_MAX_EVENTS is
    DCA_ERR_ARG (-1), not DCA_ERR_NO_DEVICE."""
    import ctypes as C

    from haskelldca_b200 import _ffi

    L = _ffi.load()
    cfg = _ffi.Config()
    assert L.dca_config_defaults(C.byref(cfg)) == 0
    cfg.n_replicas = 1
    cfg.n_events = 1_400_000_001
    h = C.c_void_p()
    assert L.dca_create(C.byref(cfg), C.byref(h)) == -1 and not h.value
    assert b"n_events" in L.dca_last_error(None)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): exactly one line on stdout, a JSON object with the
    contract's keys; whatever else a library writes to file descriptor 1 goes to stderr (bench.own_stdout)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-events", "200"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "events/s" and rec["higher_is_better"] is True
    assert rec["value"] > 0 and rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0 and rec["gpu_launches"] == 0
    # the redirection itself: a write to fd 1 after own_stdout() must not reach stdout
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.own_stdout(); os.write(1, b'library banner\\n'); "
            "bench.emit({'ok': 1})" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == '{"ok": 1}' and "library banner" in out.stderr
