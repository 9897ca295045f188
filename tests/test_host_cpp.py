"""The C++ host layer (host/dca_host.hpp) and its CLI `dca-exe` (host/main.cpp): the reference's
Opt / Stats / SimRunner API over the C ABI, in the compiled language of this image."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe():
    from haskelldca_b200 import build as b

    b.build()
    return b.build_host()


def _run(*args):
    return subprocess.run([_exe(), *args], capture_output=True, text=True, timeout=300)


def test_cli_flags_are_the_reference_flags():  # src/Opt.hs:28-82
    r = _run("--help")
    assert r.returncode == 0
    for flag in ("--call_dur", "--call_dur_hoff", "--call_rate", "--hoff_prob", "--n_events", "--log_iter", "--learning_rate",
                 "--learning_rate_avg", "--learning_rate_grad", "--verify_reuse_constraint", "--verify_nelig_bounds",
                 "--backend", "--min_loss", "--fixed_rng"):
        assert flag in r.stdout, flag
    assert "Call duration for new calls." in r.stdout and "Hand-off probability. Set to 0 to disable hand-offs." in r.stdout


def test_cli_usage_errors():
    r = _run("--bogus")
    assert r.returncode == 1 and "Invalid option `--bogus'" in r.stderr
    r = _run("--backend", "cpu")  # bkendOpt, src/Opt.hs:90-94: only the device backend exists here
    assert r.returncode == 1 and "Accepted backend is 'gpu'" in r.stderr
    r = _run("--n_events")
    assert r.returncode == 1 and "expects an argument" in r.stderr
    r = _run("--call_rate", "fast")
    assert r.returncode == 1 and "cannot parse" in r.stderr


def test_cli_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run("--fixed_rng", "--n_events", "100")
    assert r.returncode == 1 and "no CUDA device (there is no CPU fallback)" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_cli_run_matches_the_python_host_and_the_oracle():
    """`dca-exe --fixed_rng --rng mt19937 --order ref`: the same report lines as sim_runner.run_sim, which
    tests/test_gpu_sim_runner.py ties to the oracle."""
    from haskelldca_b200 import Opt, sim_runner

    r = _run("--fixed_rng", "--rng", "mt19937", "--order", "ref", "--n_events", "3000", "--log_iter", "1000")
    assert r.returncode == 0, r.stderr
    got = r.stdout.rstrip("\n").split("\n")
    want = []
    sim_runner.run_sim(Opt(n_events=3000, log_iter=1000, order="ref", rng="mt19937", rng_seed=0), out=want.append)
    want = "\n".join(want).split("\n")
    assert got[:-1] == want[:-1]  # everything but the wall-clock speed line
    assert got[-1].startswith("Simulation duration: ") and "events/second" in got[-1]


@pytest.mark.gpu
def test_cli_batched_run():
    r = _run("--n_replicas", "64", "--n_events", "2000", "--log_iter", "2000", "--hoff_prob", "0.15", "--verify_reuse_constraint")
    assert r.returncode == 0, r.stderr
    assert "Success" in r.stdout and "speed 128000 events" in r.stdout
