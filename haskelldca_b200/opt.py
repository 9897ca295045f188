"""Opt record and command-line flags, mirroring src/Opt.hs:10-94 flag for flag.

New flags only add to the reference's set: --backend gains `gpu`; --n_replicas,
--order and --rng select the batched device path.
"""
from __future__ import annotations

import argparse
import dataclasses
import random


@dataclasses.dataclass
class Opt:  # src/Opt.hs:10-25
    call_dur: float = 3.0            # callDurNew
    call_dur_hoff: float = 1.0       # callDurHoff
    call_rate: float = 200.0         # callRate
    hoff_prob: float = 0.0           # hoffProb
    n_events: int = 10000            # nEvents
    log_iter: int = 1000             # logIter
    learning_rate: float = 2.52e-6   # alphaNet
    learning_rate_avg: float = 0.06  # alphaAvg
    learning_rate_grad: float = 5e-6 # alphaGrad
    verify_reuse_constraint: bool = False
    verify_nelig_bounds: bool = False
    backend: str = "gpu"             # reference: interp | cpu ; here the device path
    min_loss: float = 0.0
    rng_seed: int = 0
    # additions of the batched build
    n_replicas: int = 1
    order: str = "fast"              # fast | ref   (fp32 reduction order)
    rng: str = "philox"              # philox | mt19937 (tape replay of the reference stream)
    device: int = 0
    first_replica: int = 0
    trace: int = 0                   # per-replica trace capacity
    warps_per_replica: int = 0
    hoff_lookahead: bool = False     # extension (the reference's TODO, README.md:71-72); dca_run in either order


def get_opts(argv=None, rand: int | None = None) -> Opt:
    """optparse-applicative parser of src/Opt.hs:28-82 restated with argparse."""
    d = Opt()
    p = argparse.ArgumentParser(description="DCA - Dynamic Channel Allocation by Reinforcement Learning")
    p.add_argument("--call_dur", type=float, default=d.call_dur, metavar="MINUTES", help="Call duration for new calls.")
    p.add_argument("--call_dur_hoff", type=float, default=d.call_dur_hoff, metavar="MINUTES", help="Call duration for handed-off calls.")
    p.add_argument("--call_rate", type=float, default=d.call_rate, metavar="PER_HOUR", help="Call arrival rate (new calls).")
    p.add_argument("--hoff_prob", type=float, default=d.hoff_prob, metavar="PROBABILITY", help="Hand-off probability. Set to 0 to disable hand-offs.")
    p.add_argument("--n_events", type=int, default=d.n_events, metavar="N", help="Simulation duration, in number of processed events.")
    p.add_argument("--log_iter", type=int, default=d.log_iter, metavar="N", help="How often to show run time statistics such as call blocking probability.")
    p.add_argument("--learning_rate", type=float, default=d.learning_rate, metavar="F", help="For neural net, i.e. state value update.")
    p.add_argument("--learning_rate_avg", type=float, default=d.learning_rate_avg, metavar="F", help="Learning rate for the average reward estimate.")
    p.add_argument("--learning_rate_grad", type=float, default=d.learning_rate_grad, metavar="F", help="Learning rate for gradient correction.")
    p.add_argument("--verify_reuse_constraint", action="store_true")
    p.add_argument("--verify_nelig_bounds", action="store_true")
    p.add_argument("--backend", type=str.lower, default=d.backend, choices=["gpu"],
                   help="'gpu' = libdca_b200 (sm_100a). The reference's 'interp'/'cpu' Accelerate backends are not part of this build.")
    p.add_argument("--min_loss", type=float, default=d.min_loss, metavar="F", help="Abort simulation if loss goes below given absolute value. Set to 0 to disable.")
    p.add_argument("--fixed_rng", action="store_true", help="Use a fixed (at 0) seed for the RNG. If this switch is not enabled, the seed is selected at random.")
    p.add_argument("--n_replicas", type=int, default=d.n_replicas)
    p.add_argument("--order", choices=["fast", "ref"], default=d.order)
    p.add_argument("--rng", choices=["philox", "mt19937"], default=d.rng)
    p.add_argument("--device", type=int, default=d.device)
    p.add_argument("--hoff_lookahead", action="store_true",
                   help="Extension (the reference's TODO): on the END half of a hand-off, choose the channel to free by looking ahead "
                        "at the hand-off arrival in the target cell (dca_run, either --order).")
    a = p.parse_args(argv)
    seed = 0 if a.fixed_rng else (rand if rand is not None else random.getrandbits(64))  # seedParser, Opt.hs:79-82
    kw = {k: v for k, v in vars(a).items() if k != "fixed_rng"}
    return Opt(rng_seed=seed, **kw)
