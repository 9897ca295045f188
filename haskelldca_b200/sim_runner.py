"""runSim / runPeriod (src/SimRunner.hs:132-243) over the fused device kernel.

One call to `Replicas.run(log_iter)` replaces log_iter iterations of runStep; the
period and end reports are formatted exactly like src/Stats.hs.
"""
from __future__ import annotations

import time

from . import _ffi, stats
from .opt import Opt
from .simulator import Replicas

SIM_STOP = {  # SimStop, SimRunner.hs:45-50
    _ffi.RUNNING: "Paused", _ffi.SUCCESS: "Success", _ffi.ZERO_LOSS: "ZeroLoss", _ffi.NAN_LOSS: "NaNLoss",
    _ffi.REUSE_VIOLATED: "InternalError ReuseConstraintViolated", _ffi.NELIG_OVERFLOW: "InternalError NEligOverflow",
    _ffi.TAPE_EXHAUSTED: "InternalError TapeExhausted", _ffi.UNINITIALIZED: "InternalError Uninitialized",
    _ffi.QUEUE_OVERFLOW: "InternalError EventQueueOverflow",
}


def run_sim(opt: Opt, out=print, replica: int = 0):
    """Run the simulation(s) to completion, printing replica `replica`'s reports like the reference CLI."""
    t0 = time.time()
    sims = Replicas(opt)
    try:
        device_ms = 0.0
        status, last_it = _ffi.RUNNING, -1
        while status == _ffi.RUNNING:
            device_ms += sims.run(opt.log_iter)  # one period (SimRunner.hs:167-168)
            status, it, _ = sims.read_status(replica)
            if it == last_it:  # a period without progress can only be a library fault: never spin on it
                raise _ffi.DcaError(f"replica {replica} made no progress in a period (status {status}, iter {it})")
            last_it = it
            st = sims.read_stats(replica)
            loss_sum, loss_n = sims.read_period(replica, reset=True)
            avg = sims.read_state(replica)["avg_reward"] if opt.n_replicas == 1 else sims.read_summary()["avg_reward"][replica]
            out(stats.stats_report_period(st, it, opt.log_iter, loss_sum, loss_n, avg))
            if status != _ffi.RUNNING:
                out(SIM_STOP[status])
        state = sims.read_state(replica)
        out(stats.stats_report_end(state["stats"], int(state["grid"].sum()), state["iter"]))
        dt = time.time() - t0
        total = int(sims.read_summary()["iter"].sum())
        t_sim = state["event"][0]
        out("\nSimulation duration: %dm%ds in sim time, %dm%ds wall clock with speed %d events at %.0f events/second"
            % (int(t_sim), int((t_sim - int(t_sim)) * 60.0), int(dt // 60), int(dt) - 60 * int(dt // 60), total, total / dt))
        return dict(status=status, state=state, summary=sims.read_summary(), device_ms=device_ms, wall_s=dt)
    finally:
        sims.close()


if __name__ == "__main__":  # CLI analogue of `dca-exe` (src/Main.hs:10-22)
    from .opt import get_opts

    run_sim(get_opts())
