"""Stats arithmetic and report strings, mirroring src/Stats.hs:83-152."""
from __future__ import annotations

import numpy as np

ARR_NEW, ACC_NEW, REJ_NEW, ARR_HOFF, ENDED, REJ_HOFF, CURR_ARR_NEW, CURR_REJ_NEW = range(8)


def stats_cumus(st):  # Stats.hs:83-92
    rej_new, arr_new = float(st[REJ_NEW]), float(st[ARR_NEW])
    rej_hoff, arr_hoff = float(st[REJ_HOFF]), float(st[ARR_HOFF])
    return (rej_new / (arr_new + 1.0), rej_hoff / (arr_hoff + 1.0),
            (rej_new + rej_hoff) / (arr_new + arr_hoff + 1.0))


def stats_report_period(st, i: int, log_iter: int, loss_sum: float, loss_n: int, avg_reward: float) -> str:
    """statsReportPeriod, Stats.hs:94-124 (the caller resets the period counters)."""
    cumu_new, _, _ = stats_cumus(st)
    n_rej, n_arr = int(st[CURR_REJ_NEW]), int(st[CURR_ARR_NEW])
    avg_loss = float(np.float32(loss_sum) / np.float32(loss_n)) if loss_n else float("nan")
    bp = n_rej / (n_arr + 1)
    return ("Blocking probability events %d-%d: %.4f, cumulative %.4f, avg. loss: %.1f, avg. reward %.4f"
            % (max(i - log_iter, 0), i, bp, cumu_new, avg_loss, avg_reward))


def stats_report_end(st, n_calls_in_progress: int, i: int) -> str:
    """statsReportEnd, Stats.hs:126-152."""
    cumu_new, cumu_hoff, cumu_tot = stats_cumus(st)
    s = "Finished %d events. Blocking probability %.4f for new calls" % (i, cumu_new)
    s += (", %.4f for hand-offs, %.4f total." % (cumu_hoff, cumu_tot)) if st[REJ_HOFF] > 0 else "."
    reported = int(st[ARR_NEW] + st[ARR_HOFF] - st[REJ_NEW] - st[REJ_HOFF] - st[ENDED])
    s += "\n%d new arrivals of which %d have been rejected. %d calls in progress at simulation end." % (
        st[ARR_NEW], st[REJ_NEW], n_calls_in_progress)
    if reported != n_calls_in_progress:
        s += ("\n\n\nSOME CALLS WERE LOST. This should never happen on runs that exits with 'Success'. According to"
              " reported calls there should be %d calls currently in progress at simulation end but there are %d."
              % (reported, n_calls_in_progress))
    return s
