"""Host-side logic of the multi-GPU path: which replicas a rank owns, the per-replica parameter
grid of BASELINE config 5, and the one collective of the path -- the sum of the Stats counters
(src/Stats.hs:15-40) that yields the blocking probabilities (statsCumus, src/Stats.hs:83-92).

Replicas are independent simulators, so ranks never exchange data inside the event loop; a rank is
one process bound to one GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import itertools

import numpy as np

N_SUM = 10  # dca_sum_stats: 8 Stats counters, events processed, replicas that failed


def shard(total_replicas: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [first, first + count) of rank `rank`; the first `total % world` ranks get one more."""
    if not (0 <= rank < world) or total_replicas < 0:
        raise ValueError("bad rank / world / total")
    base, extra = divmod(total_replicas, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def sweep_grid(call_rates, learning_rates, seeds_per_point: int):
    """BASELINE config 5: call_rate x (alpha_net, alpha_avg, alpha_grad) x seeds as per-replica arrays.

    Returns (per_replica dict for `Replicas(per_replica=...)`, point index of every replica, points).
    Replica i belongs to point i // seeds_per_point; its Philox stream is its global index."""
    points = list(itertools.product(call_rates, learning_rates))
    n = len(points) * seeds_per_point
    rate = np.empty(n, np.float64)
    an, aa, ag = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    point = np.repeat(np.arange(len(points)), seeds_per_point)
    for i, (cr, (a_net, a_avg, a_grad)) in enumerate(points):
        sl = slice(i * seeds_per_point, (i + 1) * seeds_per_point)
        rate[sl], an[sl], aa[sl], ag[sl] = cr, a_net, a_avg, a_grad
    per = dict(call_rate=rate, learning_rate=an, learning_rate_avg=aa, learning_rate_grad=ag)
    return per, point, points


def slice_params(per_replica: dict, first: int, count: int) -> dict:
    return {k: np.ascontiguousarray(v[first:first + count]) for k, v in per_replica.items()}


def all_reduce_stats(local, group=None):
    """Sum the int64[10] of dca_sum_stats over all ranks.  `local` is a numpy array (reduced through a
    CPU tensor: gloo) or a CUDA int64 tensor (NCCL, in place)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    if isinstance(local, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy()
    dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def blocking_probabilities(stats):
    """statsCumus (src/Stats.hs:83-92) on summed counters: (new, hand-off, total)."""
    st = np.asarray(stats, dtype=np.float64)
    return (st[2] / (st[0] + 1.0), st[5] / (st[3] + 1.0), (st[2] + st[5]) / (st[0] + st[3] + 1.0))


def per_point_blocking(stats_per_replica, point, n_points):
    """mean blocking probability per sweep point from [n_replicas][8] counters (config 5's report)."""
    st = np.asarray(stats_per_replica, dtype=np.float64)
    rej = np.bincount(point, weights=st[:, 2], minlength=n_points)
    arr = np.bincount(point, weights=st[:, 0], minlength=n_points)
    return rej / (arr + 1.0)
