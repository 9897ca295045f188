"""The N > 1 path on CPU: two gloo ranks shard the replicas, run them (here through the oracle -- there is
no GPU in this container), and reduce the Stats counters; the result must equal the single-process run.
Covers haskelldca_b200/replica_shards.py, which bench.py and the multi-GPU driver use unchanged."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_blocks_partition_the_replicas():
    from haskelldca_b200.replica_shards import shard

    for total in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    assert shard(65536, 8, 3) == (3 * 8192, 8192)  # BASELINE configs[3]
    with pytest.raises(ValueError):
        shard(8, 2, 2)


def test_sweep_grid_is_config5():
    from haskelldca_b200.replica_shards import per_point_blocking, slice_params, sweep_grid

    rates = [150, 175, 200, 225, 250, 275, 300]
    lrs = [(2.52e-6 * k, 0.06, 5e-6) for k in (0.25, 0.5, 1, 2, 4, 8, 16, 32)]
    per, point, points = sweep_grid(rates, lrs, seeds_per_point=4)
    assert len(points) == 56 and len(point) == 224 and per["call_rate"].shape == (224,)
    assert per["call_rate"][0] == 150 and per["call_rate"][-1] == 300
    assert per["learning_rate"].dtype == np.float32 and np.isclose(per["learning_rate"][4 * 7], 2.52e-6 * 32)
    sl = slice_params(per, 112, 112)
    assert sl["call_rate"][0] == per["call_rate"][112] and len(sl["learning_rate_grad"]) == 112
    st = np.zeros((224, 8), np.int64)
    st[:, 0], st[:, 2] = 99, point
    bp = per_point_blocking(st, point, 56)
    assert np.allclose(bp, np.arange(56) * 4 / (99 * 4 + 1.0))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, total, n_events, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import pyoracle as po
    from haskelldca_b200.replica_shards import all_reduce_stats, blocking_probabilities, shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard(total, world, rank)
        o = po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=5, n_events=n_events, hoff_prob=0.1)
        done, st8, _, _ = po.run_replicas(o, first, count, n_events, 1)
        local = np.zeros(10, np.int64)
        local[:8], local[8] = st8, done
        summed = all_reduce_stats(local)
        out.put((rank, first, count, summed.tolist(), blocking_probabilities(summed)))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process():
    import torch.multiprocessing as mp

    import pyoracle as po
    from haskelldca_b200.replica_shards import blocking_probabilities

    total, n_events, world = 6, 600, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, total, n_events, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 3)]
    assert res[0][3] == res[1][3]  # every rank holds the same reduced counters
    o = po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=5, n_events=n_events, hoff_prob=0.1)
    done, st8, _, _ = po.run_replicas(o, 0, total, n_events, 1)
    assert res[0][3][:8] == list(st8) and res[0][3][8] == done == total * n_events
    assert np.allclose(res[0][4], blocking_probabilities(list(st8) + [done, 0]))
