"""BASELINE configs[4]: a call_rate x learning-rate grid of replicas, sharded over the GPUs of one box.

  python tools/sweep_grid.py [--seeds 64] [--events 20000]                      (one GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_grid.py ...

Every sweep point (call_rate, (alpha_net, alpha_avg, alpha_grad)) gets `seeds` replicas with their own
Philox streams; rank r runs a contiguous block of the replica list (replica_shards.shard) with per-replica
parameter arrays; the per-point (rejected, arrived) sums are all-reduced over NCCL and rank 0 prints the
mean blocking probability per point."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=64)
    ap.add_argument("--events", type=int, default=20000)
    ap.add_argument("--hoff_prob", type=float, default=0.0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    from haskelldca_b200 import Opt, Replicas
    from haskelldca_b200.replica_shards import shard, slice_params, sweep_grid

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rates = [150.0, 175.0, 200.0, 225.0, 250.0, 275.0, 300.0]
    lrs = [(2.52e-6 * k, 0.06, 5e-6) for k in (0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0)]
    per, point, points = sweep_grid(rates, lrs, a.seeds)
    first, count = shard(len(point), world, rank)
    with Replicas(Opt(n_replicas=count, n_events=a.events, order="fast", rng="philox", rng_seed=1, device=local,
                      first_replica=first, hoff_prob=a.hoff_prob), per_replica=slice_params(per, first, count)) as sims:
        ms = sims.run(a.events)
        st = sims.read_summary()["stats"].astype(np.float64)
    pt = point[first:first + count]
    sums = np.stack([np.bincount(pt, weights=st[:, 2], minlength=len(points)), np.bincount(pt, weights=st[:, 0], minlength=len(points))])
    t = torch.from_numpy(sums).cuda()
    if world > 1:
        dist.all_reduce(t)  # the only collective: per-point (rejected, arrived) sums
    sums = t.cpu().numpy()
    if rank == 0:
        bp = sums[0] / (sums[1] + 1.0)
        print(f"# {len(point)} replicas ({len(points)} points x {a.seeds} seeds) x {a.events} events on {world} GPU(s); "
              f"rank 0: {count} replicas in {ms:.1f} ms")
        print("call_rate/h  " + "  ".join(f"lr={lr[0]:.2e}" for lr in lrs))
        for i, r in enumerate(rates):
            print(f"{r:10.0f}   " + "  ".join(f"{bp[i * len(lrs) + j]:11.4f}" for j in range(len(lrs))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
