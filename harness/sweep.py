#!/usr/bin/env python
"""BASELINE config 5: dense 128^3 occupancy/loss sweep over 4096 random rotated SQ pairs (visu.py-style evaluation:
IoUAccuracy(128) and ExplicitLoss(128) forward, torch/visu.py:71-73), sharded by sample over the ranks.

    python harness/sweep.py [--pairs 4096]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 harness/sweep.py

No collective during compute; the per-pair results are gathered at the end.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sq_recovery_b200 as S                              # noqa: E402
from sq_recovery_b200 import distributed as D            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--render", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="", help="also append the JSON line to this file")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from sq_recovery_b200 import inputs as O          # seeded randsq / randquat workloads
    b0, b1 = D.shard_range(args.pairs, rank, world)
    true = O.random_params(args.pairs, 0)[b0:b1].to(dev)
    pred = O.random_params(args.pairs, 1)[b0:b1].to(dev)
    R = args.render
    iou = S.IoUAccuracy(R, dev, reduce=False)
    ex = S.ExplicitLoss(R, dev)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            out = fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, t.item()

    with torch.no_grad():
        per_pair_iou, ms_iou = timed(lambda: iou(true, pred))
        loss, ms_ex = timed(lambda: ex(true, pred))
    inter, union = iou.counts(true, pred)
    g_iou = D.global_iou(inter, union)
    g_loss = D.global_mean(loss, b1 - b0)
    if world > 1:
        parts = [torch.empty_like(per_pair_iou) for _ in range(world)] if rank == 0 else None
        dist.gather(per_pair_iou, parts, dst=0)            # equal shards at the BASELINE sizes
        if rank == 0:
            per_pair_iou = torch.cat(parts)
    if rank == 0:
        n_ex = ex._n
        line = json.dumps({"harness": "sweep", "n_gpus": world, "pairs": args.pairs, "render_size": R,
                           "iou_ms": ms_iou, "iou_gpoints_per_s": args.pairs * R ** 3 / ms_iou / 1e6,
                           "explicit_fwd_ms": ms_ex, "explicit_gpoints_per_s": args.pairs * n_ex ** 3 / ms_ex / 1e6,
                           "batch_iou": g_iou.item(), "mean_explicit_loss": g_loss.item(),
                           "per_pair_iou_mean": per_pair_iou.mean().item(), "per_pair_iou_count": per_pair_iou.numel()})
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
