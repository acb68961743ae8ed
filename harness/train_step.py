#!/usr/bin/env python
"""Train-step harness modelled on torch/train.py:80-100 (BASELINE configs 3 and 4).

    python harness/train_step.py [--batch 128] [--steps 30] [--loss b200|oracle-cuda]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 harness/train_step.py --batch 1024

One step = zero_grad, CNN forward, ImplicitLoss(64, dev, 1.5, 260)(depth, pred), backward, Adam step (lr 1e-4).
Depth maps are synthetic: soft renders (R=256) of random true parameters.  With torchrun the model is wrapped in
DistributedDataParallel (NCCL), the GLOBAL batch is sharded by sample, and the loss needs no collective of its own.
`--loss oracle-cuda` runs the same step with the oracle's per-sample torch-op loop on the GPU (the reference's own
GPU code path, fp64) for the "share of step time in the loss" comparison.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness.model import SQRegressor                      # noqa: E402
import sq_recovery_b200 as S                              # noqa: E402
from sq_recovery_b200 import distributed as D            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128, help="GLOBAL batch (config 3: 128, config 4: 1024)")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--render", type=int, default=64)
    ap.add_argument("--loss", default="b200", choices=["b200", "b200-heads", "oracle-cuda"],
                    help="b200-heads: head activations fused into the loss kernels (ImplicitLoss.from_heads)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from oracle import sq_oracle as O                      # input distributions; and the comparator when asked for
    b0, b1 = D.shard_range(args.batch, rank, world)
    nb = b1 - b0
    true = O.random_params(args.batch, 0)[b0:b1].to(dev)
    images = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
    torch.manual_seed(0)
    net = SQRegressor().to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    if args.loss in ("b200", "b200-heads"):
        crit = S.ImplicitLoss(args.render, dev, 1.5, 260)
    else:
        crit = O.ImplicitLoss(args.render, dev, 1.5, 260, form="loop")
    iou = S.IoUAccuracy(args.render, dev)

    def step():
        opt.zero_grad(set_to_none=True)
        if args.loss == "b200-heads":
            pred = model(images, raw=True)
            loss = crit.from_heads(images, pred)
        else:
            pred = model(images)
            loss = crit(images, pred)
        loss.backward()
        opt.step()
        return loss, pred

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev[0].record()
    for _ in range(args.steps):
        loss, pred = step()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    if args.loss == "b200-heads":                          # raw head outputs -> parameters, for the timings / IoU below
        pred = torch.cat([torch.sigmoid(pred[:, :8]), pred[:, 8:] / pred[:, 8:].norm(dim=1, keepdim=True)], dim=1)
    # loss share: time loss forward+backward alone on the same predictions
    p = pred.detach().requires_grad_(True)
    for _ in range(3):
        crit(images, p).backward()
    ev[2].record()
    reps = 10 if args.loss.startswith("b200") else 2
    for _ in range(reps):
        crit(images, p).backward()
    ev[3].record()
    torch.cuda.synchronize()
    loss_ms = ev[2].elapsed_time(ev[3]) / reps
    t = torch.tensor([ms, loss_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    with torch.no_grad():
        inter, union = iou.counts(true, pred.detach())
    g_loss = D.global_mean(loss, nb)
    g_iou = D.global_iou(inter, union)
    if rank == 0:
        print(json.dumps({"harness": "train_step", "loss_impl": args.loss, "n_gpus": world, "global_batch": args.batch,
                          "render_size": args.render, "steps": args.steps, "ms_per_step": t[0].item(),
                          "steps_per_s": 1e3 / t[0].item(), "loss_fwd_bwd_ms": t[1].item(),
                          "loss_share_of_step": t[1].item() / t[0].item(), "loss": g_loss.item(), "val_iou": g_iou.item(),
                          "cnn_params": sum(p.numel() for p in net.parameters())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
