#!/usr/bin/env python
"""Training harness with the structure of torch/train.py:72-175 (BASELINE configs 3 and 4).

    python harness/train_step.py [--batch 128] [--epochs 2] [--steps-per-epoch 20] [--loss b200|b200-heads]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 harness/train_step.py --batch 1024

Per epoch, like the reference: a training pass (train.py:80-100: batch to the device, zero_grad, CNN forward,
``ImplicitLoss(64, dev, 1.5, 260)(data, pred)``, backward, Adam step, ``loss.item()``), a validation pass under
``torch.no_grad()`` with the loss and ``IoUAccuracy(64)`` (train.py:135-154), ``ReduceLROnPlateau(patience=25)`` on the
validation loss (train.py:52,161; without ``verbose``, which torch 2.11 rejects) and a checkpoint whenever the validation
loss improves (train.py:164-171) in the dict format of torch/helpers.py:42-48 (``epoch``, ``model_state_dict``,
``optimizer_state_dict``, ``loss`` = {"loss", "val_loss", "val_acc"} histories); ``--resume`` continues from it
(train.py:56-58).  The dataset is synthetic -- soft renders (256 x 256) of seeded random superquadrics, generated on the
GPU by harness/make_dataset.py, split 0.9 / 0.1 like ``H5Dataset(train_split=0.9)`` -- and lives in pinned host memory;
every step copies its batch to the device (train.py:83).

With torchrun the model is wrapped in DistributedDataParallel (NCCL), each GLOBAL batch is sharded by sample, the loss
needs no collective of its own; the logged loss / IoU are reduced with sq_recovery_b200.distributed.  Rank 0 prints one
JSON line: steps/s, the share of the step spent in the loss, the cost of the CNN-gradient all-reduce and how much of it is
exposed (DDP step against the same step under ``no_sync()``).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness.model import SQRegressor                      # noqa: E402
from harness import make_dataset                           # noqa: E402
import sq_recovery_b200 as S                              # noqa: E402
from sq_recovery_b200 import distributed as D            # noqa: E402
from sq_recovery_b200 import inputs                        # noqa: E402


def save_checkpoint(path, epoch, model, optimizer, history):
    """torch/helpers.py:42-48: the same four keys, so helpers.load_model reads it."""
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict(), "loss": history}, path)


def load_checkpoint(path, model, optimizer, device):
    """torch/helpers.py:51-68 without the plotting."""
    ck = torch.load(path, map_location=device, weights_only=False)
    model.load_state_dict(ck["model_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return ck["epoch"], ck["loss"]


class SyntheticSplit:
    """One rank's shard of a synthetic dataset: `batches` global batches of `global_batch` samples, of which this rank
    holds rows [b0, b1) of every batch.  Images in pinned host memory (float32, (N,1,256,256) like H5Dataset's "sq"),
    labels (N,12) next to them."""

    def __init__(self, batches, global_batch, rank, world, seed, dev, size=256):
        self.b0, self.b1 = D.shard_range(global_batch, rank, world)
        self.local = self.b1 - self.b0
        self.batches = batches
        labels = [inputs.random_params(global_batch, seed + i)[self.b0:self.b1] for i in range(batches)]
        self.labels = torch.cat(labels) if batches else torch.empty(0, 12)
        self.images = torch.empty((batches * self.local, 1, size, size), dtype=torch.float32).pin_memory()
        for i in range(batches):
            sl = slice(i * self.local, (i + 1) * self.local)
            self.images[sl].copy_(make_dataset.render(self.labels[sl], dev, size=size))
        torch.cuda.synchronize()
        self.labels = self.labels.pin_memory()

    def batch(self, i, dev):
        sl = slice(i * self.local, (i + 1) * self.local)
        return self.images[sl].to(dev, non_blocking=True), self.labels[sl].to(dev, non_blocking=True)


def predict(model, data, fused_heads):
    """train.py:88-89 (the model here returns the concatenated row); raw head outputs for the fused-heads loss."""
    return model(data, raw=True) if fused_heads else model(data)


def loss_of(crit, data, pred, fused_heads):
    return crit.from_heads(data, pred) if fused_heads else crit(data, pred)


def params_of(pred, fused_heads):
    if not fused_heads:
        return pred
    return torch.cat([torch.sigmoid(pred[:, :8]), pred[:, 8:] / pred[:, 8:].norm(dim=1, keepdim=True)], dim=1)


def train_epoch(model, opt, crit, split, dev, fused_heads, losses, timer=None):
    """train.py:76-103."""
    model.train()
    for i in range(split.batches):
        data, _ = split.batch(i, dev)                      # train.py:83
        opt.zero_grad()                                    # train.py:85
        pred = predict(model, data, fused_heads)           # train.py:88-89
        loss = loss_of(crit, data, pred, fused_heads)      # train.py:92
        loss.backward()                                    # train.py:93
        opt.step()                                         # train.py:100
        losses.append(loss)                                # .item() of train.py:103 is taken once per epoch (no host sync per step)
        if timer is not None:
            timer()
    return losses


def validate(model, crit, iou, split, dev, fused_heads):
    """train.py:131-154: loss and IoU per validation batch under no_grad; returns (sum of local mean losses, batches,
    intersection, union) for the cross-rank reduction."""
    model.eval()
    loss_sum = torch.zeros((), dtype=torch.float64, device=dev)
    inter = torch.zeros((), dtype=torch.int64, device=dev)
    union = torch.zeros((), dtype=torch.int64, device=dev)
    per_batch_acc = []
    with torch.no_grad():
        for i in range(split.batches):
            data, true_labels = split.batch(i, dev)
            pred = predict(model, data, fused_heads)
            loss_sum += loss_of(crit, data, pred, fused_heads)
            bi, bu = iou.counts(true_labels, params_of(pred, fused_heads))
            inter += bi.sum(); union += bu.sum()
            per_batch_acc.append(D.global_iou(bi, bu))     # train.py:146 per batch, over the GLOBAL batch
    return loss_sum, inter, union, per_batch_acc


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128, help="GLOBAL batch (config 3: 128, config 4: 1024)")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps-per-epoch", type=int, default=20)
    ap.add_argument("--val-batches", type=int, default=2)
    ap.add_argument("--render", type=int, default=64)
    ap.add_argument("--lr", type=float, default=1e-4)       # train.py:40
    ap.add_argument("--patience", type=int, default=25)     # train.py:52
    ap.add_argument("--loss", default="b200", choices=["b200", "b200-heads"],
                    help="b200-heads: head activations fused into the loss kernels (ImplicitLoss.from_heads)")
    ap.add_argument("--checkpoint", default="")
    ap.add_argument("--resume", action="store_true")        # train.py:45 CONTINUE_TRAINING
    ap.add_argument("--out", default="", help="also append the JSON line to this file")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    fused = args.loss == "b200-heads"

    train = SyntheticSplit(args.steps_per_epoch, args.batch, rank, world, 0, dev)
    val = SyntheticSplit(args.val_batches, args.batch, rank, world, 100000, dev)
    torch.manual_seed(0)
    net = SQRegressor().to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=0)          # train.py:51
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=args.patience)  # train.py:52
    crit = S.ImplicitLoss(args.render, dev, 1.5, 260)       # train.py:64
    iou = S.IoUAccuracy(render_size=args.render, device=dev, full=True)              # train.py:66
    history = {"loss": [], "val_loss": [], "val_acc": []}
    start_epoch, best_val = 0, None
    if args.resume and args.checkpoint and os.path.exists(args.checkpoint):
        start_epoch, history = load_checkpoint(args.checkpoint, net, opt, dev)
        start_epoch += 1
        best_val = min(history["val_loss"]) if history["val_loss"] else None

    # warm-up outside the timing: cuDNN autotune, workspace allocation, NCCL channels
    warm = []
    saved = {k: v.clone() for k, v in net.state_dict().items()}
    saved_opt = opt.state_dict()
    for _ in range(2):
        train_epoch(model, opt, crit, SyntheticSplitView(train, 3), dev, fused, warm)
    net.load_state_dict(saved); opt.load_state_dict(saved_opt)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    step_ms, saved_at = [], []
    for epoch in range(start_epoch, start_epoch + args.epochs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        losses = []
        e0.record()
        train_epoch(model, opt, crit, train, dev, fused, losses)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1) / train.batches)
        mean_loss = D.global_mean(torch.stack(losses).mean(), train.local).item()
        history["loss"].append(mean_loss)
        vl, inter, union, accs = validate(model, crit, iou, val, dev, fused)
        val_loss = D.global_mean(vl / max(val.batches, 1), val.local).item()
        history["val_loss"].append(val_loss)
        history["val_acc"].append([a.item() for a in accs])                      # train.py:159 keeps the per-batch list
        sched.step(val_loss)                                                     # train.py:161
        if best_val is None or val_loss < best_val:                              # train.py:164-171
            best_val = val_loss
            if args.checkpoint and rank == 0:
                save_checkpoint(args.checkpoint, epoch, net, opt, history)
            saved_at.append(epoch)

    # ---- where the step time goes
    data, _ = train.batch(0, dev)
    pred = predict(model, data, fused).detach()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        loss_of(crit, data, pred.requires_grad_(True), fused).backward()
    ev[0].record()
    for _ in range(10):
        loss_of(crit, data, pred.requires_grad_(True), fused).backward()
    ev[1].record(); torch.cuda.synchronize()
    loss_ms = ev[0].elapsed_time(ev[1]) / 10
    n_params = sum(p.numel() for p in net.parameters())
    allreduce_ms = nosync_ms = None
    if world > 1:
        flat = torch.zeros(n_params, dtype=torch.float32, device=dev)            # the CNN gradient: 45.5 MB
        for _ in range(3):
            dist.all_reduce(flat)
        ev[0].record()
        for _ in range(10):
            dist.all_reduce(flat)
        ev[1].record(); torch.cuda.synchronize()
        allreduce_ms = ev[0].elapsed_time(ev[1]) / 10
        with model.no_sync():                                                    # the same step without the all-reduce
            timed = []
            train_epoch(model, opt, crit, SyntheticSplitView(train, 3), dev, fused, timed)
            ev[0].record()
            train_epoch(model, opt, crit, SyntheticSplitView(train, min(10, train.batches)), dev, fused, timed)
            ev[1].record(); torch.cuda.synchronize()
        nosync_ms = ev[0].elapsed_time(ev[1]) / min(10, train.batches)
    t = torch.tensor([float(np.mean(step_ms[-1:])), loss_ms, allreduce_ms or 0.0, nosync_ms or 0.0],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t[0].item()
        line = {"harness": "train_step", "loss_impl": args.loss, "n_gpus": world, "global_batch": args.batch,
                "batch_per_gpu": train.local, "render_size": args.render, "epochs": args.epochs,
                "steps_per_epoch": train.batches, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                "samples_per_s": args.batch * 1e3 / ms, "ms_per_step_by_epoch": step_ms,
                "loss_fwd_bwd_ms": t[1].item(), "loss_share_of_step": t[1].item() / ms,
                "loss_gpoints_per_s": world * train.local * args.render ** 3 / t[1].item() / 1e6,
                "train_loss": history["loss"], "val_loss": history["val_loss"],
                "val_iou": [float(np.mean(a)) for a in history["val_acc"]], "lr": opt.param_groups[0]["lr"],
                "checkpoint_saved_at_epochs": saved_at, "cnn_params": n_params,
                "data": "synthetic renders in pinned host memory, copied to the device every step (train.py:83)"}
        if world > 1:
            line.update({"grad_allreduce_bytes": 4 * n_params, "grad_allreduce_ms_standalone": t[2].item(),
                         "ms_per_step_no_sync": t[3].item(), "allreduce_exposed_ms": ms - t[3].item(),
                         "allreduce_overlap": "DDP buckets the gradient (25 MB buckets) and all-reduces each bucket on "
                                              "NCCL's stream while the backward of the earlier layers still runs; "
                                              "exposed = DDP step - no_sync step"})
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.barrier()
    return history


class SyntheticSplitView:
    """The first `batches` batches of a split (warm-up / short timing passes)."""

    def __init__(self, split, batches):
        self.split, self.batches, self.local = split, min(batches, split.batches), split.local

    def batch(self, i, dev):
        return self.split.batch(i, dev)


if __name__ == "__main__":
    main()
    if dist.is_initialized():
        dist.destroy_process_group()
