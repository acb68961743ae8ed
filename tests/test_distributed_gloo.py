"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): batch sharding by sample + the two scalar reductions.

The per-rank loss/gradient values come from the oracle here (no GPU in this test); what is tested is the host-side
logic of sq_recovery_b200/distributed.py: shards tile the batch, the all-reduced mean equals the single-process
loss, per-rank gradients rescaled by shard/global size concatenate to the single-process gradient, and IoU
counters reduce to the batch-wide ratio (torch/classes.py:437-439).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sq_oracle as O
from sq_recovery_b200 import distributed as D


def test_shard_range_tiles_the_batch():
    for total in (1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B, R = 5, 8                                             # odd batch: unequal shards
    true, pred = O.random_params(B, 1), O.random_params(B, 2)
    b, e = D.shard_range(B, rank, world)
    p = pred[b:e].clone().requires_grad_(True)
    crit = O.ExplicitLoss(R, "cpu")
    local = crit(true[b:e], p)
    local.backward()
    g_mean = D.global_mean(local, e - b)
    grad = D.scale_local_grad(p.grad, e - b, B)
    inter, union = O.IoUAccuracy(R, "cpu").counts(true[b:e], pred[b:e])
    iou = D.global_iou(inter, union)
    q.put((rank, g_mean.item(), grad.double().numpy(), iou.item()))
    dist.destroy_process_group()


def test_two_ranks_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, R = 5, 8
    true, pred = O.random_params(B, 1), O.random_params(B, 2)
    p = pred.clone().requires_grad_(True)
    ref = O.ExplicitLoss(R, "cpu")(true, p)
    ref.backward()
    iou = O.IoUAccuracy(R, "cpu")(true, pred).item()
    import numpy as np
    for rank, g_mean, grad, g_iou in out:
        assert abs(g_mean - ref.item()) < 1e-12
        assert abs(g_iou - iou) < 1e-7
    np.testing.assert_allclose(np.concatenate([o[2] for o in out]), p.grad.double().numpy(), rtol=1e-6, atol=1e-9)   # fp32 leaves
