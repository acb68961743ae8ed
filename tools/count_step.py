"""What the implicit column kernel did on one call of a workload, from the counting build (libsqloss_count.so, -DSQ_COUNT).

    python tools/count_step.py [--dense] [--batch 256] [--render 64]

Prints one JSON object: warp plane steps, deal-out rounds, queued / refined points, walked fraction of the grid, and the
MUFU-pipe (XU) warp instructions these imply (the per-event costs are read off the kernel source, see bench.py)."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sq_recovery_b200 import inputs, counting      # noqa: E402
import sq_recovery_b200 as S                        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--render", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = inputs.DENSE_SIZE_RANGE if args.dense else inputs.SIZE_RANGE
    true = inputs.random_params(args.batch, args.seed, size_range=rng)
    pred = inputs.perturbed_params(true, 7 + args.seed).to(dev)
    img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).contiguous()
    out = counting.implicit_counts(img, pred, args.render, 1.5, 260.0, want_grad=True)
    out["forward_only"] = counting.implicit_counts(img, pred, args.render, 1.5, 260.0, want_grad=False)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
