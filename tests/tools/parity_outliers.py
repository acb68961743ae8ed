"""Why do a few samples exceed the gradient tolerance at sigmoid_sharpness 260?  For every sample above 0.8x tolerance:
the worst entries, its shape parameters, and how close its closest pixel is to a tie |depth - target| = 0 (where the
MAE derivative sign(depth - target) flips; the kernel's fp32 depth is accurate to ~1e-6..1e-5 per pixel).

    python tools/parity_outliers.py [--R 32] [--seeds 36]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sq_oracle as O          # noqa: E402  (checker)
import sq_recovery_b200 as S               # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=32)
ap.add_argument("--seeds", type=int, default=36)
args = ap.parse_args()
R, B = args.R, {16: 16, 32: 8, 64: 4}[args.R]
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
names = ["a1", "a2", "a3", "e1", "e2", "t1", "t2", "t3", "qx", "qy", "qz", "qw"]
n_all = n_out = 0
closest_all, closest_out = [], []
for seed in range(100, 100 + args.seeds):
    true = O.random_params(B, seed)
    for style, pred in (("random", O.random_params(B, seed + 1000)), ("perturbed", O.perturbed_params(true, seed))):
        with torch.no_grad():
            img = O.ImplicitLoss(4 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
        p = pred.clone().requires_grad_(True)
        oc = O.ImplicitLoss(R, "cpu", 1.5, 260.0)
        ref = oc(img, p); ref.backward()
        pg = pred.to(dev).requires_grad_(True)
        crit = S.ImplicitLoss(R, dev, 1.5, 260.0)
        l = crit(img.to(dev), pg); l.backward()
        with torch.no_grad():
            d = oc.depth_projection(pred); t = oc.resize(img)[:, 0].double()
            dk = crit.depth_projection(pred.to(dev)).double().cpu()
        gap = torch.where(d > 1e-5, (d - t).abs(), torch.full_like(d, 1.0)).flatten(1).min(dim=1).values.numpy()
        flips = (((d - t) * (dk - t) < 0) & (d > 1e-5)).flatten(1).sum(dim=1).numpy()     # pixels whose sign differs
        g, gr = pg.grad.double().cpu().numpy(), p.grad.double().numpy()
        err = np.abs(g - gr) / (1e-6 + 1e-4 * np.abs(gr))
        for b in range(B):
            if gap[b] < 1e-6:
                continue                               # excluded by the tests' tie rule
            n_all += 1; closest_all.append(gap[b])
            if err[b].max() > 0.8:
                n_out += 1; closest_out.append(gap[b])
                worst = np.argsort(-err[b])[:3]
                print(f"seed {seed} {style:9s} sample {b}: " + ", ".join(f"{names[i]} {err[b][i]:.2f}x" for i in worst)
                      + f" | a {pred[b, :3].numpy().round(3)} e {pred[b, 3:5].numpy().round(3)} | closest pixel to a tie {gap[b]:.1e}"
                      + f" | pixels with flipped sign: {int(flips[b])}")
print(f"{n_out} of {n_all} samples above 0.8x tolerance; median closest-to-tie gap: outliers {np.median(closest_out) if closest_out else float('nan'):.1e}, all {np.median(closest_all):.1e}")
