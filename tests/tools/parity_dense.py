"""Gradient / loss error of the CUDA ImplicitLoss against the fp64 oracle on LARGE objects (a ~ U(0.5, 1): the `dense`
workload of bench.py), where many columns graze a face for more planes than a lane's backward queue holds.

    python tests/tools/parity_dense.py [--seeds 6]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sq_oracle as O          # noqa: E402  (checker)
import sq_recovery_b200 as S               # noqa: E402
from sq_recovery_b200 import inputs        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=6)
    ap.add_argument("--out", default="gpurun_out/parity_dense.json")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.set_num_threads(os.cpu_count())
    rows = []
    for R, B in ((32, 8), (64, 4)):
        errs, lerrs = [], []
        for seed in range(300, 300 + args.seeds):
            true = inputs.random_params(B, seed, size_range=inputs.DENSE_SIZE_RANGE)
            for pred in (inputs.random_params(B, seed + 1000, size_range=inputs.DENSE_SIZE_RANGE), inputs.perturbed_params(true, seed)):
                with torch.no_grad():
                    img = O.ImplicitLoss(2 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
                oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
                p = pred.clone().requires_grad_(True)
                ref = oc(img, p); ref.backward()
                pg = pred.to(dev).requires_grad_(True)
                l = S.ImplicitLoss(R, dev, 1.5, 260)(img.to(dev), pg); l.backward()
                with torch.no_grad():
                    d = oc.depth_projection(pred); t = oc.resize(img)[:, 0].double()
                keep = ~(((d - t).abs() < 1e-6) & (d > 1e-5)).flatten(1).any(dim=1).numpy()
                e = (np.abs(pg.grad.double().cpu().numpy() - p.grad.double().numpy()) / (1e-6 + 1e-4 * np.abs(p.grad.double().numpy()))).max(axis=1)
                errs.append(e[keep]); lerrs.append(abs(l.item() - ref.item()) / abs(ref.item()))
        e = np.concatenate(errs)
        rows.append({"workload": "dense a~U(0.5,1)", "R": R, "samples": int(e.size), "grad_err_tol_median": float(np.median(e)),
                     "grad_err_tol_p95": float(np.percentile(e, 95)), "grad_err_tol_max": float(e.max()),
                     "frac_over_tol": float((e > 1).mean()), "loss_rel_max": float(max(lerrs))})
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
