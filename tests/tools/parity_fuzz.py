"""Random configurations of ImplicitLoss (render size, tau, sharpness, object size range, fp32 / fp64 parameters) against the
fp64 oracle: loss rtol 1e-5, every gradient entry within rtol 1e-4 / atol 1e-6 (MAE-tie samples excluded as in the tests).

    python tests/tools/parity_fuzz.py [--cases 40] [--seed 0]
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
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/parity_fuzz.json")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.set_num_threads(os.cpu_count())
    rs = np.random.RandomState(args.seed)
    rows, worst = [], 0.0
    for case in range(args.cases):
        R = int(rs.choice([8, 12, 16, 20, 24, 32, 40, 48, 64, 96]))
        tau = float(rs.choice([0.5, 1.0, 1.5, 3.0]))
        k = float(rs.choice([20.0, 100.0, 260.0, 500.0]))
        lo = float(rs.choice([0.05, 0.1, 0.3, 0.5]))
        size_range = (lo, min(1.0, lo + float(rs.choice([0.1, 0.2, 0.5]))))
        dtype = torch.float64 if rs.rand() < 0.25 else torch.float32
        B = int(rs.choice([1, 3, 8])) if R <= 48 else 2
        seed = 5000 + case
        true = inputs.random_params(B, seed, dtype, size_range=size_range)
        pred = inputs.perturbed_params(true, seed, sigma=float(rs.choice([0.01, 0.03, 0.1]))) if rs.rand() < 0.6 \
            else inputs.random_params(B, seed + 1, dtype, size_range=size_range)
        Rimg = int(rs.choice([R, 2 * R, 3 * R + 1]))
        with torch.no_grad():
            img = O.ImplicitLoss(Rimg, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
        oc = O.ImplicitLoss(R, "cpu", tau, k)
        p = pred.clone().requires_grad_(True)
        ref = oc(img, p); ref.backward()
        pg = pred.to(dev).requires_grad_(True)
        l = S.ImplicitLoss(R, dev, tau, k)(img.to(dev), pg); l.backward()
        with torch.no_grad():
            d = oc.depth_projection(pred); t = oc.resize(img)[:, 0].double()
        keep = ~(((d - t).abs() < 1e-6) & (d > 1e-5)).flatten(1).any(dim=1).numpy()
        rg = p.grad.double().numpy()
        e = (np.abs(pg.grad.double().cpu().numpy() - rg) / (1e-6 + 1e-4 * np.abs(rg))).max(axis=1)
        gerr = float(e[keep].max()) if keep.any() else 0.0
        lerr = abs(l.item() - ref.item()) / max(abs(ref.item()), 1e-30)
        row = {"case": case, "R": R, "tau": tau, "k": k, "sizes": size_range, "dtype": str(dtype).split(".")[-1], "B": B, "Rimg": Rimg,
               "loss": ref.item(), "loss_rel_err": lerr, "grad_err_tol": gerr, "kept": int(keep.sum())}
        rows.append(row)
        worst = max(worst, gerr)
        flag = "  <-- " if (gerr > 1.0 or lerr > 1e-5) else ""
        print(json.dumps(row) + flag, flush=True)
    print(f"worst gradient error {worst:.3f}x tolerance; worst loss rel err {max(r['loss_rel_err'] for r in rows):.2e}; "
          f"{sum(r['grad_err_tol'] > 1 or r['loss_rel_err'] > 1e-5 for r in rows)} case(s) outside")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
