"""The parity fuzz of tests/tools/parity_fuzz.py on the HOST build of the kernels' per-point core (tests/emu) instead of the GPU:
random ImplicitLoss configurations against the fp64 oracle, loss rtol 1e-5, every gradient entry within rtol 1e-4 / atol 1e-6.
libm replaces the MUFU approximations here, so this bounds what the ALGORITHM leaves (culling and weight cuts, the pool of
gradient points and its overflow path, fp64 refinement, first-order suffix correction), not the hardware error.  Runs on the CPU.

    python tests/tools/emu_fuzz.py [--cases 60] [--seed 0] [--pool 0]      (--pool n: n slots per 32 columns instead of 480)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
from oracle import sq_oracle as O          # noqa: E402  (checker)
from sq_recovery_b200 import inputs        # noqa: E402
import emu_lib as E                        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/emu_fuzz.json")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    E.lib().emu_set_pool(args.pool)
    rs = np.random.RandomState(args.seed)
    rows, worst = [], 0.0
    for case in range(args.cases):
        R = int(rs.choice([8, 12, 16, 20, 24, 32, 40]))
        tau = float(rs.choice([0.5, 1.0, 1.5, 3.0]))
        k = float(rs.choice([20.0, 100.0, 260.0, 500.0]))
        lo = float(rs.choice([0.05, 0.1, 0.3, 0.5]))
        size_range = (lo, min(1.0, lo + float(rs.choice([0.1, 0.2, 0.5]))))
        B = int(rs.choice([1, 3, 8])) if R <= 24 else 2
        seed = 7000 + 1000 * args.seed + case
        true = inputs.random_params(B, seed, torch.float64, size_range=size_range)
        pred = inputs.perturbed_params(true, seed, sigma=float(rs.choice([0.01, 0.03, 0.1]))) if rs.rand() < 0.6 \
            else inputs.random_params(B, seed + 1, torch.float64, size_range=size_range)
        Rimg = int(rs.choice([R, 2 * R, 3 * R + 1]))
        with torch.no_grad():
            img = O.ImplicitLoss(Rimg, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
        oc = O.ImplicitLoss(R, "cpu", tau, k)
        p = pred.clone().requires_grad_(True)
        ref = oc(img, p); ref.backward()
        with torch.no_grad():
            d = oc.depth_projection(pred); t = oc.resize(img)[:, 0].double()
        keep = ~(((d - t).abs() < 1e-6) & (d > 1e-5)).flatten(1).any(dim=1).numpy()
        l, g, _ = E.implicit(pred.numpy(), oc.resize(img)[:, 0].numpy(), R, 1 / (R - 1), 1e-4, tau, k)
        rg = p.grad.numpy()
        e = (np.abs(g - rg) / (1e-6 + 1e-4 * np.abs(rg))).max(axis=1)
        gerr = float(e[keep].max()) if keep.any() else 0.0
        lerr = abs(l - ref.item()) / max(abs(ref.item()), 1e-30)
        row = {"case": case, "seed": args.seed, "R": R, "tau": tau, "k": k, "sizes": size_range, "B": B, "Rimg": Rimg,
               "loss_rel_err": lerr, "grad_err_tol": gerr, "kept": int(keep.sum())}
        rows.append(row)
        worst = max(worst, gerr)
        # (absolute floor 3e-8: a loss that small is the fp32 quantisation of the target image itself -- an object that covers the
        # whole image and a prediction that renders the same constant depth: fuzz seed 14 case 26, reference loss 7.5e-10, here 0)
        row["outside"] = bool(gerr > 1.0 or abs(l - ref.item()) > 1e-5 * abs(ref.item()) + 3e-8)
        if row["outside"]:
            print(json.dumps(row) + "  <--", flush=True)
    print(f"seed {args.seed}: worst gradient error {worst:.3f}x tolerance; worst loss rel err {max(r['loss_rel_err'] for r in rows if r['loss_rel_err'] < 0.5):.2e}; "
          f"{sum(r['outside'] for r in rows)} of {len(rows)} case(s) outside")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    E.lib().emu_set_pool(0)


if __name__ == "__main__":
    main()
