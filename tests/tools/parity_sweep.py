"""Error statistics of the CUDA losses against the CPU oracle over many seeds (run on the GPU box).

    python tools/parity_sweep.py [--out gpurun_out/parity_sweep.json]

For each (loss, R, sharpness) it reports the distribution, over samples, of the worst gradient error in units
of the north-star tolerance (rtol 1e-4, atol 1e-6) and of the loss relative error.  DESIGN.md quotes these.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sq_oracle as O          # noqa: E402  (checker)
import sq_recovery_b200 as S               # noqa: E402


def tol_units(g, ref):
    return (np.abs(g - ref) / (1e-6 + 1e-4 * np.abs(ref))).max(axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/parity_sweep.json")
    ap.add_argument("--seeds", type=int, default=6)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.set_num_threads(os.cpu_count())
    rows = []
    for R, B in ((16, 16), (32, 8), (64, 4)):
        for (tau, k) in ((1.5, 260.0), (1.0, 100.0)):
            errs, lerrs, ties = [], [], 0
            t0 = time.time()
            for seed in range(100, 100 + args.seeds):
                true = O.random_params(B, seed)
                for pred in (O.random_params(B, seed + 1000), O.perturbed_params(true, seed)):
                    with torch.no_grad():
                        img = O.ImplicitLoss(4 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
                    p = pred.clone().requires_grad_(True)
                    oc = O.ImplicitLoss(R, "cpu", tau, k)
                    ref = oc(img, p); ref.backward()
                    pg = pred.to(dev).requires_grad_(True)
                    l = S.ImplicitLoss(R, dev, tau, k)(img.to(dev), pg); l.backward()
                    with torch.no_grad():       # MAE ties closer than fp32 resolves: sign(depth - target) undefined
                        d = oc.depth_projection(pred); t = oc.resize(img)[:, 0].double()
                    keep = ~(((d - t).abs() < 1e-6) & (d > 1e-5)).flatten(1).any(dim=1).numpy()
                    ties += int((~keep).sum())
                    errs.append(tol_units(pg.grad.double().cpu().numpy(), p.grad.double().numpy())[keep])
                    lerrs.append(abs(l.item() - ref.item()) / abs(ref.item()))
            e = np.concatenate(errs)
            rows.append({"loss": "implicit", "R": R, "tau": tau, "k": k, "samples": int(e.size),
                         "grad_err_tol_median": float(np.median(e)), "grad_err_tol_p95": float(np.percentile(e, 95)),
                         "grad_err_tol_max": float(e.max()), "frac_over_tol": float((e > 1).mean()),
                         "loss_rel_max": float(max(lerrs)), "samples_excluded_mae_tie": ties,
                         "seconds": time.time() - t0})
            print(rows[-1], flush=True)
        errs, lerrs = [], []
        for seed in range(100, 100 + args.seeds):
            true = O.random_params(B, seed)
            for pred in (O.random_params(B, seed + 1000), O.perturbed_params(true, seed)):
                p = pred.clone().requires_grad_(True)
                ref = O.ExplicitLoss(R, "cpu")(true, p); ref.backward()
                pg = pred.to(dev).requires_grad_(True)
                l = S.ExplicitLoss(R, dev)(true.to(dev), pg); l.backward()
                errs.append(tol_units(pg.grad.double().cpu().numpy(), p.grad.double().numpy()))
                lerrs.append(abs(l.item() - ref.item()) / abs(ref.item()))
                i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
                i2, u2 = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
                assert torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu()), "IoU counts differ"
        e = np.concatenate(errs)
        rows.append({"loss": "explicit", "R": R, "samples": int(e.size), "grad_err_tol_median": float(np.median(e)),
                     "grad_err_tol_p95": float(np.percentile(e, 95)), "grad_err_tol_max": float(e.max()),
                     "frac_over_tol": float((e > 1).mean()), "loss_rel_max": float(max(lerrs)), "iou_counts_exact": True})
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
