"""Error statistics of the CUDA ImplicitLoss against the frozen fp64 oracle vectors (run on the GPU box).

    python tests/tools/parity_sweep.py [--out gpurun_out/parity_sweep.json] [--explicit]

tests/golden/parity_sweep_refs.npz (oracle/make_parity_refs.py: 36 seeds x 2 prediction styles per cell) holds the
inputs and the oracle's loss / gradient for ImplicitLoss at R = 16 / 32 / 64 with (tau, k) = (1.5, 260) and (1, 100).
For each cell this reports the distribution, over samples, of the worst gradient error in units of the north-star
tolerance (rtol 1e-4, atol 1e-6) and the loss relative error.  SQ_LIBSQLOSS=<variant .so> evaluates an experimental
build.  --explicit adds ExplicitLoss / IoU against the live oracle (slow on the CPU).  DESIGN.md quotes these.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sq_recovery_b200 as S               # noqa: E402

REFS = os.path.join(ROOT, "tests", "golden", "parity_sweep_refs.npz")


def tol_units(g, ref):
    return (np.abs(g - ref) / (1e-6 + 1e-4 * np.abs(ref))).max(axis=1)


def implicit_cells(dev):
    refs = np.load(REFS)
    rows = []
    for R in (16, 32, 64):
        preds, imgs = refs[f"R{R}_pred"], refs[f"R{R}_img"]
        for (tau, k) in ((1.5, 260.0), (1.0, 100.0)):
            tag = f"R{R}_t{tau:g}_k{k:g}"
            crit = S.ImplicitLoss(R, dev, tau, k)
            errs, lerrs, ties, worst = [], [], 0, None
            for c in range(preds.shape[0]):
                pg = torch.tensor(preds[c]).to(dev).requires_grad_(True)
                l = crit(torch.tensor(imgs[c]).to(dev), pg)
                l.backward()
                keep = refs[tag + "_keep"][c]
                e = tol_units(pg.grad.double().cpu().numpy(), refs[tag + "_grad"][c])
                ties += int((~keep).sum())
                if keep.any() and (worst is None or e[keep].max() > worst[0]):
                    b = int(np.argmax(np.where(keep, e, -1)))
                    worst = (float(e[keep].max()), c, b)
                errs.append(e[keep])
                lerrs.append(abs(l.item() - refs[tag + "_loss"][c]) / abs(refs[tag + "_loss"][c]))
            e = np.concatenate(errs)
            rows.append({"loss": "implicit", "R": R, "tau": tau, "k": k, "samples": int(e.size),
                         "grad_err_tol_median": float(np.median(e)), "grad_err_tol_p95": float(np.percentile(e, 95)),
                         "grad_err_tol_p999": float(np.percentile(e, 99.9)),
                         "grad_err_tol_max": float(e.max()), "frac_over_tol": float((e > 1).mean()),
                         "loss_rel_max": float(max(lerrs)), "samples_excluded_mae_tie": ties,
                         "worst_call_sample": list(worst[1:]) if worst else None})
            print(rows[-1], flush=True)
    return rows


def explicit_cells(dev, seeds):
    from oracle import sq_oracle as O      # checker
    rows = []
    for R, B in ((16, 16), (32, 8), (64, 4)):
        errs, lerrs = [], []
        for seed in range(100, 100 + seeds):
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
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/parity_sweep.json")
    ap.add_argument("--explicit", action="store_true")
    ap.add_argument("--seeds", type=int, default=36, help="seeds of the --explicit part")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.set_num_threads(os.cpu_count())
    rows = implicit_cells(dev)
    if args.explicit:
        rows += explicit_cells(dev, args.seeds)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"library": os.environ.get("SQ_LIBSQLOSS", "sq_recovery_b200/libsqloss.so"), "cells": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
