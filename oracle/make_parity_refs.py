"""Freeze the fp64 oracle's losses and gradients for the parity sweep (run in the dev container; CPU only).

    python -m oracle.make_parity_refs [--seeds 36]

TEST INFRASTRUCTURE (see oracle/__init__.py).  For every cell of the sweep -- ImplicitLoss at render sizes 16 / 32 / 64
with (tau, sharpness) = (1.5, 260) (torch/train.py:64) and the defaults (1, 100), 36 seeds x 2 prediction styles
(independent random / perturbed ground truth) -- it stores the inputs exactly as the CUDA path must see them
(predictions fp32; depth maps already at the render size, so the nearest resize of classes.py:286 is the identity: the
oracle renders the true parameters at 4R and the loss samples every 4th pixel) and the oracle's loss, gradient and
"unambiguous" mask (samples whose MAE sign is decided below fp32 resolution are excluded from the GRADIENT comparison,
tests/test_gpu_parity.py).  Output: tests/golden/parity_sweep_refs.npz (a few MB).  tests/tools/parity_sweep.py and
tests/test_gpu_parity.py::test_parity_sweep_every_sample evaluate the CUDA path against it on the GPU box, where
recomputing the oracle would take minutes per run.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sq_oracle as O      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "parity_sweep_refs.npz")
CELLS = ((16, 16), (32, 8), (64, 4))                        # (render size, batch per call)
SETTINGS = ((1.5, 260.0), (1.0, 100.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=36)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"seeds": np.arange(100, 100 + args.seeds)}
    for R, B in CELLS:
        t0 = time.time()
        preds, imgs = [], []
        res = {s: {"loss": [], "grad": [], "keep": []} for s in SETTINGS}
        for seed in range(100, 100 + args.seeds):
            true = O.random_params(B, seed)
            with torch.no_grad():
                big = O.ImplicitLoss(4 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
            img = F.interpolate(big, size=(R, R), mode="nearest")           # what classes.py:286 hands the loss
            for pred in (O.random_params(B, seed + 1000), O.perturbed_params(true, seed)):
                preds.append(pred.numpy()); imgs.append(img.numpy())
                for (tau, k) in SETTINGS:
                    oc = O.ImplicitLoss(R, "cpu", tau, k)
                    p = pred.clone().requires_grad_(True)
                    ref = oc(img, p); ref.backward()
                    with torch.no_grad():
                        d = oc.depth_projection(pred); t = img[:, 0].double()
                    keep = ~(((d - t).abs() < 1e-6) & (d > 1e-5)).flatten(1).any(dim=1).numpy()
                    res[(tau, k)]["loss"].append(ref.item())
                    res[(tau, k)]["grad"].append(p.grad.double().numpy())
                    res[(tau, k)]["keep"].append(keep)
        out[f"R{R}_pred"] = np.stack(preds).astype(np.float32)             # (calls, B, 12)
        out[f"R{R}_img"] = np.stack(imgs).astype(np.float32)               # (calls, B, 1, R, R)
        for (tau, k), r in res.items():
            tag = f"R{R}_t{tau:g}_k{k:g}"
            out[tag + "_loss"] = np.array(r["loss"])
            out[tag + "_grad"] = np.stack(r["grad"])
            out[tag + "_keep"] = np.stack(r["keep"])
        print(f"R={R}: {len(preds)} calls of {B} samples, {time.time() - t0:.0f} s", flush=True)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
