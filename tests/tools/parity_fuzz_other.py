"""Random configurations of ExplicitLoss / IoUAccuracy / LeastSquares against the fp64 oracle, including parameters outside
the clamp ranges, non-unit quaternions and odd render sizes: explicit loss rtol 1e-5 and gradients rtol 1e-4 / atol 1e-6,
IoU voxel counts exact, least squares at the reference's own fp32 precision.

    python tests/tools/parity_fuzz_other.py [--cases 60] [--seed 0]
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


def wild(p, rs):
    """push some parameters outside the clamp ranges / off the unit sphere"""
    p = p.clone()
    for b in range(p.shape[0]):
        if rs.rand() < 0.3: p[b, rs.randint(0, 3)] = float(rs.choice([0.01, 0.04, 1.2, 1.6]))
        if rs.rand() < 0.3: p[b, 3 + rs.randint(0, 2)] = float(rs.choice([0.06, 0.12, 1.3, 1.9]))
        if rs.rand() < 0.3: p[b, 5 + rs.randint(0, 3)] = float(rs.choice([-0.3, 0.0, 1.0, 1.4]))
        if rs.rand() < 0.3: p[b, 8:12] *= float(rs.choice([0.5, 1.7]))
        if rs.rand() < 0.1: p[b, 8:12] = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=p.dtype)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/parity_fuzz_other.json")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.set_num_threads(os.cpu_count())
    rs = np.random.RandomState(args.seed)
    rows, bad = [], 0
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    for case in range(args.cases):
        R = int(rs.choice([8, 12, 16, 17, 24, 32, 33, 48, 64]))
        B = int(rs.choice([1, 3, 7])) if R <= 33 else 2
        dtype = torch.float64 if rs.rand() < 0.25 else torch.float32
        lo = float(rs.choice([0.05, 0.1, 0.3]))
        size_range = (lo, min(1.0, lo + float(rs.choice([0.1, 0.3, 0.6]))))
        seed = 9000 + case
        true = wild(inputs.random_params(B, seed, dtype, size_range=size_range), rs)
        pred = wild(inputs.random_params(B, seed + 1, dtype, size_range=size_range), rs) if rs.rand() < 0.5 \
            else wild(inputs.perturbed_params(true, seed, sigma=0.05), rs)
        row = {"case": case, "R": R, "B": B, "dtype": str(dtype).split(".")[-1]}
        # ExplicitLoss
        p = pred.clone().requires_grad_(True)
        ref = O.ExplicitLoss(R, "cpu")(true, p); ref.backward()
        pg = pred.to(dev).requires_grad_(True)
        l = S.ExplicitLoss(R, dev)(true.to(dev), pg); l.backward()
        rg = p.grad.double().numpy()
        row["explicit_loss_rel"] = abs(l.item() - ref.item()) / max(abs(ref.item()), 1e-30) if ref.item() != 0 else abs(l.item())
        row["explicit_grad_tol"] = float((np.abs(pg.grad.double().cpu().numpy() - rg) / (1e-6 + 1e-4 * np.abs(rg))).max())
        # IoU (no clamp in the reference: keep the shapes positive)
        ti, pi = true.clone(), pred.clone()
        ti[:, 3:5] = ti[:, 3:5].clamp(min=0.06); pi[:, 3:5] = pi[:, 3:5].clamp(min=0.06)
        ti[:, 0:3] = ti[:, 0:3].clamp(min=0.01); pi[:, 0:3] = pi[:, 0:3].clamp(min=0.01)
        i, u = O.IoUAccuracy(R, "cpu").counts(ti, pi)
        i2, u2 = S.IoUAccuracy(R, dev).counts(ti.to(dev), pi.to(dev))
        row["iou_exact"] = bool(torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu()))
        # LeastSquares (fp32 in the reference)
        with torch.no_grad():
            img = O.ImplicitLoss(2 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
        pf = pred.float()
        p = pf.clone().requires_grad_(True)
        ref = O.LeastSquares(R, "cpu")(img, p); ref.backward()
        pg = pf.to(dev).requires_grad_(True)
        l = S.LeastSquares(R, dev)(img.to(dev), pg); l.backward()
        rg, gg = p.grad.double().numpy(), pg.grad.double().cpu().numpy()
        row["lsq_loss_rel"] = abs(l.item() - ref.item()) / max(abs(ref.item()), 1e-30)
        # rows whose fp32 reference gradient overflowed (inf / nan: pow with exponent 2 / 0.1 on a far point) are not comparable
        fin = np.isfinite(rg).all(axis=1)
        row["lsq_ref_rows_nonfinite"] = int((~fin).sum())
        ok_l = np.isfinite(gg).all() and (np.abs(gg - rg)[fin] <= (2e-3 * np.abs(rg).max(axis=1, keepdims=True) + 4e-3 * np.abs(rg) + 1e-4)[fin]).all()
        row["lsq_grad_ok"] = bool(ok_l) if np.isfinite(ref.item()) else None
        out = row["explicit_loss_rel"] > 1e-5 or row["explicit_grad_tol"] > 1 or not row["iou_exact"] or \
            (np.isfinite(ref.item()) and (row["lsq_loss_rel"] > 2e-4 or not ok_l))
        bad += bool(out)
        if out:                                             # keep the inputs and both gradients of a failing case for a closer look
            np.savez(os.path.join(os.path.dirname(args.out) or ".", f"fuzz_other_fail_s{args.seed}_c{case}.npz"), R=R, true=true.numpy(),
                     pred=pred.numpy(), img=img.numpy(), lsq_ref_grad=rg, lsq_grad=gg, lsq_ref=ref.item(), lsq=l.item())
        rows.append(row)
        print(json.dumps(row) + ("  <-- " if out else ""), flush=True)
    print(f"{bad} case(s) outside; worst explicit gradient {max(r['explicit_grad_tol'] for r in rows):.3f}x, "
          f"explicit loss {max(r['explicit_loss_rel'] for r in rows):.2e}, iou exact in {sum(r['iou_exact'] for r in rows)}/{len(rows)}")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
