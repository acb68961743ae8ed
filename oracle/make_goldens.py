"""Freeze outputs of the UNMODIFIED reference into tests/golden/ (run in the dev container only).

    python -m oracle.make_goldens

TEST INFRASTRUCTURE (see oracle/__init__.py).  Imports /root/reference/torch/classes.py through
oracle/ref_import.py, evaluates the four loss classes on CPU and stores inputs + outputs as
compressed .npz files.  The GPU box has no /root/reference, so these files are what the
``-m gpu`` tests and the oracle self-check compare against.

Cases
  fixtures.npz   the reference's own data: data/example_imgs/*.bmp + labels.txt,
                 classes.py:458-461 (IoU smoke), visu.py:77 (fixed GT vector)
  random_*.npz   seeded randsq()+randquat() inputs (visu.py:55-56, quaternion.py:139-145), SURVEY 8d
  edge.npz       params outside the clamps, on the clamp boundary, axis-aligned q with t on a grid
                 plane (zero fix-up path, classes.py:126,171-173), non-unit q, fp64 inputs
  edge_grazing.npz   one big object on a 24^3 grid at tau = 3 whose columns graze the surface for most of their length
                 (the parity fuzz's seed 14 case 68: dozens of gradient-carrying points per column)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import                      # noqa: E402
from oracle.sq_oracle import random_params, perturbed_params   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CPU = torch.device("cpu")


def _grad(fn, true, pred):
    pred = pred.clone().requires_grad_(True)
    loss = fn(true, pred)
    loss.backward()
    return loss.detach().double().numpy(), pred.grad.detach().double().numpy()


def fixtures(rc):
    imgs, labels = ref_import.example_fixtures()
    imgs_t, lab_t = torch.tensor(imgs), torch.tensor(labels)
    roll = lab_t.roll(1, 0)
    out = {"imgs_u8": np.round(imgs * 255).astype(np.uint8), "labels": labels}
    crit = rc.ImplicitLoss(64, CPU, 1.5, 260)
    out["implicit64_loss"], out["implicit64_grad"] = _grad(crit, imgs_t, lab_t)
    out["implicit64_roll_loss"], out["implicit64_roll_grad"] = _grad(crit, imgs_t, roll)
    out["implicit64_per_sample"] = np.array(
        [crit(imgs_t[i:i + 1], lab_t[i:i + 1]).item() for i in range(10)])
    out["implicit64_per_sample_next"] = np.array(
        [crit(imgs_t[i:i + 1], lab_t[(i + 1) % 10][None]).item() for i in range(10)])
    out["implicit64_depth"] = crit.depth_projection(lab_t).numpy()
    crit32 = rc.ImplicitLoss(32, CPU)                 # default tau=1, sharpness=100
    out["implicit32_default_loss"], out["implicit32_default_grad"] = _grad(crit32, imgs_t, roll)
    ex = rc.ExplicitLoss(32, CPU)
    out["explicit32_same_loss"] = ex(lab_t, lab_t).double().numpy()
    out["explicit32_roll_loss"], out["explicit32_roll_grad"] = _grad(ex, lab_t, roll)
    out["iou64_same"] = rc.IoUAccuracy(64, CPU)(lab_t, lab_t).double().numpy()
    out["iou64_roll"] = rc.IoUAccuracy(64, CPU)(lab_t, roll).double().numpy()
    out["iou64_roll_per_sample"] = rc.IoUAccuracy(64, CPU, reduce=False)(lab_t, roll).double().numpy()
    acc = rc.IoUAccuracy(64, CPU)
    a_bin, b_bin = acc.ins_outs(lab_t) <= 1, acc.ins_outs(roll) <= 1
    out["iou64_roll_inter"] = (a_bin & b_bin).flatten(1).sum(1).numpy()
    out["iou64_roll_union"] = (a_bin | b_bin).flatten(1).sum(1).numpy()
    ls = rc.LeastSquares(64, CPU)
    out["lsq64_loss"], out["lsq64_grad"] = _grad(ls, imgs_t, lab_t)
    out["lsq64_roll_loss"], out["lsq64_roll_grad"] = _grad(ls, imgs_t, roll)
    # classes.py:458-461 : identical true/pred -> IoU exactly 1
    a1, a2, a3, e1, e2 = 28.985552 / 255, 61.850255 / 255, 68.976172 / 255, 0.215097, 0.275022
    t1, t2, t3 = 137.818167 / 255, 94.702536 / 255, 118.771105 / 255
    main = np.array([[a1, a2, a3, e1, e2, t1, t2, t3, 0.699625, 0.378123, -0.090419, -0.599476]])
    out["main_params"] = main
    out["main_iou64"] = rc.IoUAccuracy(64, CPU)(torch.tensor(main), torch.tensor(main)).double().numpy()
    # visu.py:77 fixed GT vector against a seeded random prediction, ExplicitLoss(32) + IoU(128), fp64 leaf
    visu_true = np.array([[0.17840092, 0.29169756, 0.19272356, 0.564326, 0.850042, 0.5160052, 0.51887995,
                           0.41229093, 0.468217, 0.567843, -0.355409, 0.576204]])
    visu_pred = random_params(1, 7, torch.float64).numpy()
    out["visu_true"], out["visu_pred"] = visu_true, visu_pred
    out["visu_explicit32_loss"], out["visu_explicit32_grad"] = _grad(
        rc.ExplicitLoss(32, CPU), torch.tensor(visu_true), torch.tensor(visu_pred))
    out["visu_iou128"] = rc.IoUAccuracy(128, CPU)(torch.tensor(visu_true), torch.tensor(visu_pred)).double().numpy()
    np.savez_compressed(os.path.join(OUT, "fixtures.npz"), **out)
    print("fixtures:", {k: (v.shape if v.ndim else float(v)) for k, v in out.items() if k != "imgs_u8"})


def synthetic_depth(rc, params, size, seed):
    """Depth targets in [0,1): the reference's own soft render at `size`, like SURVEY 8d."""
    with torch.no_grad():
        return rc.ImplicitLoss(size, CPU, 1.5, 260).depth_projection(params).float().unsqueeze(1)


def random_cases(rc):
    for seed, B, R in ((1, 6, 16), (2, 4, 32), (3, 2, 64), (4, 8, 8)):
        true = random_params(B, seed)
        pred_far = random_params(B, seed + 100)
        pred_near = perturbed_params(true, seed)
        out = {"true": true.numpy(), "pred_far": pred_far.numpy(), "pred_near": pred_near.numpy(), "R": np.array(R)}
        img = synthetic_depth(rc, true, 4 * R, seed)      # nearest-resize stride 4, like 256 -> 64
        out["img"] = img.numpy()
        for tag, pred in (("far", pred_far), ("near", pred_near)):
            for name, crit in (("implicit_t15_k260", rc.ImplicitLoss(R, CPU, 1.5, 260)),
                               ("implicit_default", rc.ImplicitLoss(R, CPU))):
                out[f"{name}_{tag}_loss"], out[f"{name}_{tag}_grad"] = _grad(crit, img, pred)
            out[f"explicit_{tag}_loss"], out[f"explicit_{tag}_grad"] = _grad(rc.ExplicitLoss(R, CPU), true, pred)
            acc = rc.IoUAccuracy(R, CPU)
            a_bin, b_bin = acc.ins_outs(true) <= 1, acc.ins_outs(pred) <= 1
            out[f"iou_{tag}_inter"] = (a_bin & b_bin).flatten(1).sum(1).numpy()
            out[f"iou_{tag}_union"] = (a_bin | b_bin).flatten(1).sum(1).numpy()
            out[f"iou_{tag}"] = acc(true, pred).double().numpy()
            out[f"lsq_{tag}_loss"], out[f"lsq_{tag}_grad"] = _grad(rc.LeastSquares(R, CPU), img, pred)
        out["depth_true"] = rc.ImplicitLoss(R, CPU, 1.5, 260).depth_projection(true).numpy()
        out["occupancy_true"] = rc.ExplicitLoss(R, CPU).occupancy(true[:2]).numpy() if R <= 16 else np.zeros(0)
        np.savez_compressed(os.path.join(OUT, f"random_s{seed}_b{B}_r{R}.npz"), **out)
        print(f"random seed={seed} B={B} R={R}: implicit {float(out['implicit_t15_k260_near_loss']):.6g} "
              f"explicit {float(out['explicit_near_loss']):.6g} iou {float(out['iou_near']):.6g}")


def edge_cases(rc):
    R = 16
    base = random_params(8, 11, torch.float64)
    p = base.clone()
    p[0, 0:3] = torch.tensor([0.01, 1.7, 0.05])        # a below / above / on the clamp boundary
    p[1, 3:5] = torch.tensor([0.05, 1.0])              # e below / on boundary
    p[2, 5:8] = torch.tensor([-0.2, 1.3, 0.0])         # t below / above / on boundary
    p[3, 8:12] = torch.tensor([0.0, 0.0, 0.0, 1.0])    # axis aligned ...
    p[3, 5:8] = torch.tensor([8 / 15, 5 / 15, 0.4])    # ... with t on linspace(0,1,16) planes -> exact zeros
    p[4, 8:12] = p[4, 8:12] * 1.3                      # non-unit quaternion (never normalised, quaternion.py:46-67)
    p[5, 8:12] = torch.tensor([0.0, 0.0, 0.7071067811865476, 0.7071067811865476])   # 90 deg about z
    p[6, 3:5] = torch.tensor([1.0, 0.1])               # extreme exponent ratio e2/e1 = 0.1
    p[7, 3:5] = torch.tensor([0.1, 1.0])               # extreme exponent ratio e2/e1 = 10
    true = random_params(8, 12, torch.float64)
    img = synthetic_depth(rc, true.float(), 64, 0).double()
    out = {"pred": p.numpy(), "true": true.numpy(), "img": img.numpy(), "R": np.array(R)}
    out["implicit_loss"], out["implicit_grad"] = _grad(rc.ImplicitLoss(R, CPU, 1.5, 260), img.float(), p)
    out["implicit_soft_loss"], out["implicit_soft_grad"] = _grad(rc.ImplicitLoss(R, CPU, 1.0, 20), img.float(), p)
    out["explicit_loss"], out["explicit_grad"] = _grad(rc.ExplicitLoss(R, CPU), true, p)
    out["explicit_swapped_loss"], out["explicit_swapped_grad"] = _grad(rc.ExplicitLoss(R, CPU), p, true)
    # per-sample losses so a single bad row is visible
    ex, im = rc.ExplicitLoss(R, CPU), rc.ImplicitLoss(R, CPU, 1.5, 260)
    out["explicit_per_sample"] = np.array([ex(true[i:i + 1], p[i:i + 1]).item() for i in range(8)])
    out["implicit_per_sample"] = np.array([im(img[i:i + 1].float(), p[i:i + 1]).item() for i in range(8)])
    acc = rc.IoUAccuracy(R, CPU)
    a_bin, b_bin = acc.ins_outs(true) <= 1, acc.ins_outs(p) <= 1
    out["iou_inter"] = (a_bin & b_bin).flatten(1).sum(1).numpy()
    out["iou_union"] = (a_bin | b_bin).flatten(1).sum(1).numpy()
    # odd render sizes: arange(0,1+1/R,1/R) has R+2 entries for R=24 (SURVEY 7 hard parts)
    t24, p24 = random_params(2, 21), random_params(2, 22)
    out["true24"], out["pred24"] = t24.numpy(), p24.numpy()
    out["explicit24_loss"], out["explicit24_grad"] = _grad(rc.ExplicitLoss(24, CPU), t24, p24)
    out["explicit24_n"] = np.array(rc.ExplicitLoss(24, CPU).xyz.shape[1])
    np.savez_compressed(os.path.join(OUT, "edge.npz"), **out)
    print("edge: implicit", float(out["implicit_loss"]), "explicit", float(out["explicit_loss"]),
          "n(R=24) =", int(out["explicit24_n"]))


def zero_plane_cases(rc):
    """Axis-aligned rotations with the position exactly on a grid plane ALONG Z (the walk direction of the kernels): the
    reference's `s^2 == 0 -> 1e-4` fix-up (classes.py:171-173, 261-263) then fires on a whole plane.  Found by
    tests/tools/parity_fuzz_other.py (q = identity, t_z clamped to 1 = the last coordinate of every grid)."""
    R = 16
    base = random_params(6, 31, torch.float64)
    p = base.clone()
    ident = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64)
    p[0, 8:12] = ident; p[0, 5:8] = torch.tensor([0.41, 0.47, 1.4])             # t_z clamped to 1.0
    p[0, 3:5] = torch.tensor([0.97, 0.5])                                        # e1 near 1: the zero plane's C = 1e-4^(1/e1) counts
    p[1, 8:12] = ident; p[1, 5:8] = torch.tensor([0.5, 0.52, 9 / 15])            # on a linspace(0,1,16) plane
    p[2, 8:12] = ident; p[2, 5:8] = torch.tensor([0.52, 0.5, 0.5])               # on an arange(0,1+1/16,1/16) plane
    # half turns: M = diag(1,-1,-1), diag(-1,1,-1) -- every row has ONE non-zero entry, so the zeros do not depend on the
    # summation order of the reference's einsum.  (Quarter turns need sqrt(1/2): their rows keep an entry of ~1e-16 and
    # the reference's zeros then are rounding noise of its matmul -- not reproducible)
    p[3, 8:12] = torch.tensor([1.0, 0.0, 0.0, 0.0], dtype=torch.float64)
    p[3, 5:8] = torch.tensor([0.45, 0.55, 1.0])
    p[4, 8:12] = torch.tensor([0.0, 1.0, 0.0, 0.0], dtype=torch.float64)
    p[4, 5:8] = torch.tensor([0.5, 0.5, 0.5]); p[4, 3:5] = torch.tensor([1.0, 1.0])
    p[5, 8:12] = ident * 1.5; p[5, 5:8] = torch.tensor([0.48, 0.5, 1.0])         # non-unit but axis-aligned
    true = random_params(6, 32, torch.float64)
    img = synthetic_depth(rc, true.float(), 64, 0).double()
    out = {"pred": p.numpy(), "true": true.numpy(), "img": img.numpy(), "R": np.array(R)}
    out["implicit_loss"], out["implicit_grad"] = _grad(rc.ImplicitLoss(R, CPU, 1.5, 260), img.float(), p)
    out["implicit_soft_loss"], out["implicit_soft_grad"] = _grad(rc.ImplicitLoss(R, CPU, 1.0, 20), img.float(), p)
    out["explicit_loss"], out["explicit_grad"] = _grad(rc.ExplicitLoss(R, CPU), true, p)
    out["explicit_swapped_loss"], out["explicit_swapped_grad"] = _grad(rc.ExplicitLoss(R, CPU), p, true)
    out["implicit_per_sample"] = np.array([rc.ImplicitLoss(R, CPU, 1.5, 260)(img[i:i + 1].float(), p[i:i + 1]).item() for i in range(6)])
    out["explicit_per_sample"] = np.array([rc.ExplicitLoss(R, CPU)(true[i:i + 1], p[i:i + 1]).item() for i in range(6)])
    np.savez_compressed(os.path.join(OUT, "edge_zero_planes.npz"), **out)
    print("zero planes: implicit", float(out["implicit_loss"]), "explicit", float(out["explicit_loss"]))


def grazing_case(rc):
    """The one configuration of the parity fuzz (tests/tools/parity_fuzz.py --seed 14, case 68) that was outside the gradient
    tolerance until the kernels pooled the gradient points of a column group: a single object that fills a 24^3 grid, tau = 3,
    sigmoid sharpness 260, depth image at 73 x 73; its columns graze the surface for up to all 24 planes."""
    R = 24
    true = torch.tensor([[0.5378291606903076, 0.3434811234474182, 0.5789921283721924, 0.2919308543205261, 0.46377694606781006,
                          0.6156783699989319, 0.5951224565505981, 0.6020408272743225,
                          0.9947512745857239, -0.03643030673265457, -0.09445077925920486, -0.014893271960318089]])
    pred = torch.tensor([[0.48910823464393616, 0.41221243143081665, 0.7377411723136902, 0.24222418665885925, 0.4850635826587677,
                          0.5418038368225098, 0.5384869575500488, 0.6187753677368164,
                          0.9637041687965393, -0.2668421268463135, 0.008339019492268562, 0.00011745292431442067]])
    img = synthetic_depth(rc, true, 73, 0)
    out = {"pred": pred.numpy(), "true": true.numpy(), "img": img.numpy(), "R": np.array(R), "tau": np.array(3.0), "k": np.array(260.0)}
    out["implicit_loss"], out["implicit_grad"] = _grad(rc.ImplicitLoss(R, CPU, 3.0, 260), img, pred)
    np.savez_compressed(os.path.join(OUT, "edge_grazing.npz"), **out)
    print("grazing: implicit", float(out["implicit_loss"]))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    rc, _ = ref_import.load()
    fixtures(rc)
    random_cases(rc)
    edge_cases(rc)
    zero_plane_cases(rc)
    grazing_case(rc)


if __name__ == "__main__":
    main()
