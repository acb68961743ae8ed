"""Host build of the kernels' per-point core (csrc/sq_core.cuh via tests/emu) against the frozen reference outputs.

This checks, on the CPU, the algebra the CUDA kernels run: log-sum-exp forward, analytic backward, two-moment
suffix-sum trick, z-range culling, finalize Jacobians (quaternion, clamps).  MUFU approximations are replaced by
libm here, so it bounds the formula error, not the hardware approximation error (the -m gpu tests do that).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_lib as E          # noqa: E402


def tol(g, ref, rtol=1e-4, atol=1e-6):
    return (np.abs(g - ref) / (atol + rtol * np.abs(ref))).max()


@pytest.mark.parametrize("fname", ["random_s1_b6_r16.npz", "random_s2_b4_r32.npz", "random_s4_b8_r8.npz"])
def test_emulated_kernels_match_reference(fname):
    g = load_golden(fname)
    R = int(g["R"])
    tgt = F.interpolate(torch.tensor(g["img"]), size=(R, R), mode="nearest")[:, 0].numpy()
    for tag in ("far", "near"):
        pred = g[f"pred_{tag}"]
        for name, (tau, k) in (("implicit_t15_k260", (1.5, 260.0)), ("implicit_default", (1.0, 100.0))):
            l, gr, _ = E.implicit(pred, tgt, R, 1 / (R - 1), 1e-4, tau, k)
            assert abs(l - g[f"{name}_{tag}_loss"]) <= 1e-5 * g[f"{name}_{tag}_loss"]
            assert tol(gr, g[f"{name}_{tag}_grad"]) <= 1.0
        l, gr = E.explicit(g["true"], pred, R + 1, 1 / R, 1e-4)
        assert abs(l - g[f"explicit_{tag}_loss"]) <= 1e-5 * g[f"explicit_{tag}_loss"]
        assert tol(gr, g[f"explicit_{tag}_grad"]) <= 1.0
        i, u = E.iou(g["true"], pred, R, 1 / (R - 1))
        assert (i == g[f"iou_{tag}_inter"]).all() and (u == g[f"iou_{tag}_union"]).all()
    _, _, d = E.implicit(g["true"], None, R, 1 / (R - 1), 1e-4, 1.5, 260.0, want_grad=False, want_depth=True)
    np.testing.assert_allclose(d, g["depth_true"], atol=2e-5)


def test_emulated_edge_cases(edge_golden):
    g = edge_golden
    R = int(g["R"])
    tgt = F.interpolate(torch.tensor(g["img"]).float(), size=(R, R), mode="nearest")[:, 0].numpy()
    l, gr, _ = E.implicit(g["pred"], tgt, R, 1 / (R - 1), 1e-4, 1.5, 260.0)
    assert abs(l - g["implicit_loss"]) <= 1e-5 * g["implicit_loss"] and tol(gr, g["implicit_grad"]) <= 1.0
    assert gr[0, 0] == 0 and gr[0, 1] == 0 and gr[1, 3] == 0 and gr[2, 5] == 0 and gr[2, 6] == 0   # clamp masks
    l, gr = E.explicit(g["true"], g["pred"], R + 1, 1 / R, 1e-4)
    assert abs(l - g["explicit_loss"]) <= 1e-5 * g["explicit_loss"] and tol(gr, g["explicit_grad"]) <= 1.0
    l, gr = E.explicit(g["true24"], g["pred24"], 26, 1 / 24, 1e-4)                                 # R = 24 -> n = 26
    assert abs(l - g["explicit24_loss"]) <= 1e-5 * g["explicit24_loss"] and tol(gr, g["explicit24_grad"]) <= 1.0
    i, u = E.iou(g["true"], g["pred"], R, 1 / (R - 1))
    assert (i == g["iou_inter"]).all() and (u == g["iou_union"]).all()


def test_emulated_fixture_least_squares(fixtures_golden):
    from oracle import sq_oracle as O
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    pts = O.LeastSquares(64, "cpu").points(imgs)
    off = np.cumsum([0] + [p.shape[1] for p in pts])
    P = torch.cat(pts, 1).T.contiguous().numpy()
    l, gr = E.lsq(g["labels"], P, off)
    assert abs(l - g["lsq64_loss"]) <= 1e-4 * g["lsq64_loss"] and tol(gr, g["lsq64_grad"], 1e-3, 1e-5) <= 1.0
    i, u = E.iou(g["labels"], np.roll(g["labels"], 1, 0), 64, 1 / 63)
    assert (i == g["iou64_roll_inter"]).all() and (u == g["iou64_roll_union"]).all()
