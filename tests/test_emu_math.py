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


@pytest.mark.parametrize("kind", ["implicit", "explicit", "iou"])
def test_culling_never_drops_occupancy(kind):
    """The column kernels skip grid points outside column_range() and whole work items the plan kernel proves empty.
    Both claims are checked here against the fp64 inside-outside function on random, clamped-extreme and off-centre
    superquadrics: a skipped point must have F >= 1 + bits / (k log2 e), i.e. occupancy below 2^-bits (DESIGN.md
    "culling"; IoU: F > 1).  Host build of the same header the kernels compile."""
    from oracle import sq_oracle as O
    log2e = 1.4426950408889634
    rs = np.random.RandomState(3)
    p = O.random_params(24, 17).double().numpy()
    p[0, 0:3] = 0.05; p[1, 0:3] = 1.0; p[2, 3:5] = 0.1; p[3, 3:5] = 1.0; p[4, 3:5] = [0.1, 1.0]; p[5, 3:5] = [1.0, 0.1]
    p[6, 5:8] = [0.0, 0.0, 0.0]; p[7, 5:8] = [1.0, 1.0, 1.0]; p[8, 8:12] = [0, 0, 0, 1]; p[9, 8:12] = [0.5, 0.5, 0.5, 0.5]
    p[10, 8:12] *= 1.7                                                  # non-unit quaternion
    if kind == "iou":
        p[11, 3:5] = [1.6, 1.3]; p[12, 5:8] = [1.4, -0.3, 0.5]; p[13, 0:3] = [0.02, 0.4, 1.5]     # no clamp for IoU
        n, step, z0, clamp, bound, fmin = 32, 1 / 31, 0.0, 0, 1.001, 1.0 + 1e-12
    elif kind == "implicit":
        kl = 260 * log2e
        n, step, z0, clamp = 32, 1 / 31, 1e-4, 1
        bound, fmin = float(np.sqrt((1 + 40 / kl) * 1.002)), 1 + 40 / kl
    else:
        kl = 5 * log2e
        n, step, z0, clamp = 24, 1 / 23, 1e-4, 1
        bound, fmin = float(np.sqrt((1 + 24 / kl) * 1.002)), 1 + 24 / kl
    bad, empty, total = E.check_culling(p, n, step, z0, clamp, bound, fmin)
    assert bad == 0
    if kind != "explicit":                                             # the proof is not vacuous: it catches a good part
        assert empty > 0.2 * total, (empty, total)


@pytest.mark.parametrize("n", [9, 12, 17, 20, 33])
@pytest.mark.parametrize("kind", ["implicit", "iou"])
def test_proven_empty_groups_in_the_x_fastest_layout(kind, n):
    """Grids whose size is not a multiple of 8 hand out 32 consecutive columns of the x-fastest order per work item; for
    n < 32 such a group wraps over several rows.  The plan kernel's "proven empty" must hold for every column of the
    group's real footprint (round-1 advisor finding: the footprint assumed two rows)."""
    from oracle import sq_oracle as O
    log2e = 1.4426950408889634
    rs = np.random.RandomState(n)
    p = O.random_params(400, 300 + n).double().numpy()
    p[:, 0:3] = rs.uniform(0.05, 0.5, (400, 3))
    p[:, 5:8] = rs.uniform(0.1, 0.9, (400, 3))
    if kind == "iou":
        bad, empty, total = E.check_culling(p, n, 1 / (n - 1), 0.0, 0, 1.001, 1.0 + 1e-12)
    else:
        kl = 260 * log2e
        bad, empty, total = E.check_culling(p, n, 1 / (n - 1), 1e-4, 1, float(np.sqrt((1 + 40 / kl) * 1.002)), 1 + 40 / kl)
    assert bad == 0 and total > 0


@pytest.mark.parametrize("n", [16, 20, 64])
def test_walk_estimate_never_empties_a_live_group(n):
    """The plan kernel orders the implicit kernels' work by the planes a walk really visits (footprint_walk).  Only the
    order depends on it -- except that an estimate of 0 means "proven empty, never processed": it must be 0 exactly where
    the culled plane count is, and never above it.  Small objects, big objects (the early exit cuts those), any tau."""
    from sq_recovery_b200 import inputs as I
    log2e = 1.4426950408889634
    kl = 260 * log2e
    bound = float(np.sqrt((1 + 40 / kl) * 1.002))
    for seed, size_range, tau in ((1, I.SIZE_RANGE, 1.5), (2, I.DENSE_SIZE_RANGE, 1.5), (3, I.DENSE_SIZE_RANGE, 0.5), (4, (0.05, 1.0), 3.0)):
        p = I.random_params(64, seed, size_range=size_range).double().numpy()
        bad, cut, live = E.check_walk_estimate(p, n, 1 / (n - 1), 1e-4, bound, 32.0 / (tau * log2e))
        assert bad == 0 and live > 0
        if size_range == I.DENSE_SIZE_RANGE and tau == 1.5 and n == 64:
            assert cut > 0.3 * live, (cut, live)            # objects that fill the grid: most groups leave early


def test_emulated_zero_planes():
    """Axis-aligned rotations with t_z exactly on a grid plane (tests/golden/edge_zero_planes.npz, frozen from the reference):
    the reference's exact-zero fix-up fires on a whole z plane; the kernels' affine walk along z reproduces it through
    Sample::zpack (zero_planes()).  Sample by sample, so that a single bad row is visible."""
    g = load_golden("edge_zero_planes.npz")
    R = int(g["R"])
    tgt = F.interpolate(torch.tensor(g["img"]).float(), size=(R, R), mode="nearest")[:, 0].numpy()
    for tau, k, name in ((1.5, 260.0, "implicit"), (1.0, 20.0, "implicit_soft")):
        l, gr, _ = E.implicit(g["pred"], tgt, R, 1 / (R - 1), 1e-4, tau, k)
        assert abs(l - g[f"{name}_loss"]) <= 1e-5 * g[f"{name}_loss"]
        assert tol(gr, g[f"{name}_grad"]) <= 1.0, (name, np.abs(gr - g[f"{name}_grad"]).max(axis=1))
    l, gr = E.explicit(g["true"], g["pred"], R + 1, 1 / R, 1e-4)
    assert abs(l - g["explicit_loss"]) <= 1e-5 * g["explicit_loss"] and tol(gr, g["explicit_grad"]) <= 1.0
    l, gr = E.explicit(g["pred"], g["true"], R + 1, 1 / R, 1e-4)
    assert abs(l - g["explicit_swapped_loss"]) <= 1e-5 * g["explicit_swapped_loss"] and tol(gr, g["explicit_swapped_grad"]) <= 1.0
    for i in range(len(g["pred"])):
        l, _, _ = E.implicit(g["pred"][i:i + 1], tgt[i:i + 1], R, 1 / (R - 1), 1e-4, 1.5, 260.0, want_grad=False)
        assert abs(l - g["implicit_per_sample"][i]) <= 1e-5 * g["implicit_per_sample"][i], i
        l, _ = E.explicit(g["true"][i:i + 1], g["pred"][i:i + 1], R + 1, 1 / R, 1e-4, want_grad=False)
        assert abs(l - g["explicit_per_sample"][i]) <= 1e-5 * g["explicit_per_sample"][i], i


def test_emulated_pool_overflow_and_on_the_spot_paths():
    """The warp's pool of gradient-carrying points (sq_core.cuh BwdQueue): with all 480 slots nothing overflows on these
    inputs; with 8 slots per 32-column group most points take the on-the-spot two-moment path (no fp64 refinement, host libm
    accuracy); with the pool off all do.  All three must agree with the reference within the tolerance, and the first two
    must really have taken different paths."""
    g = load_golden("random_s2_b4_r32.npz")
    R = int(g["R"])
    tgt = F.interpolate(torch.tensor(g["img"]), size=(R, R), mode="nearest")[:, 0].numpy()
    lib = E.lib()
    res = {}
    try:
        for name, (pool, queue) in (("full", (0, 1)), ("small", (8, 1)), ("off", (0, 0))):
            lib.emu_set_pool(pool)
            lib.emu_set_queue(queue)
            l, gr, _ = E.implicit(g["pred_near"], tgt, R, 1 / (R - 1), 1e-4, 1.5, 260.0)
            assert abs(l - g["implicit_t15_k260_near_loss"]) <= 1e-5 * g["implicit_t15_k260_near_loss"], name
            assert tol(gr, g["implicit_t15_k260_near_grad"]) <= 1.0, name
            res[name] = gr
    finally:
        lib.emu_set_pool(0)
        lib.emu_set_queue(1)
    assert not np.array_equal(res["full"], res["small"]) and not np.array_equal(res["small"], res["off"])


def test_emulated_grazing_columns_need_the_whole_pool():
    """tests/golden/edge_grazing.npz: a single object that fills a 24^3 grid at tau = 3, whose columns graze the surface for
    most of their 24 planes (the parity fuzz's seed 14 case 68).  With the whole pool every gradient point is refined and the
    gradient is within the tolerance; with a pool of 16 slots per 32 columns -- most points then take the unrefined
    on-the-spot path, as the columns beyond 15 points did with the per-lane queues of the first two thirds of round 2 -- it is
    not."""
    g = load_golden("edge_grazing.npz")
    R, tau, k = int(g["R"]), float(g["tau"]), float(g["k"])
    tgt = F.interpolate(torch.tensor(g["img"]).float(), size=(R, R), mode="nearest")[:, 0].numpy()
    lib = E.lib()
    try:
        lib.emu_set_pool(0)
        l, gr, _ = E.implicit(g["pred"], tgt, R, 1 / (R - 1), 1e-4, tau, k)
        assert abs(l - g["implicit_loss"]) <= 1e-5 * g["implicit_loss"]
        full = tol(gr, g["implicit_grad"])
        lib.emu_set_pool(16)
        _, gr, _ = E.implicit(g["pred"], tgt, R, 1 / (R - 1), 1e-4, tau, k)
        small = tol(gr, g["implicit_grad"])
    finally:
        lib.emu_set_pool(0)
    assert full <= 1.0, full
    assert small > 1.3 * full and small > 1.0, (small, full)


def test_pool_entry_tags_round_trip():
    """Lane and link ride in the 13 low mantissa bits of an entry's plane index: every plane of grids up to 1024, plane 0's
    non-integer index (restored from the sample), every lane, links up to "none" (255)."""
    import ctypes
    lib = E.lib()
    lib.emu_tag_roundtrip.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
    for cf0 in (0.0, 1e-4 * 63, 1e-4 * 1023, 0.5):
        for c in list(range(1, 40)) + [63, 64, 127, 128, 255, 256, 511, 512, 1000, 1023]:
            for lane, link in ((0, 0), (31, 255), (17, 254), (5, 129)):
                assert lib.emu_tag_roundtrip(float(c), cf0, lane, link) == 0, (c, cf0, lane, link)
        for lane, link in ((0, 0), (31, 255), (9, 77)):
            assert lib.emu_tag_roundtrip(np.float32(cf0), np.float32(cf0), lane, link) == 0, (cf0, lane, link)      # plane 0
