"""Parity of the CUDA product (through the reference-facing classes and the C ABI) against the oracle.

All tests here need a B200 (``-m gpu``).  Tolerances are the north-star ones: loss rtol 1e-5, parameter
gradients rtol 1e-4 / atol 1e-6 (fp32 kernels vs the fp64 reference); IoU voxel counts are exact.
Nothing here reads /root/reference: the reference's outputs come from tests/golden/*.npz (frozen by
oracle/make_goldens.py) and from the oracle restatement evaluated on the CPU.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, random_golden_files
from oracle import sq_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL, GRAD_ATOL = 1e-4, 1e-6


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def S():
    import sq_recovery_b200 as pkg
    return pkg


def run(crit, true, pred, dev):
    """loss (float) and d loss / d pred (np.float64) through the public call + autograd."""
    true = torch.as_tensor(true).to(dev)
    pred = torch.as_tensor(pred).to(dev).clone().requires_grad_(True)
    loss = crit(true, pred)
    loss.backward()
    assert pred.grad is not None and pred.grad.dtype == pred.dtype and pred.grad.shape == pred.shape
    return loss.item(), pred.grad.double().cpu().numpy()


def check(loss, grad, ref_loss, ref_grad, loss_rtol=LOSS_RTOL, rtol=GRAD_RTOL, atol=GRAD_ATOL, what="", keep=None):
    """Loss within loss_rtol; EVERY gradient entry within rtol/atol (`keep`: per-sample mask, see unambiguous())."""
    ref_loss = float(ref_loss)
    assert abs(loss - ref_loss) <= loss_rtol * abs(ref_loss) + 1e-12, f"{what}: loss {loss} vs {ref_loss}"
    err = np.abs(grad - ref_grad) / (atol + rtol * np.abs(ref_grad))
    if keep is not None:
        err = err[np.asarray(keep)]
    if err.size == 0:
        return
    assert err.max() <= 1.0, f"{what}: worst gradient error {err.max():.2f}x tolerance at {np.unravel_index(err.argmax(), err.shape)}"


def unambiguous(oracle_crit, img, pred):
    """Per-sample mask: False where the MAE derivative sign(depth - target) (classes.py:292) is decided by less than
    fp32 can resolve.  |.| is not differentiable at a tie; a pixel whose fp64 render differs from its fp32 target by
    < 1e-6 flips that pixel's whole contribution, so such samples are excluded from the GRADIENT comparison (same
    spirit as the north-star's exclusion of points within 1e-6 of a clamp boundary).  The loss is still compared."""
    with torch.no_grad():
        d = oracle_crit.depth_projection(torch.as_tensor(pred))
        t = oracle_crit.resize(torch.as_tensor(img))[:, 0].double()
    tie = ((d - t).abs() < 1e-6) & (d > 1e-5)
    return ~tie.flatten(1).any(dim=1).numpy()


# ------------------------------------------------------------------ the reference's own fixtures
def test_fixture_implicit(fixtures_golden, dev, S):
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    lab = g["labels"]
    crit = S.ImplicitLoss(64, dev, 1.5, 260)
    l, gr = run(crit, imgs, lab, dev)
    check(l, gr, g["implicit64_loss"], g["implicit64_grad"], what="implicit64")
    assert abs(l - 0.007077226864072453) < 1e-5 * 0.0071
    l, gr = run(crit, imgs, np.roll(lab, 1, 0), dev)
    check(l, gr, g["implicit64_roll_loss"], g["implicit64_roll_grad"], what="implicit64 rolled labels")
    depth = crit.depth_projection(torch.tensor(lab).to(dev)).cpu().numpy()
    np.testing.assert_allclose(depth, g["implicit64_depth"], atol=2e-5)
    l, gr = run(S.ImplicitLoss(32, dev), imgs, np.roll(lab, 1, 0), dev)      # default tau / sharpness
    check(l, gr, g["implicit32_default_loss"], g["implicit32_default_grad"], what="implicit32 defaults")
    # per-sample: image i with its own label is small, with the next label >= 10x larger
    own = [crit(imgs[i:i + 1].to(dev), torch.tensor(lab[i:i + 1]).to(dev)).item() for i in range(10)]
    np.testing.assert_allclose(own, g["implicit64_per_sample"], rtol=1e-5)


def test_fixture_explicit_iou_lsq(fixtures_golden, dev, S):
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    lab = g["labels"]
    roll = np.roll(lab, 1, 0)
    ex = S.ExplicitLoss(32, dev)
    assert ex(torch.tensor(lab).to(dev), torch.tensor(lab).to(dev)).item() == 0.0
    l, gr = run(ex, lab, roll, dev)
    check(l, gr, g["explicit32_roll_loss"], g["explicit32_roll_grad"], what="explicit32")
    acc = S.IoUAccuracy(64, dev)
    t, p = torch.tensor(lab).to(dev), torch.tensor(roll).to(dev)
    assert acc(t, t).item() == 1.0
    i, u = acc.counts(t, p)
    assert (i.cpu().numpy() == g["iou64_roll_inter"]).all() and (u.cpu().numpy() == g["iou64_roll_union"]).all()
    assert abs(acc(t, p).item() - float(g["iou64_roll"])) < 1e-7
    per = S.IoUAccuracy(64, dev, reduce=False)(t, p)
    assert per.shape == (10,) and per.dtype == torch.float64
    np.testing.assert_allclose(per.cpu().numpy(), g["iou64_roll_per_sample"], rtol=1e-12)
    ls = S.LeastSquares(64, dev)            # fp32 internals in the reference itself (classes.py:319)
    l, gr = run(ls, imgs, lab, dev)
    check(l, gr, g["lsq64_loss"], g["lsq64_grad"], loss_rtol=1e-4, rtol=1e-3, atol=1e-5, what="lsq64")
    l, gr = run(ls, imgs, roll, dev)
    check(l, gr, g["lsq64_roll_loss"], g["lsq64_roll_grad"], loss_rtol=1e-4, rtol=1e-3, atol=1e-3, what="lsq64 rolled")


def test_fixture_main_and_visu(fixtures_golden, dev, S):
    g = fixtures_golden
    main = torch.tensor(g["main_params"]).to(dev)                      # classes.py:453-473: IoU of identical params
    assert S.IoUAccuracy(render_size=64, device=dev)(main, main).item() == 1.0
    # visu.py:142-165: fp64 leaf tensors, .backward(), read pred.grad
    l, gr = run(S.ExplicitLoss(render_size=32, device=dev), g["visu_true"], g["visu_pred"], dev)
    check(l, gr, g["visu_explicit32_loss"], g["visu_explicit32_grad"], what="visu explicit32 fp64")
    a = S.IoUAccuracy(render_size=128, device=dev, full=True)(torch.tensor(g["visu_true"]).to(dev),
                                                               torch.tensor(g["visu_pred"]).to(dev))
    assert abs(a.detach().cpu().item() - float(g["visu_iou128"])) < 1e-7


# ------------------------------------------------------------------ seeded random inputs frozen from the reference
@pytest.mark.parametrize("fname", random_golden_files())
def test_random_goldens(fname, dev, S):
    g = load_golden(fname)
    R = int(g["R"])
    for tag in ("far", "near"):
        pred = g[f"pred_{tag}"]
        for name, crit in (("implicit_t15_k260", S.ImplicitLoss(R, dev, 1.5, 260)), ("implicit_default", S.ImplicitLoss(R, dev))):
            l, gr = run(crit, g["img"], pred, dev)
            check(l, gr, g[f"{name}_{tag}_loss"], g[f"{name}_{tag}_grad"], what=f"{fname} {name} {tag}")
        l, gr = run(S.ExplicitLoss(R, dev), g["true"], pred, dev)
        check(l, gr, g[f"explicit_{tag}_loss"], g[f"explicit_{tag}_grad"], what=f"{fname} explicit {tag}")
        i, u = S.IoUAccuracy(R, dev).counts(torch.tensor(g["true"]).to(dev), torch.tensor(pred).to(dev))
        assert (i.cpu().numpy() == g[f"iou_{tag}_inter"]).all() and (u.cpu().numpy() == g[f"iou_{tag}_union"]).all()
        l, gr = run(S.LeastSquares(R, dev), g["img"], pred, dev)
        check(l, gr, g[f"lsq_{tag}_loss"], g[f"lsq_{tag}_grad"], loss_rtol=1e-4, rtol=2e-3, atol=1e-4, what=f"{fname} lsq {tag}")
    d = S.ImplicitLoss(R, dev, 1.5, 260).depth_projection(torch.tensor(g["true"]).to(dev))
    np.testing.assert_allclose(d.cpu().numpy(), g["depth_true"], atol=2e-5)
    if g["occupancy_true"].size:
        occ = S.ExplicitLoss(R, dev).occupancy(torch.tensor(g["true"][:2]).to(dev))
        np.testing.assert_allclose(occ.cpu().numpy(), g["occupancy_true"], atol=2e-5)


# ------------------------------------------------------------------ edge cases
def test_edge_cases(edge_golden, dev, S):
    g = edge_golden
    R = int(g["R"])
    img = torch.tensor(g["img"]).float()
    l, gr = run(S.ImplicitLoss(R, dev, 1.5, 260), img, g["pred"], dev)          # fp64 parameters
    check(l, gr, g["implicit_loss"], g["implicit_grad"], what="edge implicit")
    l, gr = run(S.ImplicitLoss(R, dev, 1.0, 20), img, g["pred"], dev)
    check(l, gr, g["implicit_soft_loss"], g["implicit_soft_grad"], what="edge implicit soft")
    l, gr = run(S.ExplicitLoss(R, dev), g["true"], g["pred"], dev)
    check(l, gr, g["explicit_loss"], g["explicit_grad"], what="edge explicit")
    # clamp sub-gradient: exactly 0 outside the range, live on the inclusive boundary
    assert gr[0, 0] == 0 and gr[0, 1] == 0 and gr[0, 2] != 0
    assert gr[1, 3] == 0 and gr[1, 4] != 0
    assert gr[2, 5] == 0 and gr[2, 6] == 0 and gr[2, 7] != 0
    l, gr = run(S.ExplicitLoss(R, dev), g["pred"], g["true"], dev)
    check(l, gr, g["explicit_swapped_loss"], g["explicit_swapped_grad"], what="edge explicit swapped")
    i, u = S.IoUAccuracy(R, dev).counts(torch.tensor(g["true"]).to(dev), torch.tensor(g["pred"]).to(dev))
    assert (i.cpu().numpy() == g["iou_inter"]).all() and (u.cpu().numpy() == g["iou_union"]).all()
    ex24 = S.ExplicitLoss(24, dev)                                      # arange(0, 1+1/24, 1/24) has R+2 entries
    assert ex24.xyz.shape == (3, 26, 26, 26) and int(g["explicit24_n"]) == 26
    l, gr = run(ex24, g["true24"], g["pred24"], dev)
    check(l, gr, g["explicit24_loss"], g["explicit24_grad"], what="explicit R=24")


# ------------------------------------------------------------------ live oracle on fresh seeded inputs
@pytest.mark.parametrize("seed,B,R", [(21, 16, 32), (22, 10, 64), (23, 16, 16)])
def test_against_oracle(seed, B, R, dev, S):
    true = O.random_params(B, seed)
    pred_far, pred_near = O.random_params(B, seed + 500), O.perturbed_params(true, seed)
    with torch.no_grad():
        img = O.ImplicitLoss(4 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
    for pred in (pred_far, pred_near):
        for args in ((1.5, 260), (1, 100)):
            p = pred.clone().requires_grad_(True)
            oc = O.ImplicitLoss(R, "cpu", *args)
            ref = oc(img, p); ref.backward()
            l, gr = run(S.ImplicitLoss(R, dev, *args), img, pred, dev)
            keep = unambiguous(oc, img, pred)
            assert keep.sum() >= B - 2
            check(l, gr, ref.item(), p.grad.double().numpy(), what=f"implicit{args} B={B} R={R}", keep=keep)
        p = pred.clone().requires_grad_(True)
        ref = O.ExplicitLoss(R, "cpu")(true, p); ref.backward()
        l, gr = run(S.ExplicitLoss(R, dev), true, pred, dev)
        check(l, gr, ref.item(), p.grad.double().numpy(), what=f"explicit B={B} R={R}")
        i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
        i2, u2 = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
        assert torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu())


@pytest.mark.parametrize("B,R", [(5, 12), (3, 20), (2, 33)])
def test_odd_sizes_against_oracle(B, R, dev, S):
    """Render sizes that are not multiples of 8 (x-fastest column layout, masked lanes, ragged last warp) and batch
    sizes that do not fill a block; every kernel once.  (compute-sanitizer is closed on this GPU pool, so out-of-range
    accesses are hunted with these ragged cases + comparison with the oracle.)"""
    true, pred = O.random_params(B, 61), O.random_params(B, 62)
    with torch.no_grad():
        img = O.ImplicitLoss(3 * R + 1, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)   # ragged resize ratio
    p = pred.clone().requires_grad_(True)
    oc = O.ImplicitLoss(R, "cpu", 1.0, 100)
    ref = oc(img, p); ref.backward()
    l, gr = run(S.ImplicitLoss(R, dev, 1.0, 100), img, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), what=f"implicit odd B={B} R={R}", keep=unambiguous(oc, img, pred))
    p = pred.clone().requires_grad_(True)
    ref = O.ExplicitLoss(R, "cpu")(true, p); ref.backward()
    l, gr = run(S.ExplicitLoss(R, dev), true, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), what=f"explicit odd B={B} R={R}")
    i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
    i2, u2 = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
    assert torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu())
    p = pred.clone().requires_grad_(True)
    ref = O.LeastSquares(R, "cpu")(img, p); ref.backward()
    l, gr = run(S.LeastSquares(R, dev), img, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), loss_rtol=1e-4, rtol=2e-3, atol=1e-4, what=f"lsq odd B={B} R={R}")
    occ = S.ExplicitLoss(R, dev).occupancy(pred.to(dev)).cpu()
    np.testing.assert_allclose(occ.numpy(), O.ExplicitLoss(R, "cpu").occupancy(pred).numpy(), atol=2e-5)
    f = S.IoUAccuracy(R, dev).ins_outs(pred.to(dev)).cpu().double()
    fo = O.IoUAccuracy(R, "cpu").ins_outs(pred)
    ok = torch.isfinite(fo) & (fo < 1e6)
    np.testing.assert_allclose(f[ok].numpy(), fo[ok].numpy(), rtol=2e-5)


# ------------------------------------------------------------------ full BASELINE sizes through size-independent properties
def test_full_size_properties(dev, S):
    B, R = 256, 64                                                      # BASELINE config 2
    true = O.random_params(B, 0).to(dev)
    pred = O.perturbed_params(O.random_params(B, 0), 5).to(dev)
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    target = crit.depth_projection(true).unsqueeze(1)                   # synthetic depth maps
    p1 = pred.clone().requires_grad_(True)
    l1 = crit(target, p1); l1.backward()
    p2 = pred.clone().requires_grad_(True)
    l2 = crit(target, p2); l2.backward()
    assert l1.item() == l2.item() and torch.equal(p1.grad, p2.grad)    # run-to-run bit reproducible
    assert torch.isfinite(p1.grad).all()
    # rendering the true params and comparing them with themselves gives (numerically) zero loss
    assert crit(target, true).item() < 1e-6
    # batch split invariance: the loss is the mean of per-sample losses, gradients scale with 1/B
    halves = []
    for sl in (slice(0, 128), slice(128, 256)):
        ph = pred[sl].clone().requires_grad_(True)
        lh = crit(target[sl], ph); lh.backward()
        halves.append((lh.item(), ph.grad))
    assert abs(0.5 * (halves[0][0] + halves[1][0]) - l1.item()) < 1e-12
    torch.testing.assert_close(torch.cat([halves[0][1], halves[1][1]]) * 0.5, p1.grad, rtol=1e-6, atol=1e-12)
    # a slice of the full batch against the oracle
    idx = [0, 31, 77, 100, 128, 160, 200, 222, 240, 255]
    po = pred[idx].cpu().clone().requires_grad_(True)
    oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
    ref = oc(target[idx].cpu(), po); ref.backward()
    ps = pred[idx].clone().requires_grad_(True)
    ls = crit(target[idx], ps); ls.backward()
    check(ls.item(), ps.grad.double().cpu().numpy(), ref.item(), po.grad.double().numpy(), what="slice of config 2",
          keep=unambiguous(oc, target[idx].cpu(), pred[idx].cpu()))
    # ExplicitLoss / IoU identities at full size
    ex = S.ExplicitLoss(R, dev)
    assert ex(true, true).item() == 0.0
    assert S.IoUAccuracy(R, dev)(true, true).item() == 1.0
    i, u = S.IoUAccuracy(R, dev).counts(true, pred)
    assert (i <= u).all() and (u > 0).all()
    pe = pred.clone().requires_grad_(True)
    le = ex(true, pe); le.backward()
    assert torch.isfinite(pe.grad).all() and le.item() > 0


# ------------------------------------------------------------------ API / autograd behaviour
def test_autograd_contract(dev, S):
    B, R = 4, 32
    true = O.random_params(B, 3).to(dev)
    img = S.ImplicitLoss(128, dev, 1.5, 260).depth_projection(true).unsqueeze(1)
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    # non-leaf pred built with torch.cat like train.py:89; gradient must reach the heads
    heads = [O.random_params(B, 4).to(dev)[:, s].clone().requires_grad_(True) for s in (slice(0, 3), slice(3, 5), slice(5, 8), slice(8, 12))]
    pred = torch.cat(heads, dim=1)
    loss = crit(img, pred)
    assert loss.dim() == 0 and loss.dtype == torch.float64 and loss.device.type == "cuda"
    (3.0 * loss).backward()
    g3 = torch.cat([h.grad for h in heads], dim=1)
    p = pred.detach().clone().requires_grad_(True)
    crit(img, p).backward()
    torch.testing.assert_close(g3, 3.0 * p.grad, rtol=1e-6, atol=0)
    with torch.no_grad():                                               # train.py:135-146 validation path
        lv = crit(img, pred)
        assert not lv.requires_grad and abs(lv.item() - loss.item()) < 1e-7
    assert not crit(img, pred.detach()).requires_grad
    # fp64 leaf (visu.py:142-153): grad comes back in fp64
    p64 = pred.detach().double().requires_grad_(True)
    S.ExplicitLoss(R, dev)(true.double(), p64).backward()
    assert p64.grad.dtype == torch.float64
    # public attributes the reference sets (classes.py:114-127, 208-222, 381-392)
    assert crit.render_size == R and crit.tau == 1.5 and crit.sigmoid_sharpness == 260 and crit.reduce is True
    assert crit.xyz.shape == (3, R, R, R) and crit.xyz.dtype == torch.float64 and crit.xyz.min().item() == 1e-4
    assert S.ExplicitLoss(R, dev).xyz.shape == (3, R + 1, R + 1, R + 1)
    assert S.IoUAccuracy(R, dev).xyz.min().item() == 0.0
    pp = S.ExplicitLoss.preprocess_sq(torch.tensor([[2., .01, .5, 0., 2., -1., .5, 2., 1., 2., 3., 4.]]))
    np.testing.assert_allclose(pp.numpy(), [[1., .05, .5, .1, 1., 0., .5, 1., 1., 2., 3., 4.]], rtol=1e-7)
    with pytest.raises(RuntimeError):
        S.ImplicitLoss(R, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        crit(img.cpu(), pred.detach().cpu())


def test_quaternion_helpers(dev, S):
    q = torch.tensor(O.randquat(np.random.RandomState(0)), device=dev)
    np.testing.assert_allclose(S.quaternion.mat_from_quaternion(q)[0].cpu().numpy(), O.mat_from_quaternion(q.cpu()).numpy(), atol=1e-15)
    assert torch.equal(S.quaternion.conjugate(q).cpu(), O.conjugate(q.cpu()))


# ------------------------------------------------------------------ host-buffer C ABI (include/sqloss.h)
def test_host_entry_points(dev, S):
    from sq_recovery_b200.functional import HostContext
    B, R = 8, 32
    true, pred = O.random_params(B, 31), O.random_params(B, 32)
    img = S.ImplicitLoss(128, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).cpu()
    ctx = HostContext(0)
    l, g = ctx.implicit_loss(pred.numpy(), img.numpy(), R, 1.5, 260.0)
    l2, g2 = run(S.ImplicitLoss(R, dev, 1.5, 260), img, pred, dev)
    assert l == l2 and np.array_equal(g.astype(np.float64), g2)
    l, g = ctx.explicit_loss(true.numpy(), pred.numpy(), R)
    l2, g2 = run(S.ExplicitLoss(R, dev), true, pred, dev)
    assert l == l2 and np.array_equal(g.astype(np.float64), g2)
    i, u = ctx.iou_counts(true.numpy(), pred.numpy(), R)
    i2, u2 = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
    assert np.array_equal(i, i2.cpu().numpy()) and np.array_equal(u, u2.cpu().numpy())
    # two calls in flight on the context's two slots (submit / wait), fp32 and 8-bit images, pinned and pageable
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    u8 = (img * 255.0).round().clamp(0, 255).to(torch.uint8)
    as_f32 = u8.float() * np.float32(1.0 / 255.0)          # what the device makes of the 8-bit pixels
    want_f = run(crit, img, pred, dev)
    want_u = run(crit, as_f32, pred, dev)
    pin_u8 = torch.empty(u8.shape, dtype=torch.uint8).pin_memory(); pin_u8.copy_(u8)
    pin_f = torch.empty(img.shape, dtype=torch.float32).pin_memory(); pin_f.copy_(img)
    for a, b in ((pin_f.numpy(), pin_u8.numpy()), (img.numpy(), u8.numpy())):
        ctx.submit_implicit(0, pred.numpy(), a, R, 1.5, 260.0)
        ctx.submit_implicit(1, pred.numpy(), b, R, 1.5, 260.0)
        with pytest.raises(RuntimeError):                   # a slot holds one call at a time
            ctx.submit_implicit(1, pred.numpy(), b, R, 1.5, 260.0)
        l1, g1 = ctx.result(1)
        l0, g0 = ctx.result(0)
        assert l0 == want_f[0] and np.array_equal(g0.astype(np.float64), want_f[1])
        assert l1 == want_u[0] and np.array_equal(g1.astype(np.float64), want_u[1])
    ctx.submit_implicit(0, pred.numpy(), pin_u8.numpy(), R, 1.5, 260.0, want_grad=False)
    l0, g0 = ctx.result(0)
    assert g0 is None and abs(l0 - want_u[0]) <= 1e-6 * abs(want_u[0])      # the forward-only kernel (packed fp32 pairs)
    ctx.close()


def test_host_image_transfer_modes(dev, S):
    """sq_implicit_loss_host_submit picks how the pixels cross PCIe from the image shape (copy-engine rows / rows read in
    place / single pixels / pageable copy): every mode must give the bits of the class path on the same pixels, also when
    shapes, batch sizes and slots alternate (the resize tables are kept on the device while the shape stays the same) and
    when another host call reuses the slot's arena in between."""
    from sq_recovery_b200.functional import HostContext
    ctx = HostContext(0)
    rs = np.random.RandomState(3)
    seq = [(8, 16, 64, 64), (8, 16, 64, 64), (8, 16, 40, 64), (8, 16, 40, 56), (3, 16, 64, 64), (8, 16, 64, 64),
           (5, 32, 128, 96), (8, 16, 64, 64), (8, 24, 48, 96)]
    for n, (B, R, H, W) in enumerate(seq):
        pred = O.random_params(B, 40 + n)
        for dtype in (torch.uint8, torch.float32):
            raw = torch.from_numpy(rs.randint(0, 256, size=(B, 1, H, W)).astype(np.uint8))
            host = raw if dtype == torch.uint8 else raw.float() / 255.0
            as_f32 = raw.float() * np.float32(1.0 / 255.0) if dtype == torch.uint8 else host
            want = run(S.ImplicitLoss(R, dev, 1.5, 260), as_f32, pred, dev)
            pin = torch.empty(host.shape, dtype=dtype).pin_memory(); pin.copy_(host)
            slot = n & 1
            ctx.submit_implicit(slot, pred.numpy(), pin.numpy(), R, 1.5, 260.0)
            ctx.submit_implicit(1 - slot, pred.numpy(), host.numpy(), R, 1.5, 260.0)          # pageable
            for sl in (slot, 1 - slot):
                l, g = ctx.result(sl)
                assert l == want[0] and np.array_equal(g.astype(np.float64), want[1]), (n, dtype, sl)
        if n == 4:
            ctx.explicit_loss(O.random_params(4, 1).numpy(), O.random_params(4, 2).numpy(), 12)
    ctx.close()


# ------------------------------------------------------------------ work queues, scratch contract, culling corner cases
def _scratch_control_words(S, dev):
    from sq_recovery_b200 import functional as Fn
    torch.cuda.synchronize()
    return [buf[:256].view(torch.int32).cpu() for key, buf in Fn._scratch.items() if key[0] == dev.index]


def test_scratch_is_reusable_and_left_clean(dev, S):
    """include/sqloss.h scratch contract: one buffer serves any sequence of calls (different losses, batch and grid
    sizes) and every call leaves the queue counters of the control block zero again."""
    rs = np.random.RandomState(5)
    cases = [(7, 16), (33, 32), (2, 64), (64, 24), (1, 20)]
    first = {}
    for rep in range(2):
        for B, R in cases:
            true = O.random_params(B, 100 + B).to(dev)
            pred = O.perturbed_params(O.random_params(B, 100 + B), 3).to(dev)
            img = S.ImplicitLoss(64, dev, 1.5, 260).depth_projection(true).unsqueeze(1)
            out = []
            p = pred.clone().requires_grad_(True)
            l = S.ImplicitLoss(R, dev, 1.5, 260)(img, p); l.backward(); out += [l.item(), p.grad.clone()]
            p = pred.clone().requires_grad_(True)
            l = S.ExplicitLoss(R, dev)(true, p); l.backward(); out += [l.item(), p.grad.clone()]
            out += [t.clone() for t in S.IoUAccuracy(R, dev).counts(true, pred)]
            p = pred.clone().requires_grad_(True)
            l = S.LeastSquares(R, dev)(img, p); l.backward(); out += [l.item(), p.grad.clone()]
            for w in _scratch_control_words(S, dev):
                assert int(w[2]) == 0 and not w[4:36].any(), "queue counters not left clean"
            if rep == 0:
                first[(B, R)] = out
            else:                                                       # same inputs, dirty-then-cleaned scratch: same bits
                for a, b in zip(first[(B, R)], out):
                    assert (a == b) if isinstance(a, float) else torch.equal(a, b)


def test_objects_off_the_grid(dev, S):
    """Samples whose superquadric lies (partly or entirely) outside the unit cube: the plan kernel proves most or all
    work items empty; loss and gradients must still match the oracle (all-empty: loss = mean |target|, zero gradient)."""
    B, R = 6, 32
    pred = O.random_params(B, 11)
    pred[0, 5:8] = torch.tensor([3.0, 0.5, 0.5])        # clamped to t = (1, .5, .5): half outside
    pred[1, 5:8] = torch.tensor([-2.0, -2.0, -2.0])     # clamped to the corner (0, 0, 0)
    pred[2, 0:3] = 0.05                                   # smallest allowed size
    pred[3, 5:8] = torch.tensor([0.5, 0.5, 5.0])        # at the camera plane
    true = O.random_params(B, 12)
    img = S.ImplicitLoss(64, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1)
    oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
    po = pred.clone().requires_grad_(True)
    ref = oc(img.cpu(), po); ref.backward()
    l, g = run(S.ImplicitLoss(R, dev, 1.5, 260), img, pred, dev)
    check(l, g, ref.item(), po.grad.double().numpy(), what="off-grid implicit", keep=unambiguous(oc, img.cpu(), pred))
    pe = pred.clone().requires_grad_(True)
    refe = O.ExplicitLoss(R, "cpu")(true, pe); refe.backward()
    l, g = run(S.ExplicitLoss(R, dev), true, pred, dev)
    check(l, g, refe.item(), pe.grad.double().numpy(), what="off-grid explicit")
    i, u = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
    ri, ru = O.IoUAccuracy(R, "cpu").counts(true, pred)
    assert torch.equal(i.cpu(), ri) and torch.equal(u.cpu(), ru)
    # the loss assembled from the plan kernel's sum |target| and the column kernel's per-column differences equals the
    # mean absolute difference to the kernel's own render (most columns of a tiny object are never walked)
    tiny = pred[:2].clone().to(dev)
    tiny[:, 0:3] = 0.05
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    blank = torch.full((2, 1, 64, 64), 0.25, device=dev)
    d = crit.depth_projection(tiny)
    assert d[:, :2, :].abs().max().item() == 0.0          # image border: proven empty, exactly 0
    l = crit(blank, tiny)
    expect = (0.25 - d.double()).abs().mean().item()
    assert abs(l.item() - expect) <= 1e-6 * expect


def test_render_256_and_index_order_path(dev, S):
    """depth_projection at the data-generation size (R = 256) against the oracle, and a grid so large that a
    sample has more work items than the plan kernel classifies (index-order hand-out, no queues), checked against a
    render assembled from the full occupancy field (sq_field) with torch ops."""
    p = O.random_params(2, 31)
    d = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(p.to(dev))
    with torch.no_grad():
        ref = O.ImplicitLoss(256, "cpu", 1.5, 260).depth_projection(p)
    err = (d.double().cpu() - ref).abs()                 # fp32 kernel vs fp64: grazing pixels see k-amplified rounding
    assert err.max().item() < 3e-5 and err.mean().item() < 3e-7, (err.max().item(), err.mean().item())
    R = 520                                              # 520^2 / 32 = 8450 items per sample > 8192
    crit = S.ImplicitLoss(R, dev, 1.0, 100)
    q = O.random_params(1, 32).to(dev)
    d = crit.depth_projection(q)
    from sq_recovery_b200 import functional as Fn
    occ = Fn.field(q, R, 1.0 / (R - 1), 1e-4, 1, 100.0)[0].double()             # (x, y, z) occupancy
    cs = torch.cumsum(torch.flip(occ, dims=[2]), dim=2)
    depth = 1.0 - torch.exp(-1.0 * cs).sum(dim=2) / R
    img = depth.permute(1, 0).flip(0)                    # classes.py:279
    assert (d[0].double() - img).abs().max().item() < 1e-5
    for w in _scratch_control_words(S, dev):
        assert int(w[2]) == 0 and not w[4:36].any()


# ------------------------------------------------------------------ harnesses
def test_batched_descent_follows_the_oracle(dev, S):
    """harness/optimize.py (visu.py:120-186 for many SQs at once): a few descent steps with the CUDA ExplicitLoss land
    where the same steps with the fp64 oracle land."""
    from harness import optimize
    B, R, steps = 5, 32, 6
    true = O.random_params(B, 41)
    start = O.perturbed_params(O.random_params(B, 41), 4, sigma=0.05)
    ref = start.clone().double()
    optimize.descend(O.ExplicitLoss(R, "cpu"), true.double(), ref, steps)
    got = start.clone().to(dev)
    rec = []
    optimize.descend(S.ExplicitLoss(R, dev), true.to(dev), got, steps, record=rec)
    assert rec[-1].item() < rec[0].item()
    np.testing.assert_allclose(got.cpu().double().numpy(), ref.numpy(), rtol=0, atol=2e-6)


def test_dataset_generator(dev, S, tmp_path):
    from harness import make_dataset
    p = O.random_params(3, 51)
    imgs = make_dataset.render(p, dev, size=256, chunk=2)
    assert imgs.shape == (3, 1, 256, 256) and imgs.dtype == torch.float32
    assert 0.0 <= imgs.min().item() and imgs.max().item() < 1.0 and 0.02 < (imgs > 0).float().mean().item() < 0.6
    assert torch.equal(imgs[:, 0], S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(p.to(dev)))
    with torch.no_grad():                                               # one image against the oracle's render of the same label
        ref = O.ImplicitLoss(256, "cpu", 1.5, 260).depth_projection(p[2:3])
    np.testing.assert_allclose(imgs[2, 0].cpu().double().numpy(), ref[0].numpy(), rtol=0, atol=3e-5)
    # and the label file the reference's parse_csv reads back gives the parameters that were rendered
    rows = make_dataset.parse_rows(make_dataset.label_rows(p.numpy(), [f"synth/{i:06d}.bmp" for i in range(3)]))
    np.testing.assert_allclose(np.stack(rows), p.numpy(), rtol=0, atol=1e-6)


def test_fused_heads_match_torch_heads(dev, S):
    """ImplicitLoss.from_heads(raw) == ImplicitLoss(heads(raw)) with the heads of torch/models.py.
    At sharpness 260 the gradient is discontinuous in the parameters (sign(depth - target) per pixel), so the two paths
    are compared on bit-identical parameters: the heads are evaluated here the way the kernel does (fp64, rounded to
    fp32 like the reference's fp32 heads), the plain loss gives d loss / d params, and the head Jacobians are applied
    in fp64.  The loss is also compared with torch's own fp32 heads."""
    B, R = 12, 32
    true = O.random_params(B, 61).to(dev)
    img = S.ImplicitLoss(128, dev, 1.5, 260).depth_projection(true).unsqueeze(1)
    g = torch.Generator().manual_seed(3)
    p = O.perturbed_params(O.random_params(B, 61), 8).clamp(1e-3, 1 - 1e-3)
    raw = torch.cat([torch.logit(p[:, :8]), p[:, 8:] * (0.5 + torch.rand(B, 1, generator=g))], dim=1)
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    r1 = raw.clone().to(dev).requires_grad_(True)
    l1 = crit.from_heads(img, r1); l1.backward()
    # the kernel's heads, restated
    r64 = raw.double()
    hp = torch.sigmoid(r64[:, :8]).float().double()
    rn = r64[:, 8:].norm(dim=1, keepdim=True)
    q = (r64[:, 8:] / rn).float().double()
    pe = torch.cat([hp, q], dim=1).float().to(dev).requires_grad_(True)
    l2 = crit(img, pe); l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-9 * abs(l2.item())
    gp = pe.grad.double().cpu()
    expect = torch.cat([gp[:, :8] * hp * (1 - hp),
                        (gp[:, 8:] - q * (q * gp[:, 8:]).sum(dim=1, keepdim=True)) / rn], dim=1)
    torch.testing.assert_close(r1.grad.double().cpu(), expect, rtol=2e-6, atol=1e-10)
    # and against torch's own fp32 heads (models.py:28,52,75,98 + train.py:89): same loss to fp32 head precision
    r2 = raw.clone().to(dev)
    q2 = r2[:, 8:]
    pred = torch.cat([torch.sigmoid(r2[:, 0:3]), torch.sigmoid(r2[:, 3:5]), torch.sigmoid(r2[:, 5:8]),
                      q2 / torch.norm(q2, 2, -1, keepdim=True)], dim=1)
    assert abs(l1.item() - crit(img, pred).item()) <= 1e-5 * abs(l1.item())


def test_cuda_graph_capture_and_replay(dev, S):
    """a step (loss + backward) captured in a CUDA graph replays to the same bits as the eager
    call, over many replays (the queue counters in the scratch control block are restored by every run)."""
    B, R = 24, 32
    true = O.random_params(B, 71).to(dev)
    pred = O.perturbed_params(O.random_params(B, 71), 6).to(dev)
    img = S.ImplicitLoss(128, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    p0 = pred.clone().requires_grad_(True)
    l0 = crit(img, p0); l0.backward()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            pw = pred.clone().requires_grad_(True)
            crit(img, pw).backward()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    # a capture on a stream that has never run the call eagerly is refused (its workspace would be born inside the graph)
    cold = torch.cuda.Stream(dev)
    with pytest.raises(RuntimeError, match="capture"):
        with torch.cuda.graph(torch.cuda.CUDAGraph(), stream=cold):
            crit(img, pred.clone().requires_grad_(True))
    torch.cuda.synchronize()
    # two graphs captured on the warmed stream; only the SECOND is replayed at first (round-1 advisor finding: the
    # workspace's one-time initialisation must not live inside the first captured graph)
    graphs = []
    for _ in range(2):
        g = torch.cuda.CUDAGraph()
        p = pred.clone().requires_grad_(True)
        with torch.cuda.graph(g, stream=side):
            lg = crit(img, p)
            lg.backward()
        graphs.append((g, lg, p))
    g, lg, p = graphs[1]
    for _ in range(25):
        g.replay()
    torch.cuda.synchronize()
    assert lg.item() == l0.item() and torch.equal(p.grad, p0.grad)
    g, lg, p = graphs[0]
    g.replay()
    torch.cuda.synchronize()
    assert lg.item() == l0.item() and torch.equal(p.grad, p0.grad)
    # prepare_stream() is the explicit way to make a stream capturable
    from sq_recovery_b200 import functional as Fn
    s3 = torch.cuda.Stream(dev)
    Fn.prepare_stream(dev, B, R, s3)
    g3 = torch.cuda.CUDAGraph()
    p3 = pred.clone().requires_grad_(True)
    with torch.cuda.graph(g3, stream=s3):
        l3 = crit(img, p3)
        l3.backward()
    g3.replay()
    torch.cuda.synchronize()
    assert l3.item() == l0.item() and torch.equal(p3.grad, p0.grad)
    # the eager path on the default stream still works next to the captured one
    p1 = pred.clone().requires_grad_(True)
    l1 = crit(img, p1); l1.backward()
    assert l1.item() == l0.item() and torch.equal(p1.grad, p0.grad)


# ------------------------------------------------------------------ round-2 additions
def test_least_squares_energy_function(fixtures_golden, dev, S):
    """LeastSquares.energy_function(batch_points, params) (classes.py:318-356) on explicit, ragged point lists: per-sample
    energies and their gradients against the oracle (fp32 arithmetic in the reference itself, classes.py:319)."""
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    lab = torch.tensor(g["labels"])
    ols = O.LeastSquares(64, "cpu")
    pts = ols.points(imgs)                                              # ragged (3, m_i) lists from the depth images
    pts[3] = pts[3][:, :1]                                              # a single point
    pts[4] = pts[4][:, :0]                                              # an empty list
    for params in (lab, torch.roll(lab, 1, 0)):
        po = params.clone().requires_grad_(True)
        ref = ols.energy_function(pts, po)
        w = torch.linspace(0.5, 1.5, len(pts))
        (ref * w).sum().backward()
        pg = params.clone().to(dev).requires_grad_(True)
        got = S.LeastSquares(64, dev).energy_function([p.to(dev) for p in pts], pg)
        assert got.shape == (len(pts),) and got.dtype == torch.float32
        (got * w.to(dev)).sum().backward()
        np.testing.assert_allclose(got.detach().cpu().numpy(), ref.detach().numpy(), rtol=2e-4, atol=1e-7)
        rg, gg = po.grad.double().numpy(), pg.grad.double().cpu().numpy()
        assert (np.abs(gg - rg) <= 1e-3 * np.abs(rg).max(axis=1, keepdims=True) + 2e-3 * np.abs(rg) + 1e-5).all()
    # the image path and the point-list path are the same function
    po = lab.clone().to(dev)
    a = S.LeastSquares(64, dev)(imgs.to(dev), po).item()
    b = S.LeastSquares(64, dev).energy_function([p.to(dev) for p in ols.points(imgs)], po).mean().item()
    assert abs(a - b) <= 1e-5 * abs(a)


def test_forward_only_methods_and_no_grad(dev, S):
    """The grid-returning conveniences refuse tensors that ask for a gradient (they are forward-only here); under
    torch.no_grad() a prediction that requires grad runs the forward-only kernels and yields a history-free loss."""
    B, R = 4, 16
    true, pred = O.random_params(B, 5).to(dev), O.random_params(B, 6).to(dev)
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    img = crit.depth_projection(true).unsqueeze(1)
    pr = pred.clone().requires_grad_(True)
    for fn in (crit.depth_projection, S.ExplicitLoss(R, dev).occupancy, S.IoUAccuracy(R, dev).ins_outs):
        with pytest.raises(RuntimeError, match="forward-only"):
            fn(pr)
        with torch.no_grad():
            assert not fn(pr).requires_grad
        assert torch.equal(fn(pr.detach()), fn(pred))
    ref = crit(img, pred).item()
    with torch.no_grad():
        for c, t in ((crit, img), (S.ExplicitLoss(R, dev), true), (S.LeastSquares(R, dev), img)):
            l = c(t, pr)
            assert not l.requires_grad and l.grad_fn is None
        assert crit(img, pr).item() == ref
    with pytest.raises(ValueError):
        S.LeastSquares(R, dev)(img[:2], pred)


def test_config1_exact_shape(dev, S):
    """BASELINE config 1 exactly: batch 32, render size 32, both grid losses, test_random.py-style random SQs."""
    B, R = 32, 32
    true, pred = O.random_params(B, 0), O.random_params(B, 1)
    with torch.no_grad():
        img = O.ImplicitLoss(4 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
    oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
    p = pred.clone().requires_grad_(True)
    ref = oc(img, p); ref.backward()
    l, gr = run(S.ImplicitLoss(R, dev, 1.5, 260), img, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), what="config 1 implicit", keep=unambiguous(oc, img, pred))
    p = pred.clone().requires_grad_(True)
    ref = O.ExplicitLoss(R, "cpu")(true, p); ref.backward()
    l, gr = run(S.ExplicitLoss(R, dev), true, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), what="config 1 explicit")
    i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
    i2, u2 = S.IoUAccuracy(R, dev).counts(true.to(dev), pred.to(dev))
    assert torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu())


def test_config5_explicit128_and_iou128(dev, S):
    """BASELINE config 5's sizes: ExplicitLoss(128) forward (129^3 grid, visu.py:71) and IoUAccuracy(128) against the
    oracle on a few pairs."""
    B = 3
    true, pred = O.random_params(B, 90), O.random_params(B, 91)
    with torch.no_grad():
        ref = O.ExplicitLoss(128, "cpu").per_sample(true, pred)
        ex = S.ExplicitLoss(128, dev)
        assert ex._n == 129
        for b in range(B):
            got = ex(true[b:b + 1].to(dev), pred[b:b + 1].to(dev)).item()
            assert abs(got - ref[b].item()) <= LOSS_RTOL * abs(ref[b].item())
        assert abs(ex(true.to(dev), pred.to(dev)).item() - ref.mean().item()) <= LOSS_RTOL * ref.mean().item()
    p = pred.clone().requires_grad_(True)
    refl = O.ExplicitLoss(128, "cpu")(true[:1], p[:1]); refl.backward()
    l, gr = run(S.ExplicitLoss(128, dev), true[:1], pred[:1], dev)
    check(l, gr, refl.item(), p.grad[:1].double().numpy(), what="explicit128 fwd+bwd")
    i, u = O.IoUAccuracy(128, "cpu").counts(true, pred)
    i2, u2 = S.IoUAccuracy(128, dev).counts(true.to(dev), pred.to(dev))
    assert torch.equal(i, i2.cpu()) and torch.equal(u, u2.cpu())


def test_config2_whole_batch_against_oracle(dev, S):
    """The real BASELINE config 2 call -- batch 256, 64^3, tau 1.5, sharpness 260 -- against the fp64 oracle on EVERY
    sample (the oracle runs in chunks of 32 samples to bound its memory)."""
    B, R = 256, 64
    torch.set_num_threads(os.cpu_count() or 1)
    true = O.random_params(B, 0)
    pred = O.perturbed_params(true, 5)
    crit = S.ImplicitLoss(R, dev, 1.5, 260)
    img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).contiguous()
    l, gr = run(crit, img, pred, dev)
    oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
    img_c = img.cpu()
    ref_loss, ref_grad, keep = 0.0, [], []
    for c in range(0, B, 32):
        p = pred[c:c + 32].clone().requires_grad_(True)
        part = oc(img_c[c:c + 32], p) * (32 / B)
        part.backward()
        ref_loss += part.item()
        ref_grad.append(p.grad.double().numpy())
        keep.append(unambiguous(oc, img_c[c:c + 32], pred[c:c + 32]))
    keep = np.concatenate(keep)
    assert keep.sum() >= B - 8
    check(l, gr, ref_loss, np.concatenate(ref_grad), what="config 2 whole batch", keep=keep)


def test_against_the_live_reference(dev, S):
    """The UNMODIFIED reference classes (staged under oracle/_ref by oracle/build_ref.py; /root/reference itself in the dev
    container) run on the CPU next to the CUDA path on the same inputs: all four classes, R = 16 and 32."""
    from oracle import ref_import
    if not (ref_import.available() or ref_import.staged()):
        pytest.skip("reference not staged under oracle/_ref (python -m oracle.build_ref in the dev container)")
    rc, rq = ref_import.load()
    cpu = torch.device("cpu")
    for R, B, seed in ((16, 6, 201), (32, 4, 202)):
        true = O.random_params(B, seed)
        for pred in (O.random_params(B, seed + 50), O.perturbed_params(true, seed)):
            with torch.no_grad():
                img = rc.ImplicitLoss(2 * R, cpu, 1.5, 260).depth_projection(true).float().unsqueeze(1)
            oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
            for args in ((1.5, 260), (1, 100)):
                p = pred.clone().requires_grad_(True)
                ref = rc.ImplicitLoss(R, cpu, *args)(img, p); ref.backward()
                l, gr = run(S.ImplicitLoss(R, dev, *args), img, pred, dev)
                check(l, gr, ref.item(), p.grad.double().numpy(), what=f"reference implicit{args} R={R}",
                      keep=unambiguous(O.ImplicitLoss(R, "cpu", *args), img, pred))
            p = pred.clone().requires_grad_(True)
            ref = rc.ExplicitLoss(R, cpu)(true, p); ref.backward()
            l, gr = run(S.ExplicitLoss(R, dev), true, pred, dev)
            check(l, gr, ref.item(), p.grad.double().numpy(), what=f"reference explicit R={R}")
            ref = rc.IoUAccuracy(R, cpu)(true, pred).item()
            assert abs(S.IoUAccuracy(R, dev)(true.to(dev), pred.to(dev)).item() - ref) < 1e-7
            per = rc.IoUAccuracy(R, cpu, reduce=False)(true, pred)
            np.testing.assert_allclose(S.IoUAccuracy(R, dev, reduce=False)(true.to(dev), pred.to(dev)).cpu().numpy(),
                                       per.numpy(), rtol=1e-12)
            p = pred.clone().requires_grad_(True)
            ref = rc.LeastSquares(R, cpu)(img, p); ref.backward()
            l, gr = run(S.LeastSquares(R, dev), img, pred, dev)
            rg = p.grad.double().numpy()            # the reference itself is fp32 here (classes.py:319): row-relative tolerance
            assert abs(l - ref.item()) <= 1e-4 * abs(ref.item())
            assert (np.abs(gr - rg) <= 1e-3 * np.abs(rg).max(axis=1, keepdims=True) + 2e-3 * np.abs(rg) + 1e-5).all()
    q = torch.tensor(O.randquat(np.random.RandomState(1)))
    np.testing.assert_allclose(S.quaternion.mat_from_quaternion(q.to(dev)).cpu().numpy(), rq.mat_from_quaternion(q).numpy(), atol=1e-15)


def test_parity_sweep_every_sample(dev, S):
    """Every sample of the frozen parity sweep (tests/golden/parity_sweep_refs.npz: 36 seeds x 2 prediction styles at
    R = 16 / 32 / 64, tau 1.5 / k 260 and the defaults; 4032 samples) meets the gradient tolerance entry by entry --
    the criterion round 1 could only meet statistically at k = 260."""
    refs = load_golden("parity_sweep_refs.npz")
    for R in (16, 32, 64):
        preds, imgs = refs[f"R{R}_pred"], refs[f"R{R}_img"]
        for (tau, k) in ((1.5, 260.0), (1.0, 100.0)):
            tag = f"R{R}_t{tau:g}_k{k:g}"
            crit = S.ImplicitLoss(R, dev, tau, k)
            for c in range(preds.shape[0]):
                l, gr = run(crit, imgs[c], preds[c], dev)
                check(l, gr, refs[tag + "_loss"][c], refs[tag + "_grad"][c], what=f"{tag} call {c}", keep=refs[tag + "_keep"][c])


def test_train_harness(dev, S, tmp_path):
    """harness/train_step.py (the structure of torch/train.py:72-175): a few epochs on a tiny synthetic set -- finite,
    the training loss goes down, validation loss / IoU are recorded, the best-val checkpoint is written in the reference's
    dict format and resuming continues from it; and the fused-heads loss equals the unfused one on the same network."""
    from harness import train_step
    from harness.model import SQRegressor
    ck = str(tmp_path / "model.pt")
    hist = train_step.main(["--batch", "16", "--epochs", "3", "--steps-per-epoch", "4", "--val-batches", "1", "--lr", "1e-3",
                            "--checkpoint", ck])
    assert len(hist["loss"]) == 3 and all(np.isfinite(hist["loss"])) and all(np.isfinite(hist["val_loss"]))
    assert hist["loss"][-1] < hist["loss"][0]
    assert all(0.0 <= a <= 1.0 for e in hist["val_acc"] for a in e)
    saved = torch.load(ck, weights_only=False)
    assert set(saved) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"}
    assert set(saved["loss"]) == {"loss", "val_loss", "val_acc"}
    best = int(np.argmin(hist["val_loss"]))
    assert saved["epoch"] == best and len(saved["loss"]["loss"]) == best + 1
    more = train_step.main(["--batch", "16", "--epochs", "1", "--steps-per-epoch", "4", "--val-batches", "1", "--checkpoint", ck,
                            "--resume"])
    assert len(more["loss"]) == best + 2                                 # history continues after the restored epoch
    # heads fused into the loss kernels == sigmoid / normalise / cat in torch, on one network and one batch
    torch.manual_seed(1)
    net = SQRegressor().to(dev)
    true = O.random_params(8, 5).to(dev)
    crit = S.ImplicitLoss(64, dev, 1.5, 260)
    images = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
    l1 = crit(images, net(images)); l1.backward()
    g1 = net.trunk.fc[2].weight.grad.clone()
    net.zero_grad()
    l2 = crit.from_heads(images, net(images, raw=True)); l2.backward()
    g2 = net.trunk.fc[2].weight.grad
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l1.item())
    assert (g1 - g2).norm().item() <= 2e-2 * g1.norm().item()           # fp32 heads vs fp64-in-kernel heads at k = 260


def test_objects_that_fill_the_grid(dev, S):
    """bench.py's `dense` workload: sizes a ~ U(0.5, 1) -- every pixel is covered, columns graze faces for more planes than a
    lane's backward queue holds, and with a prediction close to the target the loss (~3e-3) is a small difference of depths
    near 1.  Round 2 found the loss 2e-5 off here (an fp32 sum of 32 nearly equal per-column terms rounds one way); the
    partial rows now carry the terms relative to one of them."""
    from sq_recovery_b200 import inputs
    B, R = 4, 64
    for seed in (302, 304):
        true = inputs.random_params(B, seed, size_range=inputs.DENSE_SIZE_RANGE)
        for pred in (inputs.random_params(B, seed + 1000, size_range=inputs.DENSE_SIZE_RANGE), inputs.perturbed_params(true, seed)):
            with torch.no_grad():
                img = O.ImplicitLoss(2 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
            oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
            p = pred.clone().requires_grad_(True)
            ref = oc(img, p); ref.backward()
            l, gr = run(S.ImplicitLoss(R, dev, 1.5, 260), img, pred, dev)
            check(l, gr, ref.item(), p.grad.double().numpy(), what=f"dense seed {seed}", keep=unambiguous(oc, img, pred))
            with torch.no_grad():                                       # the forward-only kernel shares the bookkeeping
                lf = S.ImplicitLoss(R, dev, 1.5, 260)(img.to(dev), pred.to(dev)).item()
            assert abs(lf - ref.item()) <= LOSS_RTOL * abs(ref.item())


def test_soft_sigmoid_on_a_fine_grid(dev, S):
    """sigmoid_sharpness 20 on a 96^3 grid: the band of points between the backward's weight cut and the culling bound is tens of
    planes thick instead of one or two, and what a fixed 2^-24 cut dropped there added up to 1.1x the gradient tolerance on a
    small entry (tests/tools/parity_fuzz.py --seed 8, case 111).  The cut now widens with the band (implicit_active_bits)."""
    from sq_recovery_b200 import inputs
    R, tau, k = 96, 1.5, 20.0
    true = inputs.random_params(2, 5111, size_range=(0.3, 0.4))
    pred = inputs.perturbed_params(true, 5111, sigma=0.03)
    with torch.no_grad():
        img = O.ImplicitLoss(3 * R + 1, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
    oc = O.ImplicitLoss(R, "cpu", tau, k)
    p = pred.clone().requires_grad_(True)
    ref = oc(img, p); ref.backward()
    l, gr = run(S.ImplicitLoss(R, dev, tau, k), img, pred, dev)
    check(l, gr, ref.item(), p.grad.double().numpy(), rtol=0.5 * GRAD_RTOL, atol=0.5 * GRAD_ATOL, what="k = 20 on 96^3",
          keep=unambiguous(oc, img, pred))


def test_grazing_columns(dev, S):
    """One object that fills a 24^3 grid at tau = 3 (tests/golden/edge_grazing.npz, frozen from the reference): its columns graze
    the surface for most of their planes -- dozens of gradient-carrying points per column.  The parity fuzz's one outlier (4.6x
    the tolerance) while those points sat in 15-deep per-lane queues; the warp-shared pool takes them all."""
    g = load_golden("edge_grazing.npz")
    img = torch.tensor(g["img"]).float()
    l, gr = run(S.ImplicitLoss(int(g["R"]), dev, float(g["tau"]), float(g["k"])), img, g["pred"], dev)
    check(l, gr, g["implicit_loss"], g["implicit_grad"], what="grazing columns")


def test_zero_planes(dev, S):
    """Axis-aligned rotations with t_z exactly on a grid plane (e.g. a position clamped to 1): the reference's exact-zero
    fix-up fires on a whole z plane -- the walk direction of the kernels (tests/golden/edge_zero_planes.npz, frozen from the
    reference).  Found by tests/tools/parity_fuzz_other.py in round 2."""
    g = load_golden("edge_zero_planes.npz")
    R = int(g["R"])
    img = torch.tensor(g["img"]).float()
    l, gr = run(S.ImplicitLoss(R, dev, 1.5, 260), img, g["pred"], dev)
    check(l, gr, g["implicit_loss"], g["implicit_grad"], what="zero planes implicit")
    l, gr = run(S.ImplicitLoss(R, dev, 1.0, 20), img, g["pred"], dev)
    check(l, gr, g["implicit_soft_loss"], g["implicit_soft_grad"], what="zero planes implicit soft")
    l, gr = run(S.ExplicitLoss(R, dev), g["true"], g["pred"], dev)
    check(l, gr, g["explicit_loss"], g["explicit_grad"], what="zero planes explicit")
    l, gr = run(S.ExplicitLoss(R, dev), g["pred"], g["true"], dev)
    check(l, gr, g["explicit_swapped_loss"], g["explicit_swapped_grad"], what="zero planes explicit swapped")
    for i in range(len(g["pred"])):
        t, p = torch.tensor(g["true"][i:i + 1]).to(dev), torch.tensor(g["pred"][i:i + 1]).to(dev)
        with torch.no_grad():
            assert abs(S.ImplicitLoss(R, dev, 1.5, 260)(img[i:i + 1].to(dev), p).item() - g["implicit_per_sample"][i]) <= LOSS_RTOL * g["implicit_per_sample"][i], i
            assert abs(S.ExplicitLoss(R, dev)(t, p).item() - g["explicit_per_sample"][i]) <= LOSS_RTOL * g["explicit_per_sample"][i], i
