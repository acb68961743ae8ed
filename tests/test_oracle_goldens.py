"""The CPU oracle (oracle/sq_oracle.py) against vectors frozen from the unmodified reference.

These pin the oracle: tests/golden/*.npz were produced by oracle/make_goldens.py from
/root/reference/torch/classes.py.  Everything here runs on the CPU.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, random_golden_files
from oracle import sq_oracle as O
from oracle import ref_import


def grad_of(crit, true, pred):
    pred = pred.clone().requires_grad_(True)
    loss = crit(true, pred)
    loss.backward()
    return loss.item(), pred.grad.double().numpy()


def close(a, b, rtol=1e-10, atol=1e-13):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("form", ["batch", "loop"])
def test_fixture_images_and_labels(fixtures_golden, form):
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    lab = torch.tensor(g["labels"])
    roll = lab.roll(1, 0)
    crit = O.ImplicitLoss(64, "cpu", 1.5, 260, form=form)
    l, gr = grad_of(crit, imgs, lab)
    close(l, g["implicit64_loss"]); close(gr, g["implicit64_grad"])
    assert abs(l - 0.007077226864072453) < 1e-15          # SURVEY 4 golden table
    l, gr = grad_of(crit, imgs, roll)
    close(l, g["implicit64_roll_loss"]); close(gr, g["implicit64_roll_grad"])
    close(crit.depth_projection(lab).numpy(), g["implicit64_depth"])
    close(crit.per_sample(imgs, lab).numpy(), g["implicit64_per_sample"])
    assert (g["implicit64_per_sample_next"] > 10 * g["implicit64_per_sample"]).all()
    l, gr = grad_of(O.ImplicitLoss(32, "cpu", form=form), imgs, roll)
    close(l, g["implicit32_default_loss"]); close(gr, g["implicit32_default_grad"])
    ex = O.ExplicitLoss(32, "cpu", form=form)
    assert ex(lab, lab).item() == 0.0
    l, gr = grad_of(ex, lab, roll)
    close(l, g["explicit32_roll_loss"]); close(gr, g["explicit32_roll_grad"])
    assert abs(l - 4.868018668436481) < 1e-12
    acc = O.IoUAccuracy(64, "cpu", form=form)
    assert acc(lab, lab).item() == 1.0
    close(acc(lab, roll).item(), g["iou64_roll"], rtol=1e-7)
    i, u = acc.counts(lab, roll)
    assert (i.numpy() == g["iou64_roll_inter"]).all() and (u.numpy() == g["iou64_roll_union"]).all()
    close(O.IoUAccuracy(64, "cpu", reduce=False, form=form)(lab, roll).numpy(), g["iou64_roll_per_sample"])


def test_fixture_least_squares(fixtures_golden):
    g = fixtures_golden
    imgs = torch.tensor(g["imgs_u8"].astype(np.float32) / 255.0)
    lab = torch.tensor(g["labels"])
    ls = O.LeastSquares(64, "cpu")
    l, gr = grad_of(ls, imgs, lab)                        # fp32 internals in the reference (classes.py:319)
    close(l, g["lsq64_loss"], rtol=1e-5); close(gr, g["lsq64_grad"], rtol=1e-4, atol=1e-6)
    l, gr = grad_of(ls, imgs, lab.roll(1, 0))
    close(l, g["lsq64_roll_loss"], rtol=1e-5); close(gr, g["lsq64_roll_grad"], rtol=1e-4, atol=1e-4)


def test_fixture_main_and_visu(fixtures_golden):
    g = fixtures_golden
    main = torch.tensor(g["main_params"])
    assert O.IoUAccuracy(64, "cpu")(main, main).item() == 1.0 == float(g["main_iou64"])   # classes.py:453-473
    t, p = torch.tensor(g["visu_true"]), torch.tensor(g["visu_pred"])
    l, gr = grad_of(O.ExplicitLoss(32, "cpu"), t, p)      # visu.py:142-165 call pattern, fp64 leaf
    close(l, g["visu_explicit32_loss"]); close(gr, g["visu_explicit32_grad"])
    close(O.IoUAccuracy(128, "cpu")(t, p).item(), g["visu_iou128"], rtol=1e-7)


@pytest.mark.parametrize("fname", random_golden_files())
def test_random_cases(fname):
    g = load_golden(fname)
    R = int(g["R"])
    true, img = torch.tensor(g["true"]), torch.tensor(g["img"])
    for tag in ("far", "near"):
        pred = torch.tensor(g[f"pred_{tag}"])
        for name, crit in (("implicit_t15_k260", O.ImplicitLoss(R, "cpu", 1.5, 260)),
                           ("implicit_default", O.ImplicitLoss(R, "cpu"))):
            l, gr = grad_of(crit, img, pred)
            close(l, g[f"{name}_{tag}_loss"]); close(gr, g[f"{name}_{tag}_grad"], rtol=1e-9, atol=1e-12)
        l, gr = grad_of(O.ExplicitLoss(R, "cpu"), true, pred)
        close(l, g[f"explicit_{tag}_loss"]); close(gr, g[f"explicit_{tag}_grad"], rtol=1e-9, atol=1e-12)
        i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
        assert (i.numpy() == g[f"iou_{tag}_inter"]).all() and (u.numpy() == g[f"iou_{tag}_union"]).all()
        l, gr = grad_of(O.LeastSquares(R, "cpu"), img, pred)
        close(l, g[f"lsq_{tag}_loss"], rtol=1e-5); close(gr, g[f"lsq_{tag}_grad"], rtol=1e-3, atol=1e-5)
    close(O.ImplicitLoss(R, "cpu", 1.5, 260).depth_projection(true).numpy(), g["depth_true"])


def test_edge_cases(edge_golden):
    g = edge_golden
    R = int(g["R"])
    true, pred, img = torch.tensor(g["true"]), torch.tensor(g["pred"]), torch.tensor(g["img"]).float()
    l, gr = grad_of(O.ImplicitLoss(R, "cpu", 1.5, 260), img, pred)
    close(l, g["implicit_loss"]); close(gr, g["implicit_grad"], rtol=1e-9, atol=1e-12)
    l, gr = grad_of(O.ImplicitLoss(R, "cpu", 1.0, 20), img, pred)
    close(l, g["implicit_soft_loss"]); close(gr, g["implicit_soft_grad"], rtol=1e-9, atol=1e-12)
    l, gr = grad_of(O.ExplicitLoss(R, "cpu"), true, pred)
    close(l, g["explicit_loss"]); close(gr, g["explicit_grad"], rtol=1e-9, atol=1e-12)
    # clamp sub-gradient: 0 outside the range, 1 on the inclusive boundary (SURVEY 4 edge tests)
    assert gr[0, 0] == 0 and gr[0, 1] == 0 and gr[0, 2] != 0
    assert gr[1, 3] == 0 and gr[1, 4] != 0
    assert gr[2, 5] == 0 and gr[2, 6] == 0 and gr[2, 7] != 0
    close(O.ExplicitLoss(R, "cpu").per_sample(true, pred).numpy(), g["explicit_per_sample"])
    close(O.ImplicitLoss(R, "cpu", 1.5, 260).per_sample(img, pred).numpy(), g["implicit_per_sample"])
    i, u = O.IoUAccuracy(R, "cpu").counts(true, pred)
    assert (i.numpy() == g["iou_inter"]).all() and (u.numpy() == g["iou_union"]).all()
    ex24 = O.ExplicitLoss(24, "cpu")
    assert ex24.xyz.shape[1] == int(g["explicit24_n"]) == 26     # arange(0,1+1/24,1/24) has R+2 entries
    l, gr = grad_of(ex24, torch.tensor(g["true24"]), torch.tensor(g["pred24"]))
    close(l, g["explicit24_loss"]); close(gr, g["explicit24_grad"], rtol=1e-9, atol=1e-12)


def test_zero_plane_cases():
    """Axis-aligned rotations with t_z exactly on a grid plane: the exact-zero fix-up on a whole z plane (frozen from
    the reference by oracle/make_goldens.py:zero_plane_cases)."""
    g = load_golden("edge_zero_planes.npz")
    R = int(g["R"])
    true, pred, img = torch.tensor(g["true"]), torch.tensor(g["pred"]), torch.tensor(g["img"]).float()
    for form in ("batch", "loop"):
        l, gr = grad_of(O.ImplicitLoss(R, "cpu", 1.5, 260, form=form), img, pred)
        close(l, g["implicit_loss"]); close(gr, g["implicit_grad"], rtol=1e-9, atol=1e-12)
        l, gr = grad_of(O.ExplicitLoss(R, "cpu", form=form), true, pred)
        close(l, g["explicit_loss"]); close(gr, g["explicit_grad"], rtol=1e-9, atol=1e-12)
    l, gr = grad_of(O.ImplicitLoss(R, "cpu", 1.0, 20), img, pred)
    close(l, g["implicit_soft_loss"]); close(gr, g["implicit_soft_grad"], rtol=1e-9, atol=1e-12)
    close(O.ExplicitLoss(R, "cpu").per_sample(true, pred).numpy(), g["explicit_per_sample"])
    close(O.ImplicitLoss(R, "cpu", 1.5, 260).per_sample(img, pred).numpy(), g["implicit_per_sample"])


def test_grazing_case():
    """One big object on a 24^3 grid at tau = 3 (oracle/make_goldens.py:grazing_case, the parity fuzz's former outlier)."""
    g = load_golden("edge_grazing.npz")
    pred, img = torch.tensor(g["pred"]), torch.tensor(g["img"]).float()
    for form in ("batch", "loop"):
        l, gr = grad_of(O.ImplicitLoss(int(g["R"]), "cpu", float(g["tau"]), float(g["k"]), form=form), img, pred)
        close(l, g["implicit_loss"]); close(gr, g["implicit_grad"], rtol=1e-9, atol=1e-12)


def test_quaternion_helpers():
    q = torch.tensor([0.699625, 0.378123, -0.090419, -0.599476], dtype=torch.float64)
    assert torch.equal(O.conjugate(q), torch.tensor([-0.699625, -0.378123, 0.090419, -0.599476], dtype=torch.float64))
    m = O.mat_from_quaternion(q / q.norm())
    close((m @ m.T).numpy(), np.eye(3), atol=1e-12)
    close(torch.det(m).item(), 1.0)
    close(O.mat_from_quaternion(O.conjugate(q)).numpy(), O.mat_from_quaternion(q).T.numpy())
    close(O.mat_from_quaternion(torch.tensor([0., 0., 0., 1.], dtype=torch.float64)).numpy(), np.eye(3))
    # not normalised: scaling q by s scales the off-identity part by s^2 (quaternion.py:46-67)
    close((O.mat_from_quaternion(2 * q) - torch.eye(3)).numpy(), 4 * (O.mat_from_quaternion(q) - torch.eye(3)).numpy())


@pytest.mark.skipif(not (ref_import.available() or ref_import.staged()),
                    reason="reference neither mounted nor staged under oracle/_ref (python -m oracle.build_ref)")
def test_live_reference_matches_oracle():
    """Where /root/reference is mounted, run the real classes next to the oracle on fresh inputs."""
    rc, rq = ref_import.load()
    true, pred = O.random_params(3, 41), O.random_params(3, 42)
    img = torch.rand(3, 1, 64, 64, generator=torch.Generator().manual_seed(0))
    cpu = torch.device("cpu")
    for ref, mine, args in ((rc.ImplicitLoss(16, cpu, 1.5, 260), O.ImplicitLoss(16, "cpu", 1.5, 260, form="loop"), (img, pred)),
                            (rc.ExplicitLoss(16, cpu), O.ExplicitLoss(16, "cpu", form="loop"), (true, pred)),
                            (rc.LeastSquares(16, cpu), O.LeastSquares(16, "cpu"), (img, pred))):
        l0, g0 = grad_of(ref, *args)
        l1, g1 = grad_of(mine, *args)
        close(l1, l0, rtol=1e-12); close(g1, g0, rtol=1e-9, atol=1e-13)
    close(O.IoUAccuracy(16, "cpu")(true, pred).item(), rc.IoUAccuracy(16, cpu)(true, pred).item(), rtol=1e-7)
    q = torch.tensor(O.randquat(np.random.RandomState(0)))
    close(O.mat_from_quaternion(q).numpy(), rq.mat_from_quaternion(q)[0].numpy(), rtol=0, atol=0)
    close(O.conjugate(q).numpy(), rq.conjugate(q).numpy(), rtol=0, atol=0)
