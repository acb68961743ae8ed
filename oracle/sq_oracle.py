"""CPU restatement of the sq-recovery loss hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Arithmetic is torch fp64 on the host, like the reference
(``torch/classes.py:139,233,395`` upcast to double; ``LeastSquares`` stays in
the input dtype because ``params.double()`` is commented out at ``:319``).

Two evaluation forms are provided for each loss:

* ``form="loop"``  -- one sample at a time, the same sequence of tensor ops per
  sample as the reference's Python loop.  This is what ``bench.py`` times as the
  CPU baseline (it has the reference's cost structure).
* ``form="batch"`` -- one broadcast expression over the whole batch; used by the
  tests because it is ~B times fewer op dispatches.  Same maths.

Parameter row layout (``torch/train.py:89``): ``[a1 a2 a3 | e1 e2 | t1 t2 t3 | qx qy qz qw]``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

ZERO_FIX = 1e-4  # torch/classes.py:126,171-173,221,261-263,342-344


# --------------------------------------------------------------------------- quaternion helpers
def conjugate(q: torch.Tensor) -> torch.Tensor:
    """(x,y,z,w) -> (-x,-y,-z,w).  torch/quaternion.py:19-21."""
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def mat_from_quaternion(q: torch.Tensor) -> torch.Tensor:
    """Rotation matrix of an (x,y,z,w) quaternion, NOT normalised.  torch/quaternion.py:46-67.

    Returns ``(..., 3, 3)``; the reference returns ``(1,3,3)`` for one quaternion and callers take ``[0]``.
    """
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    m = torch.stack(
        (1.0 - (tyy + tzz), txy - twz, txz + twy,
         txy + twz, 1.0 - (txx + tzz), tyz - twx,
         txz - twy, tyz + twx, 1.0 - (txx + tyy)), dim=-1)
    return m.reshape(q.shape[:-1] + (3, 3))


def preprocess_sq(p: torch.Tensor) -> torch.Tensor:
    """Clamp a to [0.05,1], e to [0.1,1], t to [0,1]; q untouched.  torch/classes.py:129-136,224-230,310-316."""
    a = torch.clamp(p[..., 0:3], min=0.05, max=1)
    e = torch.clamp(p[..., 3:5], min=0.1, max=1)
    t = torch.clamp(p[..., 5:8], min=0, max=1)
    return torch.cat([a, e, t, p[..., 8:12]], dim=-1)


# --------------------------------------------------------------------------- grids
def explicit_axis(render_size: int) -> np.ndarray:
    """``np.arange(0, 1+step, step)`` -- n = R+1 (R+2 for R=24,96).  torch/classes.py:122-123."""
    step = 1 / render_size
    return np.arange(0, 1 + step, step).astype(np.float64)


def linspace_axis(render_size: int) -> np.ndarray:
    """``np.linspace(0, 1, R)``.  torch/classes.py:218,389."""
    return np.linspace(0, 1, render_size).astype(np.float64)


def make_xyz(axis: np.ndarray, fix_zero: bool) -> torch.Tensor:
    """(3,n,n,n) ij-meshgrid; zeros -> 1e-4 when ``fix_zero``.  torch/classes.py:124-126,219-221,390-391."""
    r = torch.tensor(axis)
    xyz = torch.stack(torch.meshgrid([r, r, r], indexing="ij"))
    if fix_zero:
        xyz = torch.where(xyz == 0, xyz + ZERO_FIX, xyz)
    return xyz


# --------------------------------------------------------------------------- inside-outside function
def _fix(v: torch.Tensor) -> torch.Tensor:
    # in-place ``v[v == 0] += 1e-4`` of the reference; gradient passes straight through (classes.py:171-173)
    return torch.where(v == 0, v + ZERO_FIX, v)


def inside_outside_single(p: torch.Tensor, pts: torch.Tensor, clamp: bool, fix_zero: bool) -> torch.Tensor:
    """F for ONE sample on points ``pts`` of shape (3, ...).  torch/classes.py:142-184,236-273,398-424,322-351."""
    if clamp:
        p = preprocess_sq(p)
    a, e, t, q = p[0:3], p[3:5], p[5:8], p[8:12]
    rot = mat_from_quaternion(conjugate(q))
    tr = torch.matmul(rot, t)
    cs = torch.einsum("ij,jabc->iabc" if pts.dim() == 4 else "ij,ja->ia", rot, pts)
    A1 = torch.pow((cs[0] - tr[0]) / a[0], 2)
    B1 = torch.pow((cs[1] - tr[1]) / a[1], 2)
    C1 = torch.pow((cs[2] - tr[2]) / a[2], 2)
    if fix_zero:                    # same in-place masked add as the reference (classes.py:171-173)
        A1[A1 == 0] += ZERO_FIX
        B1[B1 == 0] += ZERO_FIX
        C1[C1 == 0] += ZERO_FIX
    A = torch.pow(A1, 1 / e[1])
    B = torch.pow(B1, 1 / e[1])
    C = torch.pow(C1, 1 / e[0])
    E = torch.pow(A + B, e[1] / e[0])
    return torch.pow(E + C, e[0])


def inside_outside_batch(p: torch.Tensor, xyz: torch.Tensor, clamp: bool, fix_zero: bool) -> torch.Tensor:
    """F for a batch on a shared grid ``xyz`` (3,n,n,n) -> (B,n,n,n).  Same maths as inside_outside_single."""
    if clamp:
        p = preprocess_sq(p)
    a, e, t, q = p[:, 0:3], p[:, 3:5], p[:, 5:8], p[:, 8:12]
    rot = mat_from_quaternion(conjugate(q))                      # (B,3,3)
    tr = torch.einsum("bij,bj->bi", rot, t)                      # (B,3)
    cs = torch.einsum("bij,jxyz->bixyz", rot, xyz)               # (B,3,n,n,n)
    s = (cs - tr[:, :, None, None, None]) / a[:, :, None, None, None]
    sq = torch.pow(s, 2)
    if fix_zero:
        sq = _fix(sq)
    e1 = e[:, 0, None, None, None]
    e2 = e[:, 1, None, None, None]
    A = torch.pow(sq[:, 0], 1 / e2)
    B = torch.pow(sq[:, 1], 1 / e2)
    C = torch.pow(sq[:, 2], 1 / e1)
    E = torch.pow(A + B, e2 / e1)
    return torch.pow(E + C, e1)


# --------------------------------------------------------------------------- losses
class ExplicitLoss:
    """MSE*100 between occupancy grids of true and predicted params.  torch/classes.py:109-201."""

    def __init__(self, render_size, device="cpu", reduce=True, form="batch"):
        self.render_size, self.device, self.reduce, self.form = render_size, torch.device(device), reduce, form
        self.axis = explicit_axis(render_size)
        self.xyz = make_xyz(self.axis, fix_zero=True).to(self.device)

    def occupancy(self, p):
        p = p.double()
        if self.form == "batch":
            f = inside_outside_batch(p, self.xyz, True, True)
        else:
            f = torch.stack([inside_outside_single(p[i], self.xyz, True, True) for i in range(p.shape[0])])
        return torch.sigmoid(5 * (1 - f))                                   # classes.py:187

    def __call__(self, true, pred):
        a, b = self.occupancy(true), self.occupancy(pred)
        if self.form == "batch":
            per = torch.pow(a - b, 2).flatten(1).mean(dim=1) * 100          # classes.py:198
        else:
            per = torch.stack([torch.mean(torch.pow(ai - bi, 2)) * 100 for ai, bi in zip(a, b)])
        return torch.mean(per)                                              # classes.py:200

    def per_sample(self, true, pred):
        a, b = self.occupancy(true), self.occupancy(pred)
        return torch.pow(a - b, 2).flatten(1).mean(dim=1) * 100


class ImplicitLoss:
    """MAE between the (nearest-resized) depth image and a soft depth render of the predicted SQ.

    torch/classes.py:203-295.
    """

    def __init__(self, render_size, device="cpu", tau=1, sigmoid_sharpness=100, reduce=True, form="batch"):
        self.render_size, self.device, self.reduce, self.form = render_size, torch.device(device), reduce, form
        self.tau, self.sigmoid_sharpness = tau, sigmoid_sharpness
        self.axis = linspace_axis(render_size)
        self.xyz = make_xyz(self.axis, fix_zero=True).to(self.device)

    def _render(self, f):
        o = torch.sigmoid(self.sigmoid_sharpness * (1 - f))                 # classes.py:274
        cd = torch.exp(-self.tau * torch.cumsum(o.flip(dims=[-1]), dim=-1))  # classes.py:277
        depth = 1 - cd.sum(dim=-1) / self.render_size                       # classes.py:278
        return depth.transpose(-1, -2).flip(dims=(-2,))                     # classes.py:279  img[row,col]=depth[col,n-1-row]

    def depth_projection(self, p):
        p = p.double()
        if self.form == "batch":
            return self._render(inside_outside_batch(p, self.xyz, True, True))
        return torch.stack([self._render(inside_outside_single(p[i], self.xyz, True, True))
                            for i in range(p.shape[0])])

    def resize(self, true):
        return F.interpolate(true, size=(self.render_size, self.render_size), mode="nearest")  # classes.py:286

    def per_sample(self, true, pred):
        t = self.resize(true)
        d = self.depth_projection(pred).unsqueeze(1)
        return torch.abs(t - d).flatten(1).mean(dim=1)                      # classes.py:292

    def __call__(self, true, pred):
        if self.form == "batch":
            return torch.mean(self.per_sample(true, pred))                  # classes.py:294
        t = self.resize(true)
        d = self.depth_projection(pred).unsqueeze(1)
        return torch.mean(torch.stack([torch.mean(torch.abs(ai - bi)) for ai, bi in zip(t, d)]))


class IoUAccuracy:
    """IoU of the F<=1 grids; no clamp, no zero fix-up.  torch/classes.py:374-447."""

    def __init__(self, render_size, device="cpu", reduce=True, full=False, form="batch"):
        self.render_size, self.device, self.reduce, self.full, self.form = \
            render_size, torch.device(device), reduce, full, form
        self.axis = linspace_axis(render_size)
        self.xyz = make_xyz(self.axis, fix_zero=False).to(self.device)

    def ins_outs(self, p):
        p = p.double()
        if self.form == "batch":
            return inside_outside_batch(p, self.xyz, False, False)
        return torch.stack([inside_outside_single(p[i], self.xyz, False, False) for i in range(p.shape[0])])

    def counts(self, true, pred):
        """(intersection[B], union[B]) int64 voxel counts.  classes.py:433-438."""
        a_bin, b_bin = self.ins_outs(true) <= 1, self.ins_outs(pred) <= 1
        return (a_bin & b_bin).flatten(1).sum(dim=1), (a_bin | b_bin).flatten(1).sum(dim=1)

    def __call__(self, true, pred):
        inter, union = self.counts(true, pred)
        if not self.reduce:
            return inter.double() / union.double()                          # classes.py:441-445
        return torch.sum(inter) / torch.sum(union)                          # classes.py:439


class LeastSquares:
    """Solina-Bajcsy energy on the points back-projected from the depth image.  torch/classes.py:297-371."""

    def __init__(self, render_size, device="cpu", reduce=True, form="loop"):
        self.render_size, self.device, self.reduce = render_size, torch.device(device), reduce

    def points(self, true):
        """Per-sample (3,m) lists (col/R, 1-row/R, depth).  classes.py:359-369."""
        t = F.interpolate(true, size=(self.render_size, self.render_size), mode="nearest")
        out = []
        for i in range(true.shape[0]):
            rc = torch.where(t[i][0] > 0)
            x = rc[0].float() / self.render_size
            y = rc[1].float() / self.render_size
            out.append(torch.stack([y, 1 - x, t[i][0][rc]]))
        return out

    def energy_function(self, batch_points, params):
        res = []
        for i in range(params.shape[0]):
            p = preprocess_sq(params[i])
            f = inside_outside_single(params[i], batch_points[i], True, True)
            res.append(torch.pow(torch.sqrt(p[0] * p[1] * p[2]) * (f - 1), 2).sum())   # classes.py:354-355
        return torch.stack(res)

    def __call__(self, true, pred):
        return self.energy_function(self.points(true), pred).mean()         # classes.py:370-371


# --------------------------------------------------------------------------- input distributions
def randquat(rng: np.random.RandomState) -> np.ndarray:
    """Uniform unit quaternion.  torch/quaternion.py:139-145."""
    u = rng.uniform(0, 1, (3,))
    return np.array([np.sqrt(1 - u[0]) * np.sin(2 * np.pi * u[1]),
                     np.sqrt(1 - u[0]) * np.cos(2 * np.pi * u[1]),
                     np.sqrt(u[0]) * np.sin(2 * np.pi * u[2]),
                     np.sqrt(u[0]) * np.cos(2 * np.pi * u[2])])


def randsq(rng: np.random.RandomState) -> np.ndarray:
    """a~U(.1,.3)^3, e~U(.1,1)^2, t~U(.34,.65)^3.  torch/visu.py:55-56."""
    return np.concatenate((rng.uniform(0.1, 0.3, (3,)), rng.uniform(0.1, 1, (2,)), rng.uniform(0.34, 0.65, (3,))))


def random_params(batch: int, seed: int, dtype=torch.float32) -> torch.Tensor:
    """(B,12) rows of randsq()+randquat() from ``np.random.RandomState(seed)``."""
    rng = np.random.RandomState(seed)
    rows = [np.concatenate((randsq(rng), randquat(rng))) for _ in range(batch)]
    return torch.tensor(np.stack(rows), dtype=dtype)


def perturbed_params(true: torch.Tensor, seed: int, sigma: float = 0.02) -> torch.Tensor:
    """pred = true + N(0, sigma), quaternion re-normalised (train-like gradients)."""
    g = torch.Generator().manual_seed(seed)
    p = true.double() + sigma * torch.randn(true.shape, generator=g, dtype=torch.float64)
    p[:, 8:12] = p[:, 8:12] / p[:, 8:12].norm(dim=1, keepdim=True)
    return p.to(true.dtype)
