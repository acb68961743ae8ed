"""Synthetic superquadric parameters for benchmarks and harnesses.

The reference draws its test objects with ``randsq()`` (torch/visu.py:55-56; torch/test_random.py:34-37 is the same
box after the /255 scaling) and orientations with ``randquat()`` (torch/quaternion.py:139-145).  These are the same
distributions, seeded, so that bench.py, harness/ and tools/ build their workloads without touching ``oracle/`` (which
is test infrastructure).  Rows are ``[a1 a2 a3 | e1 e2 | t1 t2 t3 | qx qy qz qw]`` (torch/train.py:89).
"""
from __future__ import annotations

import numpy as np
import torch

SIZE_RANGE = (0.1, 0.3)          # visu.py:55
SHAPE_RANGE = (0.1, 1.0)
POSITION_RANGE = (0.34, 0.65)
DENSE_SIZE_RANGE = (0.5, 1.0)    # "dense" bench workload: objects that fill the grid, where culling cannot help


def randquat(rng: np.random.RandomState) -> np.ndarray:
    """One quaternion (x, y, z, w) uniform on the unit 3-sphere from three uniform numbers (quaternion.py:139-145)."""
    u1, u2, u3 = rng.uniform(0, 1, (3,))
    r1, r2 = np.sqrt(1 - u1), np.sqrt(u1)
    a, b = 2 * np.pi * u2, 2 * np.pi * u3
    return np.array([r1 * np.sin(a), r1 * np.cos(a), r2 * np.sin(b), r2 * np.cos(b)])


def randsq(rng: np.random.RandomState, size_range=SIZE_RANGE) -> np.ndarray:
    """Sizes, shapes, position of one object: 3 + 2 + 3 uniform numbers in that order (visu.py:55-56)."""
    a = rng.uniform(size_range[0], size_range[1], (3,))
    e = rng.uniform(SHAPE_RANGE[0], SHAPE_RANGE[1], (2,))
    t = rng.uniform(POSITION_RANGE[0], POSITION_RANGE[1], (3,))
    return np.concatenate((a, e, t))


def random_params(batch: int, seed: int, dtype=torch.float32, size_range=SIZE_RANGE) -> torch.Tensor:
    """(batch, 12) rows, each ``randsq()`` followed by ``randquat()`` from one ``RandomState(seed)`` stream."""
    rng = np.random.RandomState(seed)
    out = np.empty((batch, 12), dtype=np.float64)
    for i in range(batch):
        out[i, :8] = randsq(rng, size_range)
        out[i, 8:] = randquat(rng)
    return torch.tensor(out, dtype=dtype)


def perturbed_params(true: torch.Tensor, seed: int, sigma: float = 0.02) -> torch.Tensor:
    """A prediction near ``true``: every entry + N(0, sigma), quaternion brought back to unit length -- the kind of
    gradient a half-trained network sees."""
    gen = torch.Generator().manual_seed(seed)
    noisy = true.double() + sigma * torch.randn(true.shape, generator=gen, dtype=torch.float64)
    noisy[:, 8:12] = noisy[:, 8:12] / noisy[:, 8:12].norm(dim=1, keepdim=True)
    return noisy.to(true.dtype)
