"""The two quaternion helpers the reference losses call (torch/quaternion.py:19-21, 46-67).

Inside the CUDA kernels these are fused into the per-sample prologue (csrc/sq_core.cuh: prep_sample); the torch
versions below exist so scripts that import them from the reference's ``quaternion`` module keep working.  They
are plain tensor expressions on whatever device the input lives on.
"""
import torch


def conjugate(quaternion: torch.Tensor) -> torch.Tensor:
    """(x, y, z, w) -> (-x, -y, -z, w).  torch/quaternion.py:19-21."""
    return torch.cat((-quaternion[..., :3], quaternion[..., 3:]), dim=-1)


def mat_from_quaternion(quaternion: torch.Tensor) -> torch.Tensor:
    """Rotation matrix of ONE (x, y, z, w) quaternion, shape (1, 3, 3); the quaternion is NOT normalised.

    torch/quaternion.py:46-67 (callers take ``[0]``).
    """
    x, y, z, w = quaternion[..., 0], quaternion[..., 1], quaternion[..., 2], quaternion[..., 3]
    x2, y2, z2 = x + x, y + y, z + z
    rows = (1.0 - (y2 * y + z2 * z), y2 * x - z2 * w, z2 * x + y2 * w,
            y2 * x + z2 * w, 1.0 - (x2 * x + z2 * z), z2 * y - x2 * w,
            z2 * x - y2 * w, z2 * y + x2 * w, 1.0 - (x2 * x + y2 * y))
    return torch.stack(rows, dim=-1).reshape(1, 3, 3)
