"""Instrumented runs of the implicit column kernel (libsqloss_count.so: the same sources built with -DSQ_COUNT).

Measurement support for bench.py and tools/: how many z-plane steps the kernel walked, how many points went through the
compacted backward and its fp64 refinement -- and from those the fraction of the grid that was evaluated and the number
of MUFU-pipe instructions issued.  Never on the timed path: the counting build does an atomic per event."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from .functional import nearest_offsets

LIB_COUNT = os.path.join(_lib.HERE, "libsqloss_count.so")
_h = None

# MUFU-pipe (XU) warp instructions per counted event, read off csrc/sq_core.cuh / sqloss.cu (int <-> float and
# fp64 <-> fp32 conversions share the pipe: profiles/peaks_r02.json D2F / I2F):
#   plane step        3 lg2|s| + 2 (ex2 + lg2) + ex2 (F) + ex2, rcp (sigmoid) + ex2 (transmittance)              = 11
#   on-the-spot bwd   5 rcp                                                                                      = 5
#   deal-out round    point_forward 8 + point_backward 5 rcp                                                     = 13
#   refinement round  ex2, rcp + 9 fp64<->fp32 conversions                                                       = 11
#   column group      11 conversions of the fp64 column base + 2 int->double grid coordinates + 3 int->float     = 16
#   work item         sqrt (culling) + 3 conversions of the range bounds + 2 int->float grid coordinates         = 6
# Validated against ncu (sm__inst_executed_pipe_xu of the committed capture): the model is 10 % below the hardware
# count (profiles/implicit_kernel_ncu_summary_r02.json) -- most of the difference are the POPCs of the pool appends, which
# share the pipe (profiles/peaks_r02.json EX2_POPC) but are overhead, not work -- i.e. the reported fraction is conservative.
XU_PER_EVENT = {"plane_steps": 11, "spot_backward_blocks": 5, "dealout_rounds": 13, "refine_rounds": 11,
                "column_groups": 16, "items": 6}
NAMES = ("plane_steps", "spot_backward_blocks", "dealout_rounds", "refine_rounds", "items", "column_groups",
         "queued_points", "refined_points")


def lib():
    global _h
    if _h is None:
        if not os.path.exists(LIB_COUNT):
            raise RuntimeError(f"{LIB_COUNT} not found: python -m sq_recovery_b200.build")
        h = ctypes.CDLL(LIB_COUNT)
        for name, (res, args) in _lib._PROTOS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        h.sq_debug_counters.restype = ctypes.c_int
        h.sq_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _h = h
    return _h


def implicit_counts(images: torch.Tensor, pred: torch.Tensor, n: int, tau: float, sharpness: float, want_grad: bool = True):
    """One sq_implicit_loss call on the counting build; returns the counters and what they imply."""
    h = lib()
    dev = pred.device
    B = pred.shape[0]
    p = pred.detach().float().contiguous()
    img = images.detach().float().contiguous()
    row_off, col_off = nearest_offsets(img.shape[2], img.shape[3], n, dev)
    nb = h.sq_scratch_bytes(B, n)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
    loss = torch.empty((), dtype=torch.float64, device=dev)
    grad = torch.empty_like(p) if want_grad else None
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None      # noqa: E731
    cnt = (ctypes.c_ulonglong * 8)()
    torch.cuda.synchronize()
    _lib.check(h.sq_debug_counters(cnt, 1), "sq_debug_counters")
    rc = h.sq_implicit_loss(P(p), _lib.SQ_F32, B, n, 1.0 / (n - 1), 1e-4, P(img), img.shape[2] * img.shape[3], P(row_off),
                            P(col_off), tau, sharpness, P(loss), None, P(grad), None, P(scratch), nb,
                            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "sq_implicit_loss (counting build)")
    torch.cuda.synchronize()
    _lib.check(h.sq_debug_counters(cnt, 1), "sq_debug_counters")
    c = dict(zip(NAMES, (int(v) for v in cnt)))
    grid_points = B * n ** 3
    xu = sum(XU_PER_EVENT[k] * c[k] for k in XU_PER_EVENT)
    return {"counters": c, "grid_points": grid_points, "point_evaluations": 32 * c["plane_steps"],
            "walked_fraction": 32 * c["plane_steps"] / grid_points, "xu_warp_inst_model": xu,
            "loss": loss.item()}
