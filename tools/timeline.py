"""Per-warp timeline of the implicit fwd+bwd kernel (debug build with -DSQ_TIMELINE): when each persistent warp
started, fetched its last item and finished.  Shows how much of the kernel is end-game (warps idle, work left elsewhere).

    python tools/timeline.py [extra -D defs]          (SQ_BWD_STATS: also count gradient-carrying lanes per backward
                                                      block -- slows the kernel, the timeline is then not meaningful)
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sq_recovery_b200 import inputs as O      # seeded randsq / randquat workloads
from sq_recovery_b200 import _lib as L0
import sq_recovery_b200 as S
from sq_recovery_b200.functional import nearest_offsets

defs = ["-DSQ_TIMELINE"] + [f"-D{d}" for d in sys.argv[1:]]
out = "/tmp/libsq_timeline.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-ftz=true", "-std=c++17",
                       "-shared", "-Xcompiler", "-fPIC", "-o", out,
                       os.path.join(ROOT, os.environ.get("SQ_SRC", os.path.join("sq_recovery_b200", "csrc")), "sqloss.cu")] + defs)
h = ctypes.CDLL(out)
for name, (res, args) in L0._PROTOS.items():
    fn = getattr(h, name); fn.restype, fn.argtypes = res, args
B, R = int(os.environ.get("SQ_B", 256)), 64
dev = torch.device("cuda:0")
true = O.random_params(B, 0, size_range=O.DENSE_SIZE_RANGE) if os.environ.get("SQ_DENSE") else O.random_params(B, 0)
pred = O.perturbed_params(true, 7)
if os.environ.get("SQ_SORT"):
    vol = pred[:, 0].clamp(0.05, 1) * pred[:, 1].clamp(0.05, 1) * pred[:, 2].clamp(0.05, 1)
    order = torch.argsort(vol, descending=os.environ["SQ_SORT"] == "desc")
    true, pred = true[order].contiguous(), pred[order].contiguous()
true, pred = true.to(dev), pred.to(dev)
img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
row_off, col_off = nearest_offsets(256, 256, R, dev)
loss = torch.empty((), dtype=torch.float64, device=dev); grad = torch.empty_like(pred)
nb = h.sq_scratch_bytes(B, R); scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for _ in range(5):
    rc = h.sq_implicit_loss(P(pred), 0, B, R, 1.0 / (R - 1), 1e-4, P(img), 256 * 256, P(row_off), P(col_off), 1.5, 260.0,
                            P(loss), None, P(grad), None, P(scratch), nb, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
n = 8192
buf = np.zeros(3 * n, dtype=np.uint64)
h.sq_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert h.sq_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), n) == 0
buf = buf.reshape(n, 3)
buf = buf[buf[:, 1] > 0]
t0 = buf[:, 0].min()
start, end = (buf[:, 0] - t0) / 1e3, (buf[:, 1] - t0) / 1e3
items = buf[:, 2] >> np.uint64(40)
last_fetch = (buf[:, 2] & np.uint64((1 << 40) - 1)) / 1e3 + start
print(f"warps {len(buf)}  kernel span {end.max():.1f} us")
print(f"start  : min {start.min():.1f} p50 {np.median(start):.1f} max {start.max():.1f}")
print(f"end    : min {end.min():.1f} p10 {np.percentile(end,10):.1f} p50 {np.median(end):.1f} p90 {np.percentile(end,90):.1f} max {end.max():.1f}")
print(f"last item fetched at: p50 {np.median(last_fetch):.1f} max {last_fetch.max():.1f};  duration of last item: p50 {np.median(end-last_fetch):.1f} p90 {np.percentile(end-last_fetch,90):.1f} max {(end-last_fetch).max():.1f}")
print(f"items per warp: min {items.min()} p50 {np.median(items)} max {items.max()}")
busy = (end - start).sum() / (len(buf) * end.max())
print(f"warp-slot utilisation (sum of warp lifetimes / warps x span): {busy:.3f}")

cls = np.zeros(32, dtype=np.uint32)
h.sq_debug_classes.argtypes = [ctypes.c_void_p]
if h.sq_debug_classes(cls.ctypes.data_as(ctypes.c_void_p)) == 0:
    print("items per cost class (0 = longest ... last non-zero = certified empty):", [int(c) for c in cls if c])

pl = np.zeros(16, dtype=np.uint64)
h.sq_debug_plan.argtypes = [ctypes.c_void_p]
if h.sq_debug_plan(pl.ctypes.data_as(ctypes.c_void_p)) == 0:
    for o, who in ((0, "thread 0 (prep)"), (8, "thread 100 (pixel sum)")):
        n = int(pl[o + 7]); t = (pl[o:o + n] - pl[o]).astype(np.int64) / 1e3
        print(f"plan kernel block 7, {who}: stage stamps (us) {t.round(2).tolist()}")

bw = np.zeros(2, dtype=np.uint64)
h.sq_debug_bwd.argtypes = [ctypes.c_void_p]
if h.sq_debug_bwd(bw.ctypes.data_as(ctypes.c_void_p)) == 0 and bw[0]:
    print(f"backward blocks executed (all calls so far): {int(bw[0])} warp-steps, {bw[1] / bw[0]:.1f} of 32 lanes carrying gradient on average")

if "SQ_PHASES" in sys.argv[1:]:       # python tools/timeline.py SQ_PHASES: SM cycles of all warps per phase (the stamps cost ~3 %)
    ph = np.zeros(8, dtype=np.uint64)
    h.sq_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert h.sq_debug_phases(ph.ctypes.data_as(ctypes.c_void_p), 1) == 0          # discard the warm-up calls
    reps = 4
    for _ in range(reps):
        assert h.sq_implicit_loss(P(pred), 0, B, R, 1.0 / (R - 1), 1e-4, P(img), 256 * 256, P(row_off), P(col_off), 1.5, 260.0,
                                  P(loss), None, P(grad), None, P(scratch), nb, torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert h.sq_debug_phases(ph.ctypes.data_as(ctypes.c_void_p), 0) == 0
    names = ["item set-up (sample, pixels, fp32 base, ranges)", "exact base + z walk", "claim issue, signs, work-queue look-up",
             "fp64 refinement", "deal-out backward", "column values, folding into the tile",
             "item epilogue (reduction, partial row)", "waiting for the next item (claim, sample), loop ends"]
    tot = float(ph.sum())
    print(f"warp cycles per launch, all {len(buf)} warps: {tot / reps / 1e6:.2f} M  ({tot / reps / len(buf) / 1e3:.1f} k per warp)")
    for nme, c in zip(names, ph):
        print(f"  {nme:52s} {100.0 * float(c) / tot:5.1f} %   {float(c) / reps / len(buf) / 1e3:6.2f} k cycles per warp")
