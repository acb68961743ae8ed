"""Per work item of the implicit fwd+bwd kernel: the plan kernel's cost estimate next to the cycles the column kernel spent
(debug build -DSQ_ITEMLOG).  Writes gpurun_out/item_costs.npz and prints how well the hand-out order follows the real cost.

    python tools/item_costs.py [extra -D defs]
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sq_recovery_b200 import inputs as O
from sq_recovery_b200 import _lib as L0
import sq_recovery_b200 as S
from sq_recovery_b200.functional import nearest_offsets

defs = ["-DSQ_ITEMLOG"] + [f"-D{d}" for d in sys.argv[1:]]
out = "/tmp/libsq_itemlog.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-ftz=true", "-std=c++17",
                       "-shared", "-Xcompiler", "-fPIC", "-o", out, os.path.join(ROOT, "sq_recovery_b200", "csrc", "sqloss.cu")] + defs)
h = ctypes.CDLL(out)
for name, (res, args) in L0._PROTOS.items():
    fn = getattr(h, name); fn.restype, fn.argtypes = res, args
B, R = 256, 64
dev = torch.device("cuda:0")
dense = bool(os.environ.get("SQ_DENSE"))           # objects that fill the grid (bench.py's second workload)
true = O.random_params(B, 0, size_range=O.DENSE_SIZE_RANGE) if dense else O.random_params(B, 0)
pred = O.perturbed_params(true, 7).to(dev)
true = true.to(dev)
img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
row_off, col_off = nearest_offsets(256, 256, R, dev)
loss = torch.empty((), dtype=torch.float64, device=dev); grad = torch.empty_like(pred)
nb = h.sq_scratch_bytes(B, R); scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
for _ in range(4):
    assert h.sq_implicit_loss(P(pred), 0, B, R, 1.0 / (R - 1), 1e-4, P(img), 256 * 256, P(row_off), P(col_off), 1.5, 260.0,
                              P(loss), None, P(grad), None, P(scratch), nb, torch.cuda.current_stream().cuda_stream) == 0
torch.cuda.synchronize()
n = B * (R * R // 32)
plan = np.zeros((n, 4), dtype=np.float32); item = np.zeros((n, 4), dtype=np.int32)
h.sq_debug_itemlog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
assert h.sq_debug_itemlog(plan.ctypes.data_as(ctypes.c_void_p), item.ctypes.data_as(ctypes.c_void_p), n) == 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "item_costs_dense.npz" if dense else "item_costs.npz"), plan=plan, item=item)
live = item[:, 0] > 0
cyc = item[live, 0].astype(np.float64)
print(f"items with work {live.sum()} of {n}; cycles per item: mean {cyc.mean():.0f} p50 {np.median(cyc):.0f} p90 {np.percentile(cyc, 90):.0f} max {cyc.max():.0f}")
rk = lambda a: np.argsort(np.argsort(a))
print(f"  rank correlation of the real cycles with the plan kernel's plane count: {np.corrcoef(rk(plan[live, 1]), rk(cyc))[0, 1]:.3f}")
walk = item[live, 1].astype(np.float64)
q = item[live, 2].astype(np.float64); q = np.where(q >= 1000, q - 1000, q); qr = item[live, 3].astype(np.float64)
X = np.stack([np.ones_like(cyc), walk, q, qr], 1)
co = np.linalg.lstsq(X, cyc, rcond=None)[0]
print(f"  cycles ~ {co[0]:.0f} + {co[1]:.2f} x walk cycles + {co[2]:.0f} x queued points + {co[3]:.0f} x refined points "
      f"(R^2 {1 - ((cyc - X @ co) ** 2).sum() / ((cyc - cyc.mean()) ** 2).sum():.3f})")

# list scheduling of the measured item durations on the kernel's warps: what a better hand-out order could buy
import heapq
warps = 148 * 5 * 4


def makespan(order):
    hp = [0.0] * warps
    heapq.heapify(hp)
    for i in order:
        heapq.heappush(hp, heapq.heappop(hp) + cyc[i])
    return max(hp)


rng = np.random.RandomState(0)
cls = plan[live, 0].astype(int)
by_class = np.concatenate([rng.permutation(np.where(cls == k)[0]) for k in range(cls.max() + 1)])
print(f"list scheduling on {warps} warps (cycles): mean load {cyc.sum() / warps:.0f}; hand-out by cost class (as the kernel does) "
      f"{makespan(by_class):.0f}; by exact plane count {makespan(np.argsort(-plan[live, 1], kind='stable')):.0f}; "
      f"clairvoyant longest-first {makespan(np.argsort(-cyc)):.0f}; random {makespan(rng.permutation(len(cyc))):.0f}")
