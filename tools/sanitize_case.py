"""Small end-to-end case for compute-sanitizer: every kernel once at odd sizes (masked lanes, non-patched layouts)."""
import os
import sys

import subprocess

import torch

if "--bounds" in sys.argv:      # build a library with index traps and run this case on it (sets SQ_LIBSQLOSS)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = "/tmp/libsq_bounds.so"
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-ftz=true", "-std=c++17", "-shared",
                           "-Xcompiler", "-fPIC", "-DSQ_DEBUG_BOUNDS", "-o", out, os.path.join(root, "sq_recovery_b200", "csrc", "sqloss.cu")])
    os.environ["SQ_LIBSQLOSS"] = out

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sq_recovery_b200 import inputs as O      # seeded randsq / randquat workloads
import sq_recovery_b200 as S
from sq_recovery_b200.functional import HostContext

dev = torch.device("cuda:0")
for B, R in ((3, 8), (5, 24), (2, 33), (7, 16), (300, 64), (1, 128)):
    true, pred = O.random_params(B, 1).to(dev), O.random_params(B, 2).to(dev)
    img = S.ImplicitLoss(4 * R, dev, 1.5, 260).depth_projection(true).unsqueeze(1)
    p = pred.clone().requires_grad_(True)
    S.ImplicitLoss(R, dev, 1.5, 260)(img, p).backward()
    p = pred.clone().requires_grad_(True)
    S.ExplicitLoss(R, dev)(true, p).backward()
    S.IoUAccuracy(R, dev)(true, pred)
    p = pred.clone().requires_grad_(True)
    S.LeastSquares(R, dev)(img, p).backward()
    S.ExplicitLoss(R, dev).occupancy(pred)
    S.IoUAccuracy(R, dev).ins_outs(pred)
    with torch.no_grad():
        S.ImplicitLoss(R, dev)(img, pred)
ctx = HostContext(0)
hi = torch.rand(4, 1, 40, 56).pin_memory()
ctx.implicit_loss(O.random_params(4, 3).numpy(), hi.numpy(), 16, 1.5, 260.0)
ctx.implicit_loss(O.random_params(4, 3).numpy(), torch.rand(4, 1, 40, 56).numpy(), 16, 1.5, 260.0)   # pageable
ctx.explicit_loss(O.random_params(4, 3).numpy(), O.random_params(4, 4).numpy(), 12)
ctx.iou_counts(O.random_params(4, 3).numpy(), O.random_params(4, 4).numpy(), 12)
ctx.close()
torch.cuda.synchronize()
print("sanitize case done")
