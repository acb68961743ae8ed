"""Minimal driver for ncu: steps of the headline workload (ImplicitLoss fwd+bwd, B=256, R=64) over the SAME four input
sets bench.py rotates (rank 0), and optionally one call of the other kernels.  Same code path as bench.py's eager step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sq_recovery_b200 import inputs as O      # seeded randsq / randquat workloads
import sq_recovery_b200 as S

B, R = 256, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
render = S.ImplicitLoss(256, dev, 1.5, 260)
crit = S.ImplicitLoss(R, dev, 1.5, 260)
sets = []
for k in range(4):
    true = O.random_params(B, k)
    pred = O.perturbed_params(true, 7 + k).to(dev)
    sets.append((render.depth_projection(true.to(dev)).unsqueeze(1).contiguous(), pred, true.to(dev)))
torch.cuda.synchronize()
for i in range(steps):
    img, pred, _ = sets[i % 4]
    p = pred.detach().requires_grad_(True)
    loss = crit(img, p)
    loss.backward()
if "--all" in sys.argv:
    img, pred, true = sets[0]
    p = pred.detach().requires_grad_(True)
    S.ExplicitLoss(R, dev)(true, p).backward()
    S.IoUAccuracy(R, dev)(true, pred)
    with torch.no_grad():
        crit(img, pred)
    p = pred.detach().requires_grad_(True)
    S.LeastSquares(R, dev)(img, p).backward()
torch.cuda.synchronize()
print("loss", loss.item())
