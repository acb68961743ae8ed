"""Minimal driver for ncu: a few steps of the headline workload (ImplicitLoss fwd+bwd, B=256, R=64) and one call of
the other kernels, nothing else.  Same code path as bench.py's timed region."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sq_oracle as O      # input distributions only
import sq_recovery_b200 as S

B, R = 256, 64
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
true = O.random_params(B, 0)
pred = O.perturbed_params(true, 7).to(dev)
img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).contiguous()
crit = S.ImplicitLoss(R, dev, 1.5, 260)
for i in range(steps):
    p = pred.detach().requires_grad_(True)
    loss = crit(img, p)
    loss.backward()
if "--all" in sys.argv:
    p = pred.detach().requires_grad_(True)
    S.ExplicitLoss(R, dev)(true.to(dev), p).backward()
    S.IoUAccuracy(R, dev)(true.to(dev), pred)
    with torch.no_grad():
        crit(img, pred)
torch.cuda.synchronize()
print("loss", loss.item())
