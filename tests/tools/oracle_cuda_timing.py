"""BASELINE config 2's comparator: the reference's per-sample torch-op sequence (oracle loop form, fp64) run on the SAME
GPU, against the fused path, on a slice of the config-2 workload.  (The reference tree itself cannot travel to the GPU box;
the oracle's loop form issues the same ~1.6 k aten calls per sample.)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sq_oracle as O
import sq_recovery_b200 as S

dev = torch.device("cuda:0")
B, R = 32, 64
true = O.random_params(B, 0); pred = O.perturbed_params(true, 7)
img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).contiguous()
from oracle import ref_import
if ref_import.available() or ref_import.staged():       # the UNMODIFIED reference classes with device='cuda:0' 
    oc = ref_import.load()[0].ImplicitLoss(R, dev, 1.5, 260)
    ref_name = "reference torch/classes.py ImplicitLoss on cuda:0 (fp64, oracle/_ref)"
else:
    oc = O.ImplicitLoss(R, dev, 1.5, 260, form="loop")
    ref_name = "oracle loop form on cuda:0 (fp64)"
crit = S.ImplicitLoss(R, dev, 1.5, 260)
for name, fn in ((ref_name, oc), ("sq_recovery_b200", crit)):
    for rep in range(3):
        p = pred.to(dev).requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        l = fn(img, p); l.backward()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: B={B} R={R} fwd+bwd {dt * 1e3:.2f} ms (wall, eager) = {B * R ** 3 / dt / 1e9:.3f} Gpoints/s, loss {l.item():.8f}")
