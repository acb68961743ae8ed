"""One-off parity check of ImplicitLoss fwd+bwd at large grids (R = 128, 96) against the fp64 oracle (slow on the CPU)."""
import os, sys, torch, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sq_oracle as O
import sq_recovery_b200 as S
dev = torch.device('cuda:0')
torch.set_num_threads(16)
for R, B in ((128, 2), (96, 2)):
    true = O.random_params(B, 81); pred = O.perturbed_params(true, 9)
    img = S.ImplicitLoss(256, dev, 1.5, 260).depth_projection(true.to(dev)).unsqueeze(1).contiguous()
    t0 = time.time()
    po = pred.clone().requires_grad_(True)
    ref = O.ImplicitLoss(R, "cpu", 1.5, 260)(img.cpu(), po); ref.backward()
    pg = pred.to(dev).requires_grad_(True)
    l = S.ImplicitLoss(R, dev, 1.5, 260)(img, pg); l.backward()
    err = ((pg.grad.cpu().double() - po.grad.double()).abs() / (1e-6 + 1e-4 * po.grad.double().abs())).max().item()
    print(f"R={R}: loss rel err {abs(l.item()-ref.item())/ref.item():.2e}, worst gradient entry {err:.2f}x tolerance, oracle {time.time()-t0:.0f}s")
