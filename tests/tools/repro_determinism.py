"""Run the implicit loss + backward six times on the same inputs and report whether the results are bit-identical, then
compare with the oracle.  Used with SQ_LIBSQLOSS=<variant .so> to check experimental builds (e.g. -DSQ_BWD_DEPTH=2, which
forces the on-the-spot fallback of the compacted backward)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import sq_oracle as O
import sq_recovery_b200 as S
dev = torch.device('cuda:0')
B, R = 24, 32
true = O.random_params(B, 71).to(dev)
pred = O.perturbed_params(O.random_params(B, 71), 6).to(dev)
img = S.ImplicitLoss(128, dev, 1.5, 260).depth_projection(true).unsqueeze(1).contiguous()
crit = S.ImplicitLoss(R, dev, 1.5, 260)
outs = []
for i in range(6):
    p = pred.clone().requires_grad_(True)
    l = crit(img, p); l.backward()
    outs.append((l.item(), p.grad.clone()))
for i in range(1, 6):
    d = (outs[i][1] - outs[0][1]).abs().max().item()
    print(i, outs[i][0] == outs[0][0], d, (outs[i][1] != outs[0][1]).sum().item())
# against the oracle
po = pred.cpu().clone().requires_grad_(True)
ref = O.ImplicitLoss(R, "cpu", 1.5, 260)(img.cpu(), po); ref.backward()
err = ((outs[0][1].cpu().double() - po.grad.double()).abs() / (1e-6 + 1e-4 * po.grad.double().abs())).max().item()
print("vs oracle: loss rel", abs(outs[0][0] - ref.item()) / ref.item(), "grad err (tol units)", err)
