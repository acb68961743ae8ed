"""Time ImplicitLoss(256).depth_projection on a batch of 256 (the data-generation path)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sq_recovery_b200 import inputs as O      # seeded randsq / randquat workloads
import sq_recovery_b200 as S
dev = torch.device('cuda:0')
crit = S.ImplicitLoss(256, dev, 1.5, 260)
p = O.random_params(256, 0).to(dev)
for _ in range(2): d = crit.depth_projection(p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d = crit.depth_projection(p)
e1.record(); torch.cuda.synchronize()
print("R=256 B=256 render ms", e0.elapsed_time(e1) / 5)
