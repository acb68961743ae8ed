"""Per-sample loss error of the CUDA ImplicitLoss on the dense workload (debug helper for tests/tools/parity_dense.py)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sq_oracle as O
import sq_recovery_b200 as S
from sq_recovery_b200 import inputs
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
R, B = 64, 4
for seed in range(300, 306):
    true = inputs.random_params(B, seed, size_range=inputs.DENSE_SIZE_RANGE)
    for style, pred in (("rand", inputs.random_params(B, seed + 1000, size_range=inputs.DENSE_SIZE_RANGE)), ("pert", inputs.perturbed_params(true, seed))):
        with torch.no_grad():
            img = O.ImplicitLoss(2 * R, "cpu", 1.5, 260).depth_projection(true).float().unsqueeze(1)
        oc = O.ImplicitLoss(R, "cpu", 1.5, 260)
        ref = oc.per_sample(img, pred)
        crit = S.ImplicitLoss(R, dev, 1.5, 260)
        for b in range(B):
            pg = pred[b:b + 1].to(dev).requires_grad_(True)
            l = crit(img[b:b + 1].to(dev), pg); l.backward()
            with torch.no_grad():
                lf = crit(img[b:b + 1].to(dev), pred[b:b + 1].to(dev)).item()
            rel = abs(l.item() - ref[b].item()) / abs(ref[b].item())
            if rel > 5e-6:
                d = crit.depth_projection(pred[b:b + 1].to(dev))[0].double().cpu().numpy()
                dref = oc.depth_projection(pred[b:b + 1])[0].numpy()
                err = d - dref
                ij = np.unravel_index(np.abs(err).argmax(), err.shape)
                print(seed, style, b, "loss", l.item(), "fwd-only", lf, "ref", ref[b].item(), "rel", rel, "| depth err max", np.abs(err).max(), "mean", err.mean(),
                      "n>1e-6:", int((np.abs(err) > 1e-6).sum()), "at", ij, "dref", dref[ij], "params", pred[b].numpy().round(4).tolist())
