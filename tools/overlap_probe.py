"""How much of the step's fixed cost (plan / finalize / end-game of the persistent kernel) disappears when two independent
batches are in flight on two streams: the four steps of bench.py's graph, captured sequentially vs on two interleaved streams.

    python tools/overlap_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sq_recovery_b200 import inputs as O
import sq_recovery_b200 as S

B, R = 256, 64
dev = torch.device("cuda:0")
render = S.ImplicitLoss(256, dev, 1.5, 260)
crit = S.ImplicitLoss(R, dev, 1.5, 260)
sets = []
for k in range(4):
    true = O.random_params(B, k)
    sets.append((render.depth_projection(true.to(dev)).unsqueeze(1).contiguous(), O.perturbed_params(true, 7 + k).to(dev)))


def step(k):
    img, pred = sets[k]
    p = pred.detach().requires_grad_(True)
    loss = crit(img, p)
    loss.backward()
    return loss, p


def timed(g, n=200):
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n / 4 * 1e3          # us per step


s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for s in (s1, s2):
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for k in range(4):
            step(k); step(k)
torch.cuda.synchronize()
seq = torch.cuda.CUDAGraph()
keep = []
with torch.cuda.graph(seq, stream=s1):
    for k in range(4):
        keep.append(step(k))
two = torch.cuda.CUDAGraph()
with torch.cuda.graph(two, stream=s1):
    s2.wait_stream(s1)
    for k in range(4):
        with torch.cuda.stream(s1 if k % 2 == 0 else s2):
            keep.append(step(k))
    s1.wait_stream(s2)
s3, s4 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for s in (s3, s4):
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for k in range(4):
            step(k); step(k)
torch.cuda.synchronize()
four = torch.cuda.CUDAGraph()
keep4 = []
with torch.cuda.graph(four, stream=s1):
    for s in (s2, s3, s4):
        s.wait_stream(s1)
    for k, s in enumerate((s1, s2, s3, s4)):
        with torch.cuda.stream(s):
            keep4.append(step(k))
    for s in (s2, s3, s4):
        s1.wait_stream(s)
print(f"sequential: {timed(seq):.2f} us/step   two streams: {timed(two):.2f} us/step   four streams: {timed(four):.2f} us/step")
ref = [keep[k][0].item() for k in range(4)]
got = [keep[4 + k][0].item() for k in range(4)]
print("same losses:", ref == got, "same grads:", all(torch.equal(keep[k][1].grad, keep[4 + k][1].grad) for k in range(4)))
