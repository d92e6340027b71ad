"""Per-call CUDA-event timing of one TRAINING-mode step of the hot path (every dropout of the reference active).

    python tools/train_prof.py [rows]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cti_b200  # noqa: E402
from cti_b200 import kernels as KS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K, Q, A, G = 50, 12, 6, 2
dev = torch.device("cuda")
torch.manual_seed(1204)
att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
pools = [cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2) for _ in range(G)]
q_prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
a_prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
mods = torch.nn.ModuleList([att, *pools, *q_prj, *a_prj]).to(dev).train()
params = list(mods.parameters())
g = torch.Generator().manual_seed(1)
v = torch.relu(torch.randn(B, K, 2048, generator=g)).to(dev)
q = torch.tanh(torch.randn(B, Q, 1024, generator=g)).to(dev)
a = torch.tanh(torch.randn(B, A, 1024, generator=g)).to(dev)
cot = torch.randn(B, 1024, generator=g).to(dev)


def step():
    cti_b200.prepack(mods)
    for p in params:
        p.grad = None
    qq, aa = q.detach().requires_grad_(True), a.detach().requires_grad_(True)
    p_att, _ = att(v, qq, aa)
    qe, ae = qq, aa
    for gi in range(G):
        b_emb = pools[gi].forward_with_weights(v, qe, ae, p_att[:, :, :, :, gi])
        qe = q_prj[gi](b_emb.unsqueeze(1)) + qe
        ae = a_prj[gi](b_emb.unsqueeze(1)) + ae
    ((qe.sum(1) + ae.sum(1)) * cot).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
KS.STATS.prof = []
n = 3
for _ in range(n):
    step()
torch.cuda.synchronize()
rec, KS.STATS.prof = KS.STATS.prof, None
agg = {}
for name, tag, flops, nbytes, e0, e1 in rec:
    d = agg.setdefault((name, tag), [0, 0.0])
    d[0] += 1
    d[1] += e0.elapsed_time(e1)
tot = sum(d[1] for d in agg.values()) / n
print(f"rows {B}: {tot:.3f} ms of library calls per step (eager, events around every call)")
for (name, tag), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / n * 1e3:9.1f} us  {c // n:3d} x  {name}  {tag}")
