"""Timing of the BAN student's hot path (BiAttention + 2 x BCNet pooling + q_prj), fwd+bwd, BASELINE config 3."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, cti_b200
from cti_b200 import kernels as KS
from oracle import cti_oracle as O
B = int(os.environ.get("B", 256)); K, Q, G = 50, 12, 2
dev = "cuda"
torch.manual_seed(0)
att = cti_b200.BiAttention(2048, 1024, 1024, G)
pools = [cti_b200.BCNet(2048, 1024, 1024, None, k=1) for _ in range(G)]
prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
mods = torch.nn.ModuleList([att, *pools, *prj]).to(dev).eval()
params = list(mods.parameters())
v, q, _ = O.synthetic_inputs(B, K, Q, 0, seed=1)
v, q = v.to(dev), q.to(dev)
cot = torch.randn(B, 1024, device=dev)
def step():
    for p in params: p.grad = None
    qq = q.detach().requires_grad_(True)
    p_att, _ = att.forward_all(v, qq)
    qe, lst = qq, []
    for g in range(G):
        b = pools[g].forward_with_weights(v, qe, p_att[:, g])
        qe = prj[g](b.unsqueeze(1)) + qe
        lst.append(qe)
    (torch.stack(lst, 1).sum(1).sum(1) * cot).sum().backward()
for _ in range(5): step()
g = cti_b200.GraphedStep(step, [mods], [v])
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"BAN hot path fwd+bwd B={B}: {ms:.3f} ms/step, {B/ms*1e3:.0f} rows/s")
KS.STATS.prof = []
for _ in range(3): step()
torch.cuda.synchronize()
rec, KS.STATS.prof = KS.STATS.prof, None
agg = {}
for name, tag, fl, nb, a, b in rec:
    d = agg.setdefault((name, tag), [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b)
for (n, t), (c, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  {n:28s} {t:30s} x{c//3}  {ms_/c:.4f} ms")
