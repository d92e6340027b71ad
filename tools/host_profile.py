"""Where does the HOST time of an eagerly issued step go?  (cProfile of a few steps, GPU box only.)"""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, cti_b200
from oracle import cti_oracle as O
dev = "cuda"
B, K, Q, A, G = 1024, 50, 12, 6, 2
torch.manual_seed(0)
att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
pools = [cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2) for _ in range(G)]
qp = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
ap = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
mods = torch.nn.ModuleList([att, *pools, *qp, *ap]).to(dev).eval()
params = list(mods.parameters())
v, q, a = [t.to(dev) for t in O.synthetic_inputs(B, K, Q, A, seed=1)]
cot = torch.randn(B, 1024, device=dev)
def step():
    for p in params: p.grad = None
    qq, aa = q.detach().requires_grad_(True), a.detach().requires_grad_(True)
    p_att, _ = att(v, qq, aa)
    qe, ae = qq, aa
    for g in range(G):
        b = pools[g].forward_with_weights(v, qe, ae, p_att[:, :, :, :, g])
        qe = qp[g](b.unsqueeze(1)) + qe
        ae = ap[g](b.unsqueeze(1)) + ae
    ((qe.sum(1) + ae.sum(1)) * cot).sum().backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
