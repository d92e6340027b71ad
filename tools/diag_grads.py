"""Diagnostic (GPU): per-parameter gradient error of the drop-in TriAttention path vs the oracle,
for a cotangent applied (a) directly to the logits, (b) to the attention p (through softmax)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cti_b200
from oracle import cti_oracle as O

B, K, Q, A, G = int(os.environ.get("B", 8)), 50, 12, 6, 2
params = O.random_cti_params(glimpse=G, seed=1204)
v, q, a = O.synthetic_inputs(B, K, Q, A, seed=1212)
gen = torch.Generator().manual_seed(3)
cotL = torch.randn(B, K, Q, A, G, generator=gen)
cotP = torch.randn(B, K, Q, A, G, generator=gen)

def run(mode):
    pl = {k: t.clone().requires_grad_(True) for k, t in params.items() if k.startswith("v_att.")}
    ql, al = q.clone().requires_grad_(True), a.clone().requires_grad_(True)
    p_ref, l_ref = O.tri_attention(v, ql, al, pl, "v_att.TriAtt.")
    fin = torch.isfinite(l_ref)
    if mode == "logits":
        (torch.where(fin, l_ref, torch.zeros(())) * cotL).sum().backward()
    else:
        (p_ref * cotP * 3600).sum().backward()
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
    att.load_state_dict({k[len("v_att."):]: t for k, t in params.items() if k.startswith("v_att.")})
    att.cuda().eval()
    qd, ad = q.cuda().requires_grad_(True), a.cuda().requires_grad_(True)
    p, l = att(v.cuda(), qd, ad)
    if mode == "logits":
        (torch.where(torch.isfinite(l), l, torch.zeros((), device="cuda")) * cotL.cuda()).sum().backward()
    else:
        (p * cotP.cuda() * 3600).sum().backward()
    rows = []
    def add(name, got, ref):
        got = got.detach().float().cpu(); ref = ref.detach()
        rows.append((name, ((got - ref).abs().max() / ref.abs().max()).item(), ((got - ref).norm() / ref.norm()).item()))
    add("dq", qd.grad, ql.grad); add("da", ad.grad, al.grad)
    groups = {}
    for k, p_ in att.named_parameters():
        ref = pl["v_att." + k].grad
        kk = k.replace("TriAtt.", "")
        import re
        kk = re.sub(r"_net\.\d+\.", "_net.*.", kk)
        got = p_.grad.detach().float().cpu()
        d = groups.setdefault(kk, [0., 0., 0., 0.])
        d[0] = max(d[0], ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item())
        d[1] += (got - ref).pow(2).sum().item(); d[2] += ref.pow(2).sum().item()
    print(f"== cotangent on {mode}")
    for n, e1, e2 in rows: print(f"  {n:40s} maxrel {e1:.4f}  normrel {e2:.4f}")
    for kk, d in groups.items(): print(f"  {kk:40s} maxrel {d[0]:.4f}  normrel {(d[1]/max(d[2],1e-30))**.5:.4f}")
run("logits"); run("p")
