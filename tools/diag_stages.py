"""Diagnostic (GPU): stage-by-stage gradient comparison of TriLogitsFn.backward against the oracle
evaluated with the kernels' bf16 rounding points."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cti_b200
from cti_b200 import kernels as KS, functions as F_
from oracle import cti_oracle as O

B, K, Q, A, G, R = int(os.environ.get("B", 8)), 50, 12, 6, 2, 32
params = O.random_cti_params(glimpse=G, seed=1204)
v, q, a = O.synthetic_inputs(B, K, Q, A, seed=1212)
cot = torch.randn(B, K, Q, A, G, generator=torch.Generator().manual_seed(3))
pre = "v_att.TriAtt."
# ---- oracle with rounding, keeping intermediates
pl = {k: t.clone().requires_grad_(True) for k, t in params.items() if k.startswith("v_att.")}
ql, al = q.clone().requires_grad_(True), a.clone().requires_grad_(True)
with O.bf16_rounding():
    vt = O.fcnet(v, pl, pre + "v_tucker.", dropout=0.5); vt.retain_grad()
    qt = O.fcnet(ql, pl, pre + "q_tucker.", dropout=0.2); qt.retain_grad()
    at = O.fcnet(al, pl, pre + "a_tucker.", dropout=0.2); at.retain_grad()
    vc = torch.stack([O.fcnet(vt, pl, f"{pre}v_net.{r}.", dropout=0.5) for r in range(R)], 2); vc.retain_grad()
    qc = torch.stack([O.fcnet(qt, pl, f"{pre}q_net.{r}.", dropout=0.2) for r in range(R)], 2); qc.retain_grad()
    ac = torch.stack([O.fcnet(at, pl, f"{pre}a_net.{r}.", dropout=0.2) for r in range(R)], 2); ac.retain_grad()
    logits = O.trilinear_closed(vc, qc, ac, O.teff_from_tg(pl[pre + "T_g"]))
    mask = O.zero_row_mask(v)
    lm = logits.masked_fill(mask[:, :, None, None, None], float("-inf"))
    p = torch.softmax(lm.reshape(B, -1, G), 1).view(B, K, Q, A, G)
    mode = os.environ.get("MODE", "p")
    if mode == "p":
        (p * cot * 3600).sum().backward()
    else:
        (torch.where(torch.isfinite(lm), lm, torch.zeros(())) * cot).sum().backward()
ref = {"dzv": (vc.grad * (vc > 0)).reshape(B * K, -1), "dzq": (qc.grad * (qc > 0)).reshape(B * Q, -1),
       "dza": (ac.grad * (ac > 0)).reshape(B * A, -1), "dzvt": (vt.grad * (vt > 0)).reshape(B * K, -1),
       "dzqt": (qt.grad * (qt > 0)).reshape(B * Q, -1), "dzat": (at.grad * (at > 0)).reshape(B * A, -1),
       "vc": vc.detach().reshape(B * K, -1), "yv": vt.detach().reshape(B * K, -1)}
# ---- GPU, recording
rec = {}
orig_tb, orig_lb, orig_lf = KS.trilinear_bwd, F_.lin_bwd, F_.lin_fwd
def tb(*args):
    out = orig_tb(*args); rec["dl"] = args[4]; rec["dzv"], rec["dzq"], rec["dza"] = out[0], out[1], out[2]; return out
n = [0]
def lb(x, dz, V, g, pk, ng, need_dx, dx_relu_aux=None, dx_f32=False):
    out = orig_lb(x, dz, V, g, pk, ng, need_dx, dx_relu_aux, dx_f32)
    names = ["vn", "qn", "an", "vt", "qt", "at"]
    rec["dz_in_" + names[n[0]]] = dz; rec["dV_" + names[n[0]]] = out[0]; rec["dx_" + names[n[0]]] = out[2]; n[0] += 1
    return out
nf = [0]
def lf(x, pk, bias, relu, out_bf16=True, out_f32=False):
    out = orig_lf(x, pk, bias, relu, out_bf16, out_f32)
    rec["y%d" % nf[0]] = out[0]; nf[0] += 1
    return out
KS.trilinear_bwd = tb; F_.lin_bwd = lb; F_.lin_fwd = lf
att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
att.load_state_dict({k[len("v_att."):]: t for k, t in params.items() if k.startswith("v_att.")})
att.cuda().eval()
qd, ad = q.cuda().requires_grad_(True), a.cuda().requires_grad_(True)
pg, lg = att(v.cuda(), qd, ad)
if mode == "p":
    (pg * cot.cuda() * 3600).sum().backward()
else:
    (torch.where(torch.isfinite(lg), lg, torch.zeros((), device="cuda")) * cot.cuda()).sum().backward()
def cmp(name, got, want):
    got = got.detach().float().cpu(); want = want.detach().float()
    agree = (got != 0) & (want != 0)
    if agree.any() and not agree.all():
        ga, wa = got[agree], want[agree]
        print(f"{name:10s} [mask-agree subset] maxrel {((ga-wa).abs().max()/wa.abs().max()).item():.4f} normrel {((ga-wa).norm()/wa.norm()).item():.4f}")
    print(f"{name:10s} maxrel {((got-want).abs().max()/want.abs().max()).item():.4f} normrel {((got-want).norm()/want.norm()).item():.4f}"
          f"  |want| {want.abs().max().item():.3e} mismatch-zero {((got==0)!=(want==0)).float().mean().item():.5f}")
cmp("yv", rec["y0"], ref["yv"]); cmp("vc", rec["y3"], ref["vc"])
cmp("p", pg, p); 
cmp("dl", rec["dl"].permute(0, 2, 3, 4, 1), torch.nan_to_num(logits.grad) if logits.grad is not None else lm.grad) if False else None
cmp("dzv", rec["dzv"], ref["dzv"]); cmp("dzq", rec["dzq"], ref["dzq"]); cmp("dza", rec["dza"], ref["dza"])
cmp("dzvt", rec["dx_vn"], ref["dzvt"]); cmp("dzqt", rec["dx_qn"], ref["dzqt"]); cmp("dzat", rec["dx_an"], ref["dzat"])
for nm, key in (("vt", "v_tucker"), ("qt", "q_tucker"), ("at", "a_tucker")):
    cmp("dV_" + nm, rec["dV_" + nm], pl[f"{pre}{key}.main.1.weight_v"].grad)
cmp("dV_vn", rec["dV_vn"], torch.cat([pl[f"{pre}v_net.{r}.main.1.weight_v"].grad for r in range(R)], 0))
# wgrad recomputed in fp32 torch from the kernel's own dz and x:
xv = cti_b200.fc.cast_features(v.cuda())[0]
dW = rec["dz_in_vt"].float().t() @ xv.float()
dWref = ref["dzvt"].t() @ v.reshape(B * K, -1).to(torch.bfloat16).float()
cmp("dWeff_vt(torch from kernel dz)", dW, dWref)
