"""Module-level parity on the B200: the drop-in nn.Modules (through the C ABI) against
  * the reference's own outputs stored in tests/golden/cti_golden.pt (made by importing the
    reference, tests/golden/make_golden.py), and
  * the pinned CPU oracle on seeded synthetic inputs at the real model sizes.
Tolerances are the north_star's: logits / attention <= 2e-2 max-abs, gradients <= 3e-2 relative
(max-abs error over max-abs of the reference gradient), bf16 compute with fp32 accumulation."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402

DEV = "cuda"
ABS_TOL = 2e-2
GRAD_TOL = 3e-2


def rel(x, ref):
    return ((x.detach().float().cpu() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20)).item()


def maxabs(x, ref):
    return (x.detach().float().cpu() - ref.float()).abs().max().item()


def check_param_grads(module, ref_grads, tol=GRAD_TOL, prefix=""):
    worst = 0.0
    for k, p in module.named_parameters():
        assert p.grad is not None, f"{prefix}{k} received no gradient"
        assert p.grad.shape == p.shape, k
        e = rel(p.grad, ref_grads[k])
        worst = max(worst, e)
        assert e <= tol, (prefix + k, e)
    return worst


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("name", ["fc_relu", "fc_lin", "fc_nodrop"])
def test_fcnet_against_reference_golden(golden, name):
    g = golden[name]
    m = cti_b200.FCNet([24, 40], g["act"], g["dropout"]).to(DEV).eval()
    m.load_state_dict(g["sd"])
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x)
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    assert maxabs(y, g["y"]) <= ABS_TOL
    (y * g["cot"].to(DEV)).sum().backward()
    assert rel(x.grad, g["dx"]) <= GRAD_TOL
    check_param_grads(m, g["grads"])


def test_tri_attention_and_pool_against_reference_golden(golden):
    g = golden["tri_d16"]
    c = g["cfg"]
    att = cti_b200.TriAttention(c["v_dim"], c["q_dim"], c["q_dim"], c["h_mm"], 1, c["rank"], c["G"], 1).to(DEV).eval()
    att.load_state_dict(g["att_sd"])
    pools = [cti_b200.TCNet(c["v_dim"], c["q_dim"], c["q_dim"], c["h_mm"], 1, c["rank"], 1, k=c["k_pool"]).to(DEV).eval()
             for _ in range(c["G"])]
    for m, sd in zip(pools, g["pool_sd"]):
        m.load_state_dict(sd)
    v = g["v"].to(DEV)
    q = g["q"].to(DEV).requires_grad_(True)
    a = g["a"].to(DEV).requires_grad_(True)
    p, logits = att(v, q, a)
    assert p.shape == g["p"].shape and logits.shape == g["logits"].shape
    inf_ref = torch.isinf(g["logits"])
    assert torch.equal(torch.isinf(logits).cpu(), inf_ref)                     # -inf positions match exactly
    assert maxabs(logits.cpu()[~inf_ref], g["logits"][~inf_ref]) <= ABS_TOL
    assert maxabs(p, g["p"]) <= ABS_TOL
    assert torch.allclose(p.sum((1, 2, 3)).cpu(), torch.ones(c["B"], c["G"]), atol=1e-5)
    pooled = [pools[i].forward_with_weights(v, q, a, p[:, :, :, :, i]) for i in range(c["G"])]
    for o, ref in zip(pooled, g["pooled"]):
        assert o.shape == ref.shape
        assert rel(o, ref) <= ABS_TOL
    loss = sum((o * ct.to(DEV)).sum() for o, ct in zip(pooled, g["cot"]))
    loss.backward()
    assert rel(q.grad, g["dq"]) <= GRAD_TOL
    assert rel(a.grad, g["da"]) <= GRAD_TOL
    check_param_grads(att, g["att_grads"], prefix="att.")
    for m, gr in zip(pools, g["pool_grads"]):
        check_param_grads(m, gr, prefix="pool.")


def test_bi_attention_and_pool_against_reference_golden(golden):
    g = golden["bi_c128"]
    c = g["cfg"]
    att = cti_b200.BiAttention(c["v_dim"], c["q_dim"], c["hid"], c["G"]).to(DEV).eval()
    att.load_state_dict(g["att_sd"])
    pools = [cti_b200.BCNet(c["v_dim"], c["q_dim"], c["hid"], None, k=1).to(DEV).eval() for _ in range(c["G"])]
    for m, sd in zip(pools, g["pool_sd"]):
        m.load_state_dict(sd)
    v = g["v"].to(DEV)
    q = g["q"].to(DEV).requires_grad_(True)
    p, logits = att.forward_all(v, q)
    inf_ref = torch.isinf(g["logits"])
    assert torch.equal(torch.isinf(logits).cpu(), inf_ref)
    assert maxabs(logits.cpu()[~inf_ref], g["logits"][~inf_ref]) <= ABS_TOL * max(1.0, g["logits"][~inf_ref].abs().max().item())
    assert maxabs(p, g["p"]) <= ABS_TOL
    pooled = [pools[i].forward_with_weights(v, q, p[:, i]) for i in range(c["G"])]
    for o, ref in zip(pooled, g["pooled"]):
        assert rel(o, ref) <= ABS_TOL
    sum((o * ct.to(DEV)).sum() for o, ct in zip(pooled, g["cot"])).backward()
    assert rel(q.grad, g["dq"]) <= GRAD_TOL
    check_param_grads(att, g["att_grads"], prefix="att.")
    for m, gr in zip(pools, g["pool_grads"]):
        check_param_grads(m, gr, prefix="pool.")


# --------------------------------------------------------------------------- #
def build_cti(params, G, device):
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1)
    pools = [cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2) for _ in range(G)]
    prj = [(cti_b200.FCNet([1024, 1024], '', .2), cti_b200.FCNet([1024, 1024], '', .2)) for _ in range(G)]
    att.load_state_dict({k[len("v_att."):]: v for k, v in params.items() if k.startswith("v_att.")})
    for i in range(G):
        pools[i].load_state_dict({k[len(f"t_net.{i}."):]: v for k, v in params.items() if k.startswith(f"t_net.{i}.")})
        prj[i][0].load_state_dict({k[len(f"q_prj.{i}."):]: v for k, v in params.items() if k.startswith(f"q_prj.{i}.")})
        prj[i][1].load_state_dict({k[len(f"a_prj.{i}."):]: v for k, v in params.items() if k.startswith(f"a_prj.{i}.")})
    mods = [att] + pools + [m for pr in prj for m in pr]
    for m in mods:
        m.to(device).eval()
    return att, pools, prj


def cti_forward(att, pools, prj, v, q, a):
    """The hot-path slice of TanModel.forward (reference src/MC/base_model.py:143-150)."""
    p, logits = att(v, q, a)
    for g in range(len(pools)):
        b_emb = pools[g].forward_with_weights(v, q, a, p[:, :, :, :, g])
        q = prj[g][0](b_emb.unsqueeze(1)) + q
        a = prj[g][1](b_emb.unsqueeze(1)) + a
    return q.sum(1) + a.sum(1), p, logits


@pytest.mark.parametrize("B,A", [(8, 6), (5, 3)])
def test_cti_hot_path_against_oracle_full_size(B, A):
    K, Q, G = 50, 12, 2
    params = O.random_cti_params(glimpse=G, seed=1204)
    v, q, a = O.synthetic_inputs(B, K, Q, A, seed=1204 + B)
    pl = {k: t.clone().requires_grad_(True) for k, t in params.items()}
    ql, al = q.clone().requires_grad_(True), a.clone().requires_grad_(True)
    joint_ref, p_ref, logits_ref = O.cti_hot_path(v, ql, al, pl, G)
    gen = torch.Generator().manual_seed(3)
    cot = torch.randn(joint_ref.shape, generator=gen)
    (joint_ref * cot).sum().backward()

    att, pools, prj = build_cti(params, G, DEV)
    qd, ad = q.to(DEV).requires_grad_(True), a.to(DEV).requires_grad_(True)
    joint, p, logits = cti_forward(att, pools, prj, v.to(DEV), qd, ad)
    inf_ref = torch.isinf(logits_ref)
    assert torch.equal(torch.isinf(logits).cpu(), inf_ref)
    assert maxabs(logits.cpu()[~inf_ref], logits_ref.detach()[~inf_ref]) <= ABS_TOL
    assert maxabs(p, p_ref.detach()) <= ABS_TOL
    assert rel(joint, joint_ref.detach()) <= ABS_TOL
    (joint * cot.to(DEV)).sum().backward()
    assert rel(qd.grad, ql.grad) <= GRAD_TOL
    assert rel(ad.grad, al.grad) <= GRAD_TOL
    named = [("v_att.", att)] + [(f"t_net.{i}.", m) for i, m in enumerate(pools)]
    named += [(f"q_prj.{i}.", pr[0]) for i, pr in enumerate(prj)] + [(f"a_prj.{i}.", pr[1]) for i, pr in enumerate(prj)]
    for prefix, m in named:
        check_param_grads(m, {k: pl[prefix + k].grad for k, _ in m.named_parameters()}, prefix=prefix)


def test_ban_hot_path_against_oracle_full_size():
    B, K, Q, G = 6, 50, 12, 2
    params = O.random_ban_params(glimpse=G, seed=1204)
    v, q, _ = O.synthetic_inputs(B, K, Q, 0, seed=77)
    pl = {k: t.clone().requires_grad_(True) for k, t in params.items()}
    ql = q.clone().requires_grad_(True)
    joint_ref, p_ref, logits_ref = O.ban_hot_path(v, ql, pl, G)
    cot = torch.randn(joint_ref.shape, generator=torch.Generator().manual_seed(5))
    (joint_ref * cot).sum().backward()

    att = cti_b200.BiAttention(2048, 1024, 1024, G)
    pools = [cti_b200.BCNet(2048, 1024, 1024, None, k=1) for _ in range(G)]
    prj = [cti_b200.FCNet([1024, 1024], '', .2) for _ in range(G)]
    att.load_state_dict({k[len("v_att."):]: t for k, t in params.items() if k.startswith("v_att.")})
    for i in range(G):
        pools[i].load_state_dict({k[len(f"b_net.{i}."):]: t for k, t in params.items() if k.startswith(f"b_net.{i}.")})
        prj[i].load_state_dict({k[len(f"q_prj.{i}."):]: t for k, t in params.items() if k.startswith(f"q_prj.{i}.")})
    for m in [att] + pools + prj:
        m.to(DEV).eval()
    vd = v.to(DEV)
    qd = q.to(DEV).requires_grad_(True)
    p, logits = att.forward_all(vd, qd)
    qe, q_list = qd, []
    for g in range(G):
        b_emb = pools[g].forward_with_weights(vd, qe, p[:, g])
        qe = prj[g](b_emb.unsqueeze(1)) + qe
        q_list.append(qe)
    joint = torch.stack(q_list, 1).sum(1).sum(1)
    inf_ref = torch.isinf(logits_ref)
    assert torch.equal(torch.isinf(logits).cpu(), inf_ref)
    scale = max(1.0, logits_ref.detach()[~inf_ref].abs().max().item())
    assert maxabs(logits.cpu()[~inf_ref], logits_ref.detach()[~inf_ref]) <= ABS_TOL * scale
    assert maxabs(p, p_ref.detach()) <= ABS_TOL
    assert rel(joint, joint_ref.detach()) <= ABS_TOL
    (joint * cot.to(DEV)).sum().backward()
    assert rel(qd.grad, ql.grad) <= GRAD_TOL
    named = [("v_att.", att)] + [(f"b_net.{i}.", m) for i, m in enumerate(pools)] + [(f"q_prj.{i}.", m) for i, m in enumerate(prj)]
    for prefix, m in named:
        check_param_grads(m, {k: pl[prefix + k].grad for k, _ in m.named_parameters()}, prefix=prefix)


def test_attention_properties_at_bench_size():
    """Size-independent properties at the benchmark batch: rows of p sum to one, masked regions get
    zero weight and -inf logits, results are deterministic, and a batch equals its two halves."""
    B, K, Q, A, G = 256, 50, 12, 6, 2
    torch.manual_seed(1204)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1).to(DEV).eval()
    v, q, a = O.synthetic_inputs(B, K, Q, A, seed=9)
    v, q, a = v.to(DEV), q.to(DEV), a.to(DEV)
    with torch.no_grad():
        p, logits = att(v, q, a)
        p2, _ = att(v, q, a)
        ph, _ = att(v[: B // 2].contiguous(), q[: B // 2].contiguous(), a[: B // 2].contiguous())
    assert torch.equal(p, p2)
    assert torch.equal(p[: B // 2], ph)
    assert torch.allclose(p.sum((1, 2, 3)), torch.ones(B, G, device=DEV), atol=1e-4)
    mask = v.abs().sum(2) == 0
    assert mask.any()
    assert torch.all(p[mask] == 0) and torch.all(torch.isinf(logits[mask]))
    assert torch.isfinite(logits[~mask]).all()


def test_modules_in_train_mode_with_dropout_zero_equal_eval():
    torch.manual_seed(0)
    m = cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, dropout=[0, 0], k=2).to(DEV)
    v, q, a = [t.to(DEV) for t in O.synthetic_inputs(4, 20, 12, 6, seed=3)]
    w = torch.softmax(torch.randn(4, 20 * 12 * 6, device=DEV), 1).view(4, 20, 12, 6)
    m.train()
    y1 = m.forward_with_weights(v, q, a, w)
    m.eval()
    y2 = m.forward_with_weights(v, q, a, w)
    assert torch.equal(y1, y2)
