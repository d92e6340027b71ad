"""The oracle (oracle/cti_oracle.py) against the reference's own outputs
(tests/golden/cti_golden.pt, made by tests/golden/make_golden.py) and the two
known answers the reference tree holds.  CPU only."""
import os
import sys

import pytest
import torch

from oracle import cti_oracle as O

TOL = dict(rtol=2e-5, atol=2e-6)


def test_kolda_bader_known_answer(golden):
    g = golden["kolda_bader"]
    # textbook result of X x_1 U (Kolda & Bader 2009, sec. 2.5) -- src/Tensor.py:31-32
    expect = torch.tensor([[[22., 130.], [49., 157.], [76., 184.], [103., 211.]],
                           [[28., 172.], [64., 208.], [100., 244.], [136., 280.]]])
    assert torch.equal(g["Y"][0, :, :, :, 0], expect)
    y = O.mode_product3(g["X"][None, :, :, :, None, None].permute(0, 1, 2, 3, 5, 4), g["U1"],
                        torch.eye(4)[None], torch.eye(2)[None])
    assert torch.equal(y[0, :, :, :, 0], expect)


def test_grad_check_known_answer():
    # tools/grad_check.py:8-26 -> q.grad = [1.0136 1.9155 3.0709]
    q = torch.tensor([[1., 2., 3.]], requires_grad=True)
    v = torch.tensor([[[2., 1., 3.], [3., 2., 1.], [1., 2., 3.]]])
    a = torch.softmax((q.unsqueeze(1) * v).sum(2), 1)
    out = (q * (a.unsqueeze(2) * v).sum(1)).sum(1)
    out.backward()
    assert torch.allclose(q.grad, torch.tensor([[1.0136, 1.9155, 3.0709]]), atol=5e-5)
    # closed-form softmax backward used by the CUDA kernel: dlogit = p * (dp - sum p dp)
    p = a.detach()[0]
    dp = (v[0] * q.detach()[0]).sum(1)
    dlog = p * (dp - (p * dp).sum())
    dq = (p[:, None] * v[0]).sum(0) + (dlog[:, None] * v[0]).sum(0)
    assert torch.allclose(dq, torch.tensor([1.0136, 1.9155, 3.0709]), atol=5e-5)


@pytest.mark.parametrize("name", ["fc_relu", "fc_lin", "fc_nodrop"])
def test_fcnet(golden, name):
    g = golden[name]
    p = {k: v.clone().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["x"].clone().requires_grad_(True)
    y = O.fcnet(x, p, "", act=g["act"], dropout=g["dropout"])
    assert torch.allclose(y, g["y"], **TOL)
    (y * g["cot"]).sum().backward()
    assert torch.allclose(x.grad, g["dx"], **TOL)
    for k, gr in g["grads"].items():
        assert torch.allclose(p[k].grad, gr, rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("R,d,G", [(2, 4, 1), (3, 4, 2), (2, 4, 3), (2, 16, 2)])
def test_teff_is_mode_product_with_identities(R, d, G):
    torch.manual_seed(0)
    tg = torch.randn(1, R, d, d, d, G, 1)
    te = O.teff_from_tg(tg)
    eye = torch.eye(d)[None]
    for r in range(R):
        assert torch.equal(O.mode_product3(tg[:, r], eye, eye, eye)[0], te[r])
    idx = O.teff_index_map(R, d, G)
    assert torch.equal(tg.reshape(-1)[idx.reshape(-1)].view(te.shape), te)
    assert sorted(idx.reshape(-1).tolist()) == list(range(tg.numel()))   # a permutation
    if G == 1:
        assert torch.equal(te, tg[0, ..., 0])


@pytest.mark.parametrize("name", ["tri_small_g2", "tri_small_g3", "tri_d16"])
def test_tri_attention_and_pool(golden, name):
    g = golden[name]
    G = g["cfg"]["G"]
    att_p = {"TriAtt." + k[len("TriAtt."):]: v.clone().requires_grad_(True) for k, v in g["att_sd"].items()}
    pool_p = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in g["pool_sd"]]
    q = g["q"].clone().requires_grad_(True)
    a = g["a"].clone().requires_grad_(True)
    p, logits = O.tri_attention(g["v"], q, a, att_p)
    assert torch.equal(torch.isinf(logits), torch.isinf(g["logits"]))
    fin = torch.isfinite(g["logits"])
    assert torch.allclose(logits[fin], g["logits"][fin], rtol=1e-4, atol=1e-4)
    assert torch.allclose(p, g["p"], rtol=1e-4, atol=1e-6)
    closed = O.tcnet_logits_closed(g["v"], q, a, att_p, "TriAtt.")
    assert torch.allclose(closed[fin], g["logits"][fin], rtol=1e-4, atol=1e-4)
    pooled = [O.tcnet_pool(g["v"], q, a, p[:, :, :, :, i], pool_p[i]) for i in range(G)]
    for o, ref in zip(pooled, g["pooled"]):
        assert torch.allclose(o, ref, rtol=1e-4, atol=1e-5)
    sum((o * c).sum() for o, c in zip(pooled, g["cot"])).backward()
    assert torch.allclose(q.grad, g["dq"], rtol=1e-3, atol=1e-5)
    assert torch.allclose(a.grad, g["da"], rtol=1e-3, atol=1e-5)
    for k, gr in g["att_grads"].items():
        assert torch.allclose(att_p[k].grad, gr, rtol=1e-3, atol=1e-5), k
    for i in range(G):
        for k, gr in g["pool_grads"][i].items():
            assert torch.allclose(pool_p[i][k].grad, gr, rtol=1e-3, atol=1e-5), k


@pytest.mark.parametrize("name", ["bi_small", "bi_c128"])
def test_bi_attention_and_pool(golden, name):
    g = golden[name]
    G = g["cfg"]["G"]
    att_p = {k: v.clone().requires_grad_(True) for k, v in g["att_sd"].items()}
    pool_p = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in g["pool_sd"]]
    q = g["q"].clone().requires_grad_(True)
    p, logits = O.bi_attention(g["v"], q, att_p)
    assert torch.equal(torch.isinf(logits), torch.isinf(g["logits"]))
    fin = torch.isfinite(g["logits"])
    assert torch.allclose(logits[fin], g["logits"][fin], rtol=1e-4, atol=1e-4)
    assert torch.allclose(p, g["p"], rtol=1e-4, atol=1e-6)
    pooled = [O.bcnet_pool(g["v"], q, p[:, i], pool_p[i]) for i in range(G)]
    for o, ref in zip(pooled, g["pooled"]):
        assert torch.allclose(o, ref, rtol=1e-4, atol=1e-5)
    sum((o * c).sum() for o, c in zip(pooled, g["cot"])).backward()
    assert torch.allclose(q.grad, g["dq"], rtol=1e-3, atol=1e-5)
    for k, gr in g["att_grads"].items():
        assert torch.allclose(att_p[k].grad, gr, rtol=1e-3, atol=1e-5), k


def test_distillation_loss_matches_torch_formula():
    torch.manual_seed(3)
    x, t = torch.randn(6, 11), torch.randn(6, 11)
    y = (torch.rand(6, 11) > 0.8).float()
    T, alpha = 5.0, 0.005
    ref = torch.nn.KLDivLoss(reduction="none")(torch.log_softmax(x / T, 1), torch.softmax(t / T, 1)).sum(1).mean() \
        * (alpha * T * T) + torch.nn.BCEWithLogitsLoss(reduction="sum")(x, y) / 6 * (1 - alpha)
    assert torch.allclose(O.distillation_loss(x, t, y, T, alpha), ref, rtol=1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the build container")
def test_oracle_vs_live_reference_real_dims():
    """Full-size dims (rank 32, d 16, h_mm 512) against the imported reference, small batch."""
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from src.attention import TriAttention
    torch.manual_seed(1204)
    m = TriAttention(2048, 1024, 1024, 512, 1, 32, 2, 1).eval()
    v, q, a = O.synthetic_inputs(2, 12, 12, 6, seed=5, min_boxes=6)
    with torch.no_grad():
        p_ref, l_ref = m(v, q, a)
        p, l = O.tri_attention(v, q, a, dict(m.state_dict()), "TriAtt.")
    fin = torch.isfinite(l_ref)
    assert torch.equal(fin, torch.isfinite(l))
    scale = l_ref[fin].abs().max()
    assert (l[fin] - l_ref[fin]).abs().max() <= 1e-5 * scale
    assert torch.allclose(p, p_ref, rtol=1e-3, atol=1e-7)


# --------------------------------------------------------------------------- #
# components right after the hot path (tests/golden/next_golden.pt, made by tests/golden/make_golden_next.py)
# --------------------------------------------------------------------------- #
NEXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_golden.pt")


@pytest.mark.parametrize("name", ["clf_mc", "clf_ffoe"])
def test_simple_classifier_matches_reference(name):
    g = torch.load(NEXT)[name]
    params = {"classifier." + k: v.clone().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["x"].clone().requires_grad_(True)
    y = O.simple_classifier(x, params)
    assert (y - g["y"]).abs().max() < 1e-5
    (y * g["cot"]).sum().backward()
    assert (x.grad - g["dx"]).abs().max() < 1e-5
    for k, ref in g["grads"].items():
        assert (params["classifier." + k].grad - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item()), k


def test_trainer_update_matches_reference_clip_and_adamax():
    t = torch.load(NEXT)["trainer"]
    params = [p.clone() for p in t["p0"]]
    m = [torch.zeros_like(p) for p in params]
    u = [torch.zeros_like(p) for p in params]
    for i, st in enumerate(t["steps"]):
        norm = O.trainer_update(params, st["grads"], m, u, i + 1, t["lr"], st["denom"], t["clip_norm"], *t["betas"],
                                t["eps"])
        assert abs(norm - st["norm"]) <= 1e-5 * max(1.0, st["norm"])
        for p, ref in zip(params, st["params"]):
            assert (p - ref).abs().max() < 1e-6


@pytest.mark.parametrize("name", ["gru_small", "gru_300"])
def test_gru_matches_reference_question_embedding(name):
    g = torch.load(NEXT)[name]
    params = {k: v.clone().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["x"].clone().requires_grad_(True)
    y = O.gru_forward_all(x, params)
    assert (y - g["y"]).abs().max() < 1e-5
    assert (y[:, -1] - g["last"]).abs().max() < 1e-5
    (y * g["cot"]).sum().backward()
    assert (x.grad - g["dx"]).abs().max() < 1e-5
    for k, ref in g["grads"].items():
        assert (params[k].grad - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item()), k
