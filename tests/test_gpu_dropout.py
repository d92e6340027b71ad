"""Training-mode dropout on the B200: keep-rate and scaling statistics,
determinism under a fixed (seed, site), forward / backward mask consistency, and p = 0 train == eval."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from cti_b200 import functions as F_, kernels as K_  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("p", [0.2, 0.5])
def test_cast_dropout_statistics_and_mask_regeneration(p):
    x = torch.randn(2000, 512, device=DEV) + 3.0                  # no zeros
    site = (p, 1204, 7)
    y, mask = K_.cast_rows_dropout(x, site, want_mask=True)
    keep = y != 0
    rate = keep.float().mean().item()
    assert abs(rate - (1 - p)) < 5e-3
    assert torch.equal(y[keep], (x[keep] / (1 - p)).to(torch.bfloat16))
    assert not mask.any()
    y2, _ = K_.cast_rows_dropout(x, site)
    assert torch.equal(y, y2)                                     # deterministic
    y3, _ = K_.cast_rows_dropout(x, (p, 1204, 8))
    assert (y3 != 0).ne(keep).float().mean().item() > 0.2         # another site, another mask
    # backward regenerates the same mask (fp32, in place) and so does the bf16 variant
    ones = torch.ones_like(x)
    K_.dropout_f32_(ones, site)
    assert torch.equal(ones != 0, keep)
    assert torch.allclose(ones[keep], torch.full_like(ones[keep], 1 / (1 - p)))
    yb = K_.dropout_bf16(x.to(torch.bfloat16), site)
    assert torch.equal(yb != 0, keep)
    # per-row and per-column keep rates are unbiased too
    assert (keep.float().mean(0) - (1 - p)).abs().max().item() < 0.06
    assert (keep.float().mean(1) - (1 - p)).abs().max().item() < 0.12


def _seeded(fn, seed=11):
    torch.manual_seed(seed)
    F_._DROP_SITES[0] = 0
    return fn()


def test_fcnet_dropout_forward_backward_use_the_same_mask():
    """For the activation-free FCNet (q_prj / a_prj, reference src/MC/base_model.py:204-205) y is linear in x for a
    fixed mask, so <dL/dx, x'> must equal <y(x'), c> when x' is pushed through the same mask."""
    torch.manual_seed(0)
    m = cti_b200.FCNet([1024, 1024], '', .2).to(DEV).train()
    x = torch.randn(64, 1, 1024, device=DEV, requires_grad=True)
    xp = torch.randn(64, 1, 1024, device=DEV)
    c = torch.randn(64, 1, 1024, device=DEV)
    y = _seeded(lambda: m(x))
    (y * c).sum().backward()
    with torch.no_grad():
        y0 = _seeded(lambda: m(torch.zeros_like(xp)))            # bias term
        yp = _seeded(lambda: m(xp))
    lhs = (x.grad * xp).sum().item()
    rhs = ((yp - y0) * c).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * abs(rhs) + 1e-2
    assert (x.grad == 0).float().mean().item() == pytest.approx(0.2, abs=0.02)
    y_again = _seeded(lambda: m(x))
    assert torch.equal(y, y_again)
    m.eval()
    assert not torch.equal(m(x), y)


def test_train_mode_tri_attention_and_pool_run_and_are_deterministic():
    B, K, Q, A, G = 6, 50, 12, 6, 2
    torch.manual_seed(1204)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1).to(DEV).train()
    pool = cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2).to(DEV).train()
    v, q, a = [t.to(DEV) for t in O.synthetic_inputs(B, K, Q, A, seed=3)]

    def run():
        qd, ad = q.clone().requires_grad_(True), a.clone().requires_grad_(True)
        for mod in (att, pool):
            for prm in mod.parameters():
                prm.grad = None
        p, logits = att(v, qd, ad)
        out = pool.forward_with_weights(v, qd, ad, p[:, :, :, :, 0])
        out.sum().backward()
        return p.detach(), out.detach(), qd.grad.clone(), att.TriAtt.T_g.grad.clone()

    p1, o1, g1, t1 = _seeded(run)
    p2, o2, g2, t2 = _seeded(run)
    assert torch.equal(p1, p2) and torch.equal(o1, o2) and torch.equal(g1, g2)
    assert torch.allclose(t1, t2, rtol=1e-3, atol=1e-6)            # atomics: order of the batch reduction varies
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all() and torch.isfinite(t1).all()
    assert torch.allclose(p1.sum((1, 2, 3)), torch.ones(B, G, device=DEV), atol=1e-4)
    for prm in list(att.parameters()) + [x for n, x in pool.named_parameters() if "tucker" in n]:
        assert prm.grad is not None and torch.isfinite(prm.grad).all()
    p3, o3, _, _ = _seeded(run, seed=12)
    assert not torch.equal(o1, o3)                                 # another seed, other masks
    att.eval(); pool.eval()
    with torch.no_grad():
        pe, _ = att(v, q, a)
    assert not torch.equal(pe, p1)


def test_train_mode_ban_runs():
    B, K, Q, G = 5, 50, 12, 2
    torch.manual_seed(7)
    att = cti_b200.BiAttention(2048, 1024, 1024, G).to(DEV).train()
    pool = cti_b200.BCNet(2048, 1024, 1024, None, k=1).to(DEV).train()
    v, q, _ = [t.to(DEV) if t is not None else None for t in O.synthetic_inputs(B, K, Q, 0, seed=5)]
    qd = q.clone().requires_grad_(True)
    p, logits = _seeded(lambda: att.forward_all(v, qd))
    out = pool.forward_with_weights(v, qd, p[:, 0])
    out.sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(qd.grad).all()
    assert torch.allclose(p.sum((2, 3)), torch.ones(B, G, device=DEV), atol=1e-4)
    for prm in list(att.parameters()) + list(pool.parameters()):
        assert prm.grad is not None and torch.isfinite(prm.grad).all()


def test_dropout_zero_in_train_mode_equals_eval():
    torch.manual_seed(0)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, 2, 1, dropout=[0, 0]).to(DEV)
    v, q, a = [t.to(DEV) for t in O.synthetic_inputs(4, 20, 12, 6, seed=3)]
    with torch.no_grad():
        att.train()
        p1, _ = att(v, q, a)
        att.eval()
        p2, _ = att(v, q, a)
    assert torch.equal(p1, p2)


@pytest.mark.parametrize("M,p", [(300, 0.5), (1000, 0.2), (64, 0.5)])
def test_per_rank_nets_draw_independent_masks_and_match_autograd(M, p):
    """The R per-rank FCNets each own a Dropout in the reference (src/tc.py:29-31).  The fused kernels (rank_proj.cu)
    regenerate the masks in registers; cti_rank_proj_dropout_mask writes the same masks out element by element.  Forward and
    backward are compared with autograd of the same masked computation in fp32."""
    torch.manual_seed(0)
    H, R, d = 512, 32, 16
    site = (p, 99, 5)
    assert K_.rank_proj_fused_ok(H, R)
    y = torch.relu(torch.randn(M, H, device=DEV)).to(torch.bfloat16)
    V = (torch.randn(R * d, H, device=DEV) / H ** 0.5).requires_grad_(True)
    g = V.detach().view(R, -1).norm(dim=1).clone().requires_grad_(True)
    bias = (torch.randn(R * d, device=DEV) * 0.1).requires_grad_(True)
    pk = F_.pack_layer(V, g, R)
    out = F_.rank_proj_fwd(y, pk, bias, site, R)
    keep = K_.rank_proj_dropout_mask(M, H, R, site, DEV).float()                                 # (R, M, H) in {0, 1}
    scale = K_.rank_proj_scale(p)
    p_eff = round(p * 256) / 256
    assert scale == pytest.approx(1 / (1 - p_eff), rel=1e-6) and abs(p_eff - p) < 2e-3
    assert set(keep.unique().tolist()) == {0.0, 1.0}
    assert abs(keep.mean().item() - (1 - p_eff)) < 3e-3
    same = (keep[0] == keep[1]).float().mean().item()
    assert abs(same - (p_eff ** 2 + (1 - p_eff) ** 2)) < 0.02                                   # ranks 0 and 1: independent masks
    assert abs((keep[:, 0] == keep[:, 1]).float().mean().item() - (p_eff ** 2 + (1 - p_eff) ** 2)) < 0.03   # rows too
    masks = keep * scale
    yl = y.float().requires_grad_(True)
    w_eff = (V.view(R, -1) * (g / V.view(R, -1).norm(dim=1))[:, None]).view(R, d, H)
    ref = torch.relu(torch.einsum("rmh,rdh->mrd", yl[None] * masks, w_eff.to(torch.bfloat16).float()) + bias.view(R, d))
    ref = ref.reshape(M, R * d)
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2
    cot = torch.randn(M, R * d, device=DEV)
    (ref * cot).sum().backward()
    dz = (cot * (out.float() > 0)).to(torch.bfloat16)
    dV, dg, dy = F_.rank_proj_bwd(y, dz, V.detach(), g.detach(), pk, site, R)
    rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
    assert dy.dtype == torch.bfloat16                                      # fused dgrad: ReLU mask of y applied, bf16
    assert rel(dy, yl.grad * (y > 0)) < 3e-2
    assert rel(dV, V.grad) < 3e-2
    assert ((dg - g.grad).abs().max() / V.grad.view(R, -1).norm(dim=1).max()).item() < 3e-2
    # same site, same masks: deterministic forward; another site, other masks
    assert torch.equal(out, F_.rank_proj_fwd(y, pk, bias, site, R))
    assert not torch.equal(out, F_.rank_proj_fwd(y, pk, bias, (p, 99, 6), R))


def test_rank_projection_batched_over_modalities_equals_single_calls():
    """One call serves the three modalities (p = .5 image side alone, the two p = .2 sides in one launch): same masks,
    same outputs as three single calls."""
    torch.manual_seed(3)
    H, R = 512, 32
    Ms, ps = (300, 1000, 70), (0.5, 0.2, 0.2)
    ys = [torch.relu(torch.randn(M, H, device=DEV)).to(torch.bfloat16) for M in Ms]
    ws = [(torch.randn(R * 16, H, device=DEV) / H ** 0.5).to(torch.bfloat16) for _ in Ms]
    bs = [torch.randn(R * 16, device=DEV) * 0.1 for _ in Ms]
    drops = [(p, 7, 10 + i) for i, p in enumerate(ps)]
    outs = K_.rank_proj_dropout_fwd(ys, ws, bs, R, drops)
    dzs = [torch.randn(M, R * 16, device=DEV).to(torch.bfloat16) for M in Ms]
    dzts = K_.rank_proj_dropout_dgrad(dzs, ws, ys, R, drops)
    dws = [torch.zeros(R * 16, H, device=DEV) for _ in Ms]
    K_.rank_proj_dropout_wgrad(dzs, ys, dws, R, drops)
    for i in range(3):
        o1 = K_.rank_proj_dropout_fwd([ys[i]], [ws[i]], [bs[i]], R, [drops[i]])[0]
        d1 = K_.rank_proj_dropout_dgrad([dzs[i]], [ws[i]], [ys[i]], R, [drops[i]])[0]
        w1 = torch.zeros(R * 16, H, device=DEV)
        K_.rank_proj_dropout_wgrad([dzs[i]], [ys[i]], [w1], R, [drops[i]])
        assert torch.equal(outs[i], o1) and torch.equal(dzts[i], d1)
        assert ((dws[i] - w1).abs().max() / w1.abs().max()).item() < 1e-5          # fp32 red.add: order varies


def test_expand_path_still_serves_other_shapes():
    """Widths the fused kernels are not built for (H != 512) go through cti_dropout_expand + block-diagonal GEMMs."""
    torch.manual_seed(0)
    M, H, R, d, p = 96, 256, 32, 16, 0.5
    site = (p, 99, 5)
    assert not K_.rank_proj_fused_ok(H, R)
    y = torch.relu(torch.randn(M, H, device=DEV)).to(torch.bfloat16)
    V = (torch.randn(R * d, H, device=DEV) / H ** 0.5)
    g = V.view(R, -1).norm(dim=1).clone()
    bias = torch.randn(R * d, device=DEV) * 0.1
    pk = F_.pack_layer(V, g, R)
    out = F_.rank_proj_fwd(y, pk, bias, site, R)
    masks = torch.stack([K_.dropout_expand(torch.ones_like(y), 1, r, site) for r in range(R)], 0).float()
    w_eff = (V.view(R, -1) * (g / V.view(R, -1).norm(dim=1))[:, None]).view(R, d, H)
    ref = torch.relu(torch.einsum("rmh,rdh->mrd", y.float()[None] * masks, w_eff.to(torch.bfloat16).float()) + bias.view(R, d))
    assert ((out.float() - ref.reshape(M, R * d)).abs().max() / ref.abs().max()).item() < 1e-2


def test_rank_dropout_modes_both_run():
    B, K, Q, A, G = 4, 50, 12, 6, 2
    torch.manual_seed(1204)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, G, 1).to(DEV).train()
    v, q, a = [t.to(DEV) for t in O.synthetic_inputs(B, K, Q, A, seed=3)]
    outs = {}
    for mode in ("independent", "shared"):
        F_.RANK_DROPOUT = mode
        try:
            qd = q.clone().requires_grad_(True)
            for prm in att.parameters():
                prm.grad = None
            p, _ = _seeded(lambda: att(v, qd, a))
            (p * torch.arange(p.numel(), device=DEV).view_as(p).float()).sum().backward()
            assert torch.isfinite(qd.grad).all()
            assert all(prm.grad is not None and torch.isfinite(prm.grad).all() for prm in att.parameters())
            outs[mode] = p.detach().clone()
        finally:
            F_.RANK_DROPOUT = "independent"
    assert not torch.equal(outs["independent"], outs["shared"])
