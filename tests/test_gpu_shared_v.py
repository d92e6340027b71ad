"""Rows that share one image: the multiple-choice trainer clones every image once per
answer candidate on the device (reference src/MC/train.py:75-76).  The drop-ins accept the un-cloned features
(v with B / n samples next to q, a with B rows) and must give what the reference gives on the cloned tensor:

  * forward: bit-identical to the same modules fed the explicit clone (the clones' projections are identical rows),
  * backward: parameter / input gradients equal to the cloned run up to the order of the fp32 sums,
  * and the usual tolerance against the fp32 oracle evaluated on the cloned tensor.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from cti_b200 import kernels as KS  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402
from test_gpu_modules import ABS_TOL, build_cti, check_grads_fp32, cti_forward, maxabs, normrel, rel, run_oracle  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16


def clone_rows(x, n):
    """What src/MC/train.py:75-76 does to v."""
    return x.unsqueeze(1).expand(x.size(0), n, *x.shape[1:]).contiguous().view(x.size(0) * n, *x.shape[1:])


@pytest.mark.parametrize("groups,rep,row", [(7, 4, 64), (300, 3, 50 * 512), (1, 1, 8)])
def test_sum_row_groups_matches_fp32_sum(groups, rep, row):
    g = torch.Generator().manual_seed(groups)
    x = torch.randn(groups * rep, row, generator=g).to(BF16)
    out = KS.sum_row_groups(x.to(DEV), rep, row)
    ref = x.float().view(groups, rep, row).sum(1).to(BF16)
    assert out.shape == (groups, row)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("Bv,rep,K,Q,A", [(6, 4, 50, 12, 6), (3, 2, 37, 9, 3), (40, 4, 50, 12, 6)])
def test_trilinear_kernels_with_shared_v_equal_explicit_clone(Bv, rep, K, Q, A):
    G, R = 2, 32
    B = Bv * rep
    g = torch.Generator().manual_seed(Bv + K)
    vc = torch.relu(torch.randn(Bv * K, R * 16, generator=g)).to(BF16).to(DEV)
    qc = torch.relu(torch.randn(B * Q, R * 16, generator=g)).to(BF16).to(DEV)
    ac = torch.relu(torch.randn(B * A, R * 16, generator=g)).to(BF16).to(DEV)
    tp = (0.1 * torch.randn(R, 16, 16 * G * 16, generator=g)).to(BF16).to(DEV)
    mask = (torch.rand(Bv, K, generator=g) < 0.2).to(torch.uint8).to(DEV)
    vc_cl = clone_rows(vc.view(Bv, K, -1), rep).view(B * K, -1)
    mask_cl = clone_rows(mask, rep)
    lo, n1 = KS.trilinear_fwd(vc, qc, ac, tp, mask, B, K, Q, A, G, R, rep, save_n1=True)
    lo_cl, n1_cl = KS.trilinear_fwd(vc_cl, qc, ac, tp, mask_cl, B, K, Q, A, G, R, save_n1=True)
    assert torch.equal(lo, lo_cl) and torch.equal(n1, n1_cl)
    dl = torch.randn(B, G, K, Q, A, generator=g).to(DEV)
    dl = torch.where(torch.isinf(lo), torch.zeros_like(dl), dl)
    out = KS.trilinear_bwd(vc, qc, ac, tp, dl, B, K, Q, A, G, R, rep, n1=n1)
    out_cl = KS.trilinear_bwd(vc_cl, qc, ac, tp, dl, B, K, Q, A, G, R, n1=n1_cl)
    dzv, dzv_cl = out[0], out_cl[0]
    assert dzv.shape == (Bv * K, R * 16)
    folded = dzv_cl.float().view(Bv, rep, K * R * 16).sum(1).to(BF16).view(Bv * K, R * 16)
    assert torch.equal(dzv, folded)                      # same per-row kernel output, folded in fp32
    assert torch.equal(out[1], out_cl[1]) and torch.equal(out[2], out_cl[2])       # dzq, dza: per-row, deterministic
    for x, y in zip(out[3:], out_cl[3:]):                # bias / core gradients: fp32 atomics, order varies
        assert normrel(x, y.cpu()) < 1e-5


@pytest.mark.parametrize("Bv,rep,K,Q,A,C", [(5, 4, 50, 12, 6, 1024), (4, 2, 36, 7, 0, 512), (37, 4, 50, 12, 6, 1024)])
def test_pool_kernels_with_shared_v_equal_explicit_clone(Bv, rep, K, Q, A, C):
    B = Bv * rep
    g = torch.Generator().manual_seed(Bv * 3 + K)
    v = torch.relu(torch.randn(Bv * K, C, generator=g)).to(BF16).to(DEV)
    q = torch.relu(torch.randn(B * Q, C, generator=g)).to(BF16).to(DEV)
    a = torch.relu(torch.randn(B * A, C, generator=g)).to(BF16).to(DEV) if A > 0 else None
    w = torch.softmax(torch.randn(B, K * Q * max(A, 1), generator=g), 1).view((B, K, Q, A) if A else (B, K, Q)).to(DEV)
    v_cl = clone_rows(v.view(Bv, K, C), rep).view(B * K, C)
    out = KS.tri_pool_fwd(v, q, a, w, w.stride(0), B, K, Q, A, C, rep)
    out_cl = KS.tri_pool_fwd(v_cl, q, a, w, w.stride(0), B, K, Q, A, C)
    assert torch.equal(out, out_cl)
    dout = torch.randn(B, C, generator=g).to(DEV)
    r = KS.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C, rep)
    r_cl = KS.tri_pool_bwd(v_cl, q, a, w, w.stride(0), dout, B, K, Q, A, C)
    folded = r_cl[0].float().view(Bv, rep, K * C).sum(1).to(BF16).view(Bv * K, C)
    assert r[0].shape == (Bv * K, C) and torch.equal(r[0], folded)
    assert torch.equal(r[1], r_cl[1])                                             # dzq
    if A > 0:
        assert torch.equal(r[2], r_cl[2])
    assert torch.equal(r[6], r_cl[6])                                             # dw
    for x, y in zip(r[3:6], r_cl[3:6]):
        if x is not None:
            assert normrel(x, y.cpu()) < 1e-5


def test_cti_hot_path_with_shared_images_equals_cloned_run_and_oracle():
    """MC shape: 3 questions x 4 answer candidates; v handed in once per question."""
    Bq, rep, K, Q, A, G = 3, 4, 50, 12, 6, 2
    B = Bq * rep
    params = O.random_cti_params(glimpse=G, seed=1204)
    v_q, _, _ = O.synthetic_inputs(Bq, K, Q, A, seed=77)
    _, q, a = O.synthetic_inputs(B, K, Q, A, seed=78)
    v_cl = clone_rows(v_q, rep)
    cot = torch.randn(B, 1024, generator=torch.Generator().manual_seed(3))

    def fn(pl, ql, al):
        joint, pp, ll = O.cti_hot_path(v_cl, ql, al, pl, G)
        return (joint * cot).sum(), (joint, pp, ll)
    (joint_ref, p_ref, logits_ref), lv32, g32 = run_oracle(fn, params, [q, a])

    def run(v_in):
        att, pools, prj = build_cti(params, G, DEV)
        qd, ad = q.to(DEV).requires_grad_(True), a.to(DEV).requires_grad_(True)
        joint, p, logits = cti_forward(att, pools, prj, v_in.to(DEV), qd, ad)
        (joint * cot.to(DEV)).sum().backward()
        mods = [("v_att.", att)] + [(f"t_net.{i}.", m) for i, m in enumerate(pools)]
        mods += [(f"q_prj.{i}.", pr[0]) for i, pr in enumerate(prj)] + [(f"a_prj.{i}.", pr[1]) for i, pr in enumerate(prj)]
        grads = {"dq": qd.grad, "da": ad.grad}
        grads.update({pre + k: t.grad for pre, m in mods for k, t in m.named_parameters()})
        return joint, p, logits, grads

    joint_s, p_s, logits_s, g_s = run(v_q)
    joint_c, p_c, logits_c, g_c = run(v_cl)
    assert p_s.shape == (B, K, Q, A, G) and joint_s.shape == (B, 1024)
    assert torch.equal(logits_s, logits_c) and torch.equal(p_s, p_c) and torch.equal(joint_s, joint_c)
    assert all(g_s[name] is not None for name in g_c)
    # bf16 fold of dzv vs fp32 accumulation over the clone rows inside the wgrad: 1e-2 in norm (scalars such as
    # weight_g are cancelling sums and are judged with the whole gradient, as in test_gpu_modules)
    check_grads_fp32(list(g_s.items()), {k: t.cpu() for k, t in g_c.items()}, tol=1e-2)
    # and against the reference arithmetic on the cloned tensor
    inf_ref = torch.isinf(logits_ref)
    assert torch.equal(torch.isinf(logits_s).cpu(), inf_ref)
    assert maxabs(logits_s.cpu()[~inf_ref], logits_ref[~inf_ref]) <= ABS_TOL
    assert maxabs(p_s, p_ref) <= ABS_TOL
    assert rel(joint_s, joint_ref) <= ABS_TOL
    ref = dict(g32, dq=lv32[0], da=lv32[1])
    check_grads_fp32([(k, g) for k, g in g_s.items() if ref.get(k) is not None], ref)


def test_batch_mismatch_raises():
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, 2, 1).to(DEV).eval()
    v, q, a = O.synthetic_inputs(3, 20, 5, 3, seed=1)
    _, q4, a4 = O.synthetic_inputs(4, 20, 5, 3, seed=2)
    with pytest.raises(RuntimeError, match="batch mismatch"):
        att(v.to(DEV), q4.to(DEV), a4.to(DEV))


def test_shared_images_in_train_mode_with_dropout_run_and_give_every_parameter_a_gradient():
    Bq, rep, K, Q, A, G = 2, 4, 30, 12, 6, 2
    B = Bq * rep
    params = O.random_cti_params(glimpse=G, seed=5)
    v_q, _, _ = O.synthetic_inputs(Bq, K, Q, A, seed=11)
    _, q, a = O.synthetic_inputs(B, K, Q, A, seed=12)
    att, pools, prj = build_cti(params, G, DEV)
    mods = [att] + pools + [m for pr in prj for m in pr]
    for m in mods:
        m.train()
    torch.manual_seed(0)
    qd, ad = q.to(DEV).requires_grad_(True), a.to(DEV).requires_grad_(True)
    joint, p, logits = cti_forward(att, pools, prj, v_q.to(DEV), qd, ad)
    assert torch.isfinite(joint).all() and torch.isfinite(p).all()
    assert torch.allclose(p.sum(dim=(1, 2, 3)), torch.ones(B, G, device=DEV), atol=1e-4)
    joint.sum().backward()
    assert torch.isfinite(qd.grad).all() and torch.isfinite(ad.grad).all()
    for m in mods:
        for name, t in m.named_parameters():
            assert t.grad is not None and torch.isfinite(t.grad).all(), name
