"""Kernel-level parity on the B200: every C-ABI entry point against the oracle / a plain torch fp32
reference of the same op on the same seeded inputs.  Inputs to the tensor-core kernels are bf16
(that is the kernels' arithmetic type); the reference is evaluated in fp32 on the bf16-rounded
operands, so the tolerances below only cover fp32 accumulation order and the bf16 rounding of
on-chip intermediates / outputs."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

import cti_b200  # noqa: E402
from cti_b200 import kernels as K_  # noqa: E402
from oracle import cti_oracle as O  # noqa: E402

DEV = "cuda"


def bf(t):
    return t.to(torch.bfloat16)


def rel_err(x, ref):
    return ((x.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20)).item()


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 136), (1000, 512, 2048), (50, 16, 512), (2600, 1024, 1024),
                                   (129, 264, 72)])
@pytest.mark.parametrize("tile_n", [0, 128, 256, 512])          # 512: CTA pairs (cta_group::2, 256 x 256 tiles)
def test_gemm_kmajor(M, N, K, tile_n):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    a = bf(torch.randn(M, K, device=DEV, generator=g))
    b = bf(torch.randn(N, K, device=DEV, generator=g) / K ** 0.5)
    bias = torch.randn(N, device=DEV, generator=g)
    ref = a.float() @ b.float().t() + bias
    ob, of = K_.gemm(a, b, M, N, K, bias=bias, out_bf16=True, out_f32=True, tile_n=tile_n)
    assert rel_err(of, ref) < 2e-5
    assert rel_err(ob, ref) < 5e-3
    ob2, _ = K_.gemm(a, b, M, N, K, bias=bias, relu=True, tile_n=tile_n)
    assert rel_err(ob2, torch.relu(ref)) < 5e-3
    aux = bf(torch.randn(M, N, device=DEV, generator=g))
    _, of3 = K_.gemm(a, b, M, N, K, relu_aux=aux, out_bf16=False, out_f32=True, alpha=0.5, tile_n=tile_n)
    ref3 = 0.5 * (a.float() @ b.float().t()) * (aux.float() > 0)
    assert rel_err(of3, ref3) < 2e-5


@pytest.mark.parametrize("M,N,K", [(256, 512, 128), (1000, 2048, 512), (333, 136, 72), (600, 1024, 16)])
def test_gemm_dgrad_layout(M, N, K):
    """dx[M, N] = dz[M, K] . W[K, N]: B operand MN-major (stored [K][N])."""
    g = torch.Generator(device=DEV).manual_seed(3)
    dz = bf(torch.randn(M, K, device=DEV, generator=g))
    w = bf(torch.randn(K, N, device=DEV, generator=g) / K ** 0.5)
    ref = dz.float() @ w.float()
    for tile_n in (128, 256, 512):
        _, of = K_.gemm(dz, w, M, N, K, b_mn=True, out_bf16=False, out_f32=True, tile_n=tile_n)
        assert rel_err(of, ref) < 2e-5


@pytest.mark.parametrize("rows,N,Kin,splits", [(640, 128, 256, 1), (5000, 512, 2048, 4), (1237, 16, 512, 3),
                                               (12800, 512, 512, 9), (777, 136, 72, 2)])
def test_gemm_wgrad_layout_splitk(rows, N, Kin, splits):
    """dW[N, Kin] = dz^T x: both operands MN-major, split over the reduction with fp32 atomics."""
    g = torch.Generator(device=DEV).manual_seed(5)
    dz = bf(torch.randn(rows, N, device=DEV, generator=g))
    x = bf(torch.randn(rows, Kin, device=DEV, generator=g))
    ref = dz.float().t() @ x.float()
    for tile_n in (128, 256, 512):
        acc = torch.zeros(N, Kin, device=DEV)
        K_.gemm(dz, x, N, Kin, rows, a_mn=True, b_mn=True, accum_f32=acc, k_splits=splits, tile_n=tile_n)
        assert rel_err(acc, ref) < 1e-4


@pytest.mark.parametrize("Mq,Ma,N,K", [(1536, 768, 1024, 1024), (12288, 6144, 512, 512), (300, 77, 136, 72), (96, 2000, 512, 256),
                                       (12288, 6144, 1024, 1024)])          # the last one runs as CTA pairs
def test_gemm_batch_pairs_independent_problems(Mq, Ma, N, K):
    """Inside gemm_batch() two GEMMs with the same operand layouts leave in ONE launch (cti_gemm_bf16_pair); results are
    bit-identical to the separate launches (forward, dgrad) or equal up to the split-K summation order (wgrad)."""
    g = torch.Generator(device=DEV).manual_seed(Mq + N)
    mk = lambda *sh: bf(torch.randn(*sh, device=DEV, generator=g))
    xq, xa, wq, wa = mk(Mq, K), mk(Ma, K), mk(N, K) / K ** 0.5, mk(N, K) / K ** 0.5
    bq, ba = torch.randn(N, device=DEV, generator=g), torch.randn(N, device=DEV, generator=g)
    want_q, _ = K_.gemm(xq, wq, Mq, N, K, bias=bq, relu=True)
    want_a, _ = K_.gemm(xa, wa, Ma, N, K, bias=ba, relu=True)
    n0 = K_.STATS.launches
    with K_.gemm_batch():
        yq, _ = K_.gemm(xq, wq, Mq, N, K, bias=bq, relu=True)
        ya, _ = K_.gemm(xa, wa, Ma, N, K, bias=ba, relu=True)
    assert K_.STATS.launches - n0 == 1
    assert torch.equal(yq, want_q) and torch.equal(ya, want_a)
    assert rel_err(yq, torch.relu(xq.float() @ wq.float().t() + bq)) < 5e-3
    # dgrad pair (B MN-major, ReLU-mask aux on one side, fp32 output on the other) + wgrad pair (split-K) in one batch:
    # four queued GEMMs, two launches
    dzq, dza = mk(Mq, N), mk(Ma, N)
    aux = mk(Mq, K)
    accq, acca = torch.zeros(N, K, device=DEV), torch.zeros(N, K, device=DEV)
    n0 = K_.STATS.launches
    with K_.gemm_batch():
        K_.gemm(dzq, xq, N, K, Mq, a_mn=True, b_mn=True, accum_f32=accq, k_splits=3)
        dxq, _ = K_.gemm(dzq, wq, Mq, K, N, b_mn=True, relu_aux=aux)
        K_.gemm(dza, xa, N, K, Ma, a_mn=True, b_mn=True, accum_f32=acca, k_splits=2)
        _, dxa = K_.gemm(dza, wa, Ma, K, N, b_mn=True, out_bf16=False, out_f32=True)
    assert K_.STATS.launches - n0 == 2
    assert rel_err(accq, dzq.float().t() @ xq.float()) < 1e-4 and rel_err(acca, dza.float().t() @ xa.float()) < 1e-4
    assert rel_err(dxq, (dzq.float() @ wq.float()) * (aux.float() > 0)) < 5e-3
    assert rel_err(dxa, dza.float() @ wa.float()) < 2e-5
    # a GEMM that reads a queued GEMM's output flushes the queue first (no pairing of dependent problems)
    n0 = K_.STATS.launches
    with K_.gemm_batch():
        h, _ = K_.gemm(xq, wq, Mq, N, K, bias=bq, relu=True)
        w2 = mk(N, N) / N ** 0.5
        y2, _ = K_.gemm(h, w2, Mq, N, N)
    assert K_.STATS.launches - n0 == 2
    assert rel_err(y2, want_q.float() @ w2.float().t()) < 5e-3


def test_gemm_rejects_bad_arguments():
    a = bf(torch.randn(8, 12, device=DEV))          # pitch 24 B: not a multiple of 16
    b = bf(torch.randn(8, 12, device=DEV))
    with pytest.raises(RuntimeError, match="argument error"):
        K_.gemm(a, b, 8, 8, 12)


# --------------------------------------------------------------------------- #
def test_cast_rows_mask():
    x = torch.randn(1000, 2048, device=DEV)
    x[::7] = 0
    x[3, 5] = 1e-30
    out, mask = K_.cast_rows(x, want_mask=True)
    assert torch.equal(out, bf(x))
    assert torch.equal(mask.bool(), x.abs().sum(1) == 0)
    out2, _ = K_.cast_rows(torch.randn(17, 36, device=DEV))     # scalar path (cols % 8 != 0)
    assert out2.shape == (17, 36)


@pytest.mark.parametrize("groups,rows,cols", [(1, 512, 2048), (32, 16, 512), (1, 40, 24), (4, 16, 64)])
def test_weight_norm_pack_and_grad(groups, rows, cols):
    g_ = torch.Generator(device=DEV).manual_seed(9)
    v = torch.randn(groups * rows, cols, device=DEV, generator=g_, requires_grad=True)
    g = (torch.rand(groups, device=DEV, generator=g_) + 0.5).requires_grad_(True)
    w_ref = (v.view(groups, -1) * (g / v.view(groups, -1).norm(dim=1))[:, None]).view_as(v)
    w, sumsq = K_.wn_pack(v.detach(), g.detach(), groups)
    assert rel_err(sumsq, v.detach().view(groups, -1).pow(2).sum(1)) < 1e-5
    assert rel_err(w, w_ref) < 5e-3
    dw = torch.randn(groups * rows, cols, device=DEV, generator=g_)
    (w_ref * dw).sum().backward()
    dv, dg = K_.wn_grad(dw, v.detach(), g.detach(), sumsq, groups)
    assert rel_err(dv, v.grad) < 1e-4
    assert rel_err(dg, g.grad) < 1e-4


def test_act_bwd_bias():
    dy = torch.randn(777, 512, device=DEV)
    y = bf(torch.randn(777, 512, device=DEV))
    db = torch.zeros(512, device=DEV)
    dz = K_.act_bwd_bias(dy, y, True, db)
    ref = dy * (y.float() > 0)
    assert torch.equal(dz, bf(ref))
    assert rel_err(db, ref.sum(0)) < 1e-5
    db2 = torch.zeros(512, device=DEV)
    K_.act_bwd_bias(bf(dy), None, False, db2)
    assert rel_err(db2, bf(dy).float().sum(0)) < 1e-5


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("rows,length", [(64, 3600), (33, 1800), (10, 600), (5, 37), (3, 5000)])
def test_masked_softmax_fwd_bwd(rows, length):
    x = torch.randn(rows, length, device=DEV) * 3
    x[:, : length // 5] = float("-inf")
    x.requires_grad_(True)
    ref = torch.softmax(x, 1)
    p = K_.softmax_fwd(x.detach(), rows, length)
    assert (p - ref).abs().max() < 1e-6
    assert torch.all(p[:, : length // 5] == 0)
    dp = torch.randn(rows, length, device=DEV)
    ref.backward(dp)
    dl = K_.softmax_bwd(p, dp, 0, length, 1, 1, rows, length)
    assert (dl - torch.nan_to_num(x.grad)).abs().max() < 1e-6


def test_softmax_known_answer_grad_check():
    """tools/grad_check.py:8-26 of the reference: q.grad = [1.0136 1.9155 3.0709]."""
    q = torch.tensor([1., 2., 3.], device=DEV)
    v = torch.tensor([[2., 1., 3.], [3., 2., 1.], [1., 2., 3.]], device=DEV)
    logits = (q[None] * v).sum(1)[None].contiguous()
    p = K_.softmax_fwd(logits, 1, 3)
    dp = (v * q).sum(1)[None].contiguous()
    dl = K_.softmax_bwd(p, dp, 0, 3, 1, 1, 1, 3)
    dq = (p[0][:, None] * v).sum(0) + (dl[0][:, None] * v).sum(0)
    assert torch.allclose(dq.cpu(), torch.tensor([1.0136, 1.9155, 3.0709]), atol=5e-5)


def test_softmax_bwd_strided_gradient():
    B, G, L = 6, 2, 600
    p = torch.softmax(torch.randn(B, G, L, device=DEV), 2)
    dp_bLG = torch.randn(B, L, G, device=DEV)                    # the layout autograd hands back for (B,K,Q,A,G)
    dl = K_.softmax_bwd(p, dp_bLG, L * G, 1, G, B, G, L)
    dp = dp_bLG.permute(0, 2, 1)
    ref = p * (dp - (p * dp).sum(2, keepdim=True))
    assert (dl - ref).abs().max() < 1e-6


# --------------------------------------------------------------------------- #
def tri_inputs(B, K, Q, A, G, R, seed):
    g = torch.Generator().manual_seed(seed)
    vc = bf(torch.relu(torch.randn(B, K, R, 16, generator=g)) * 0.3).float()
    qc = bf(torch.relu(torch.randn(B, Q, R, 16, generator=g)) * 0.3).float()
    ac = bf(torch.relu(torch.randn(B, A, R, 16, generator=g)) * 0.3).float()
    tg = bf(torch.randn(1, R, 16, 16, 16, G, 1, generator=g)).float()
    return vc, qc, ac, tg


@pytest.mark.parametrize("B,K,Q,A,G,R", [(3, 50, 12, 6, 2, 32), (2, 10, 12, 6, 2, 4), (5, 36, 12, 3, 2, 32),
                                         (2, 50, 12, 4, 3, 8), (150, 17, 5, 2, 2, 2), (1, 1, 1, 1, 2, 1),
                                         (151, 33, 16, 5, 2, 8), (7, 64, 1, 1, 2, 12)])
def test_trilinear_logits_fwd_bwd(B, K, Q, A, G, R):
    from cti_b200 import functions as F_
    vc, qc, ac, tg = tri_inputs(B, K, Q, A, G, R, B * 100 + K)
    vc_, qc_, ac_, tg_ = [t.clone().requires_grad_(True) for t in (vc, qc, ac, tg)]
    ref = O.trilinear_closed(vc_, qc_, ac_, O.teff_from_tg(tg_))                    # (B,K,Q,A,G)
    mask = torch.zeros(B, K, dtype=torch.uint8)
    mask[:, K - K // 4:] = 1
    dev = lambda t: t.to(DEV)
    tpack = F_.pack_core(dev(tg))
    vcd, qcd, acd = [bf(dev(t)).reshape(t.shape[0] * t.shape[1], R * 16).contiguous() for t in (vc, qc, ac)]
    out, n1 = K_.trilinear_fwd(vcd, qcd, acd, tpack, dev(mask).reshape(-1).contiguous(), B, K, Q, A, G, R, save_n1=True)
    assert (n1 is not None) == (G == 2 and A <= 6 and K <= 64 and R % 4 == 0)      # the tcgen05 path saves its N1 tiles
    out = out.permute(0, 2, 3, 4, 1).cpu()
    keep = mask == 0
    scale = ref.abs().max().item()
    assert torch.isinf(out[~keep]).all() and (out[~keep] < 0).all()
    assert (out[keep] - ref[keep]).abs().max().item() <= 1e-2 * scale
    # backward
    gen = torch.Generator().manual_seed(1)
    dl = torch.randn(B, G, K, Q, A, generator=gen) * keep[:, None, :, None, None]
    ref.backward(dl.permute(0, 2, 3, 4, 1))
    def chk(got, grad, act, name):
        want = (grad * (act > 0)).reshape(got.shape)
        e = (got.float().cpu() - want).abs().max().item() / want.abs().max().clamp_min(1e-12).item()
        assert e < 2e-2, (name, e)
        return want
    # with the saved N1 tiles (tcgen05 backward) and without (generic kernel)
    for saved in ([n1, None] if n1 is not None else [None]):
        dzv, dzq, dza, dbv, dbq, dba, dtp = K_.trilinear_bwd(vcd, qcd, acd, tpack, dev(dl).contiguous(), B, K, Q, A, G, R,
                                                             n1=saved)
        wv = chk(dzv, vc_.grad, vc, "dzv")
        wq = chk(dzq, qc_.grad, qc, "dzq")
        wa = chk(dza, ac_.grad, ac, "dza")
        for got, want, name in ((dbv, wv, "dbv"), (dbq, wq, "dbq"), (dba, wa, "dba")):
            assert rel_err(got.cpu(), want.sum(0)) < 2e-2, name
        dT = F_.unpack_core_grad(dtp, dev(tg)).cpu()
        assert rel_err(dT, tg_.grad) < 2e-2


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("B,K,Q,A,C", [(3, 50, 12, 6, 1024), (4, 36, 12, 3, 1024), (2, 10, 12, 6, 128), (160, 20, 7, 0, 256),
                                       (3, 50, 12, 0, 1024), (2, 7, 3, 2, 128)])
def test_pool_fwd_bwd(B, K, Q, A, C):
    g = torch.Generator().manual_seed(C + A)
    mk = lambda n: bf(torch.relu(torch.randn(B, n, C, generator=g))).float()
    v, q = mk(K), mk(Q)
    a = mk(A) if A > 0 else None
    An = max(A, 1)
    # attention weights live in a (B, G=2, K*Q*A) buffer: glimpse 1 is a strided-per-batch slice
    att = torch.softmax(torch.randn(B, 2, K * Q * An, generator=g), 2)
    w = att[:, 1].reshape(B, K, Q, An) if A > 0 else att[:, 1].reshape(B, K, Q)
    leaves = [t.clone().requires_grad_(True) for t in ([v, q, a, w] if A > 0 else [v, q, w])]
    if A > 0:
        ref = O.trilinear_pool(leaves[0], leaves[1], leaves[2], leaves[3])
    else:
        ref = torch.einsum("bkc,bkq,bqc->bc", leaves[0], leaves[2], leaves[1])
    dev = lambda t: t.to(DEV)
    vd, qd = bf(dev(v)).reshape(B * K, C).contiguous(), bf(dev(q)).reshape(B * Q, C).contiguous()
    ad = bf(dev(a)).reshape(B * A, C).contiguous() if A > 0 else None
    attd = dev(att)
    wd = attd[:, 1]
    out = K_.tri_pool_fwd(vd, qd, ad, wd, wd.stride(0), B, K, Q, A, C)
    assert rel_err(out.cpu(), ref) < 1e-2
    dout = torch.randn(B, C, generator=g)
    ref.backward(dout)
    dzv, dzq, dza, dbv, dbq, dba, dw = K_.tri_pool_bwd(vd, qd, ad, wd, wd.stride(0), dev(dout), B, K, Q, A, C)
    wv = (leaves[0].grad * (v > 0)).reshape(B * K, C)
    wq = (leaves[1].grad * (q > 0)).reshape(B * Q, C)
    assert rel_err(dzv.cpu(), wv) < 2e-2 and rel_err(dzq.cpu(), wq) < 2e-2
    assert rel_err(dbv.cpu(), wv.sum(0)) < 2e-2 and rel_err(dbq.cpu(), wq.sum(0)) < 2e-2
    if A > 0:
        wa = (leaves[2].grad * (a > 0)).reshape(B * A, C)
        assert rel_err(dza.cpu(), wa) < 2e-2 and rel_err(dba.cpu(), wa.sum(0)) < 2e-2
    assert rel_err(dw.cpu().reshape(B, -1), leaves[-1].grad.reshape(B, -1)) < 2e-2


# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("B,K,Q,G,C", [(3, 50, 12, 2, 3072), (2, 36, 14, 4, 384), (150, 10, 3, 2, 128)])
def test_bilinear_logits_fwd_bwd(B, K, Q, G, C):
    g = torch.Generator().manual_seed(C)
    vb = bf(torch.relu(torch.randn(B, K, C, generator=g)) * 0.2).float()
    qb = bf(torch.relu(torch.randn(B, Q, C, generator=g)) * 0.2).float()
    h = torch.randn(1, G, 1, C, generator=g) * 0.3
    hb = torch.randn(1, G, 1, 1, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (vb, qb, h, hb)]
    ref = O.bilinear_closed(*leaves)                                   # (B,G,K,Q)
    mask = torch.zeros(B, K, dtype=torch.uint8)
    mask[:, K - 2:] = 1
    dev = lambda t: t.to(DEV)
    vd, qd = bf(dev(vb)).reshape(B * K, C).contiguous(), bf(dev(qb)).reshape(B * Q, C).contiguous()
    hd, hbd = dev(h).reshape(G, C).contiguous(), dev(hb).reshape(G).contiguous()
    out = K_.bilinear_fwd(vd, qd, hd, hbd, dev(mask).reshape(-1).contiguous(), B, K, Q, G, C).cpu()
    keep = (mask == 0)[:, None, :, None].expand_as(ref)
    assert torch.isinf(out[~keep]).all()
    scale = ref.abs().max().item()
    assert (out[keep] - ref[keep]).abs().max().item() < 1e-2 * scale
    dl = torch.randn(B, G, K, Q, generator=g) * keep
    ref.backward(dl)
    dzv, dzq, dbv, dbq, dh, dhb = K_.bilinear_bwd(vd, qd, hd, dev(dl).contiguous(), B, K, Q, G, C)
    wv = (leaves[0].grad * (vb > 0)).reshape(B * K, C)
    wq = (leaves[1].grad * (qb > 0)).reshape(B * Q, C)
    assert rel_err(dzv.cpu(), wv) < 2e-2 and rel_err(dzq.cpu(), wq) < 2e-2
    assert rel_err(dbv.cpu(), wv.sum(0)) < 2e-2 and rel_err(dbq.cpu(), wq.sum(0)) < 2e-2
    assert rel_err(dh.cpu(), leaves[2].grad.reshape(G, C)) < 2e-2
    assert rel_err(dhb.cpu(), leaves[3].grad.reshape(G)) < 1e-3


# --------------------------------------------------------------------------- #
def test_pool_kernels_are_bit_repeatable_at_bench_size():
    """racecheck reports the cp.async.bulk write of the per-stage `dout` chunk against the epilogue's read of it in
    pool_kernel (tc_tiles.cuh bulk_load_1d vs ld_shared_f32); the two are ordered by the stage's mbarrier
    (complete_tx -> try_wait on the full barrier, read-side arrive on the empty barrier), which the tool does not
    model.  A real race would make some of these outputs depend on timing: 100 back-to-back launches at the bench size
    (1024 rows, the grid wraps 7 times) on two streams' worth of load must be bit-identical.  dbv / dbq / dba are
    excluded: they are cross-CTA float atomics, whose order is not fixed by design."""
    B, K, Q, A, C = 1024, 50, 12, 6, 1024
    g = torch.Generator(device=DEV).manual_seed(11)
    mk = lambda n: torch.relu(torch.randn(B * n, C, generator=g, device=DEV)).to(torch.bfloat16)
    v, q, a = mk(K), mk(Q), mk(A)
    att = torch.softmax(torch.randn(B, 2, K * Q * A, generator=g, device=DEV), 2)
    w = att[:, 1]
    dout = torch.randn(B, C, generator=g, device=DEV)
    out0 = K_.tri_pool_fwd(v, q, a, w, w.stride(0), B, K, Q, A, C)
    ref = K_.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C)
    noise = torch.empty(64 << 20, device=DEV)
    side = torch.cuda.Stream()
    for it in range(100):
        if it % 2:                                   # perturb timing: a bandwidth hog on another stream
            with torch.cuda.stream(side):
                noise.normal_()
        out = K_.tri_pool_fwd(v, q, a, w, w.stride(0), B, K, Q, A, C)
        assert torch.equal(out, out0), it
        got = K_.tri_pool_bwd(v, q, a, w, w.stride(0), dout, B, K, Q, A, C)
        for i in (0, 1, 2, 6):                       # dzv, dzq, dza, dw
            assert torch.equal(got[i], ref[i]), (it, i)
    torch.cuda.synchronize()
