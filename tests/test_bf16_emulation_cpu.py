"""What bf16 operands alone do to the gradients of the CTI hot path -- no kernel involved (CPU only).

The north_star asks for gradients within 3e-2 relative of the reference's fp32 path with "bf16 inputs, fp32
accumulate".  Every projection on the path ends in a ReLU, so rounding GEMM operands to bf16 flips the sign of a few
near-zero pre-activations and each flip switches gradient entries on or off.  This test pins how large that effect is
by comparing autograd of the fp32 oracle with autograd of the SAME oracle evaluated with the kernels' bf16 rounding
points (``O.bf16_rounding``), per tensor, for two losses:

  * ``hot``   -- the hot path's own output contracted with a random cotangent (what bench.py and the GPU parity tests
                 differentiate; the residual connections of src/MC/base_model.py:147-148 dominate dq / da);
  * ``smoke`` -- the sum of the two pooled embeddings without the residuals (the loss of __graft_entry__.smoke() in
                 round 1): dq / da here come only through ReLU-gated projections.

It also attributes the error to rounding sites (``bf16_rounding(keep=...)``): keeping the 96 per-rank 512->16
projections of src/tc.py:47-49 in fp32 (what a tf32 / split-bf16 version of that grouped GEMM would buy) changes
nothing, because the flips that matter come from the bf16 rounding of the *tucker* layers' operands feeding them and
from the pooling-side projections.  The GPU tests (tests/test_gpu_modules.py, tests/test_gpu_baseline_sizes.py) hold
the kernels to these emulated numbers plus a margin, and to 3e-2 on the whole flat gradient.
"""
import re

import pytest
import torch

from oracle import cti_oracle as O

B, K, Q, A, G = 4, 50, 12, 6, 2


def _grads(keep, loss_kind):
    """keep: False = plain fp32 oracle; None = every rounding point; callable = site filter."""
    params = O.random_cti_params(glimpse=G, seed=1204)
    v, q, a = O.synthetic_inputs(B, K, Q, A, seed=7)
    cot = torch.randn(B, 1024, generator=torch.Generator().manual_seed(3))
    pl = {k: t.clone().requires_grad_(True) for k, t in params.items()}
    ql, al = q.clone().requires_grad_(True), a.clone().requires_grad_(True)

    def loss_fn():
        if loss_kind == "smoke":
            p_att, _ = O.tri_attention(v, ql, al, pl, "v_att.TriAtt.")
            return sum(O.tcnet_pool(v, ql, al, p_att[:, :, :, :, i], pl, f"t_net.{i}.").sum() for i in range(G))
        joint, _, _ = O.cti_hot_path(v, ql, al, pl, G)
        return (joint * cot).sum()
    if keep is False:
        loss_fn().backward()
    else:
        with O.bf16_rounding(keep):
            loss_fn().backward()
    g = {k: t.grad for k, t in pl.items() if t.grad is not None}
    g["dq"], g["da"] = ql.grad, al.grad
    return g


def emulated_errors(loss_kind, keep=None):
    """-> dict: 'dq', 'da', 'params' (all parameter gradients as one flat vector), 'all' (parameters + dq + da) and one
    entry per parameter-tensor class (per-rank nets pooled over ranks): L2-relative error of the bf16-rounded oracle."""
    ref, got = _grads(False, loss_kind), _grads(keep, loss_kind)
    nr = lambda ks: (sum((got[k] - ref[k]).pow(2).sum().item() for k in ks) /
                     max(sum(ref[k].pow(2).sum().item() for k in ks), 1e-60)) ** 0.5
    pkeys = [k for k in ref if k not in ("dq", "da")]
    out = {"dq": nr(["dq"]), "da": nr(["da"]), "params": nr(pkeys), "all": nr(list(ref))}
    classes = {}
    for k in pkeys:
        kk = re.sub(r"_net\.\d+\.", "_net.*.", k)
        classes.setdefault(kk, []).append(k)
    out.update({kk: nr(ks) for kk, ks in classes.items()})
    return out


_RANK = re.compile(r"TriAtt\.[vqa]_net\.")


@pytest.fixture(scope="module")
def table():
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    t = {("hot", "all sites"): emulated_errors("hot"),
         ("smoke", "all sites"): emulated_errors("smoke"),
         ("smoke", "per-rank nets fp32"): emulated_errors("smoke", lambda s: not _RANK.search(s)),
         ("smoke", "pooling side only"): emulated_errors("smoke", lambda s: s.startswith("t_net") or s == "att_w")}
    print()
    for (loss, sites), e in t.items():
        worst = sorted(((v, k) for k, v in e.items() if k not in ("dq", "da", "params", "all")), reverse=True)[:3]
        print(f"bf16 emulation [{loss:5s} | {sites:20s}] dq {e['dq']:.4f} da {e['da']:.4f} params {e['params']:.4f} "
              f"all {e['all']:.4f} worst: " + ", ".join(f"{k} {v:.3f}" for v, k in worst))
    return t


def test_hot_path_loss_meets_3e2_on_the_flat_gradient(table):
    e = table[("hot", "all sites")]
    assert e["all"] <= 1e-2 and e["params"] <= 1e-2 and e["dq"] <= 3e-3 and e["da"] <= 3e-3, e


def test_smoke_loss_exceeds_3e2_by_rounding_alone(table):
    """dq / da of the residual-free loss: bf16 operand rounding alone costs 3-4 % (round 1's smoke printed 3.455e-2 for
    the kernels) -- the bound of the north_star is not attainable there by ANY bf16-operand forward."""
    e = table[("smoke", "all sites")]
    assert 2.5e-2 <= e["dq"] <= 4.5e-2 and 2.5e-2 <= e["da"] <= 5e-2, e
    assert e["params"] <= 3e-2, e


def test_wider_operands_in_the_per_rank_nets_do_not_help(table):
    """run the grouped per-rank projections in tf32 / split bf16.  Emulated here by leaving their
    x, w and y unrounded: dq, da and the per-rank nets' own gradient error do not move (the flips come from upstream)."""
    e0, e1 = table[("smoke", "all sites")], table[("smoke", "per-rank nets fp32")]
    for k in ("dq", "da", "params", "v_att.TriAtt.a_net.*.main.1.weight_v", "v_att.TriAtt.q_net.*.main.1.weight_v"):
        assert abs(e1[k] - e0[k]) <= 0.1 * e0[k] + 1e-4, (k, e0[k], e1[k])
    # almost all of dq / da's error is the pooling side's (t_net) projections
    e2 = table[("smoke", "pooling side only")]
    assert e2["dq"] >= 0.85 * e0["dq"] and e2["da"] >= 0.7 * e0["da"], (e0, e2)
