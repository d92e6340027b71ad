"""CPU oracle for the CTI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch fp32 restatement of the reference's algorithm for the compact
trilinear interaction path (aioz-ai/ICCV19_VQA-CTI).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module, and only as the checker / the timed
CPU baseline.  The product package (``iccv19_vqa-cti_b200``) never imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's
own modules from /root/reference (possible only in the build container), runs
them on seeded inputs and stores inputs, weights, outputs and gradients in
``tests/golden/*.pt``; ``tests/test_oracle.py`` checks every function below
against those files and against the two known answers the reference tree holds
(Kolda-Bader n-mode product, ``src/Tensor.py:31-32``; the softmax-attention
gradient of ``tools/grad_check.py:26``).

All functions take a flat ``dict`` of tensors keyed exactly like the reference
modules' ``state_dict()`` (old-style weight-norm keys ``...main.N.weight_g``,
``...main.N.weight_v``, ``...main.N.bias``) plus a key prefix, so a reference
``state_dict`` can be passed in unchanged.  Everything is eval-mode (dropout
is the identity), which is the parity target.

Two formulations are kept for the trilinear map:
  * ``tcnet_logits``      -- follows the reference's own order of operations
                             (rank loop, three mode products as matmuls);
                             this is the one timed as the CPU baseline "port".
  * ``tcnet_logits_closed`` -- the closed-form einsum with ``T_eff``; used to
                             cross-check and to define the CUDA kernels' math.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

Params = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------- #
# Optional bf16 rounding points.  By default every function below is the plain fp32 restatement
# of the reference.  Inside ``with bf16_rounding():`` the operands the CUDA kernels hold in bf16
# (cast inputs, folded weights, stored layer outputs, the packed core and the two on-chip
# intermediates of the contraction, the attention weights fed to the pooling) are rounded to bf16
# with a straight-through gradient.  The values then differ from the fp32 oracle by the bf16
# quantisation the kernels are allowed (north_star: "bf16 inputs, fp32 accumulate"), but every ReLU
# sees the same pre-activation sign as the kernels do, so autograd of this variant is the reference
# for the hand-written BACKWARD kernels (see tests/test_gpu_modules.py on why the ReLU masks matter).
# --------------------------------------------------------------------------- #
_ROUND = [False]


class bf16_rounding:
    """``keep``: optional predicate on the rounding site's name; sites it rejects stay fp32.  Site names are
    ``<layer prefix>{x,w,y}`` for the input, folded weight and stored output of a weight-normed layer, and
    ``core``, ``n1``, ``m``, ``att_w``, ``hq`` for the contraction / pooling operands -- used by
    tests/test_bf16_emulation_cpu.py to attribute the gradient error to individual rounding points."""

    def __init__(self, keep=None):
        self.keep = keep

    def __enter__(self):
        self.prev = _ROUND[0]
        _ROUND[0] = self.keep if self.keep is not None else True

    def __exit__(self, *exc):
        _ROUND[0] = self.prev


def _r(x: torch.Tensor, site: str = "") -> torch.Tensor:
    mode = _ROUND[0]
    if not mode or (mode is not True and not mode(site)):
        return x
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


# Optional externally supplied ReLU masks, keyed by the layer's parameter prefix
# (e.g. ``'v_att.TriAtt.v_net.3.main.1.'``).  A layer with an entry computes ``z * mask`` instead of
# ``relu(z)``: the oracle is then evaluated on the same linear piece of the network as the
# implementation whose masks were recorded, which removes the only discontinuous step of the path
# from a gradient comparison (a pre-activation within rounding distance of zero may get a different
# sign in bf16 arithmetic than in fp32; each such flip switches a gradient entry on or off).
_MASKS = [None]


class relu_masks:
    def __init__(self, masks: Dict[str, torch.Tensor]):
        self.masks = masks

    def __enter__(self):
        self.prev = _MASKS[0]
        _MASKS[0] = self.masks

    def __exit__(self, *exc):
        _MASKS[0] = self.prev


# --------------------------------------------------------------------------- #
# FCNet (reference src/fc.py:10-34)
# --------------------------------------------------------------------------- #
def fcnet_layer_index(dropout: float) -> int:
    """Index of the weight-normed Linear inside ``FCNet.main`` for a single
    layer net: a Dropout module comes first iff dropout > 0 (src/fc.py:25-27)."""
    return 1 if dropout > 0 else 0


def wn_linear(x: torch.Tensor, p: Params, prefix: str, act: str = "ReLU") -> torch.Tensor:
    """One weight-normed linear layer, ``weight_norm(nn.Linear, dim=None)``
    (src/fc.py:22,27): W = V * g / ||V||_F, y = act(x W^T + b)."""
    v = p[prefix + "weight_v"]
    g = p[prefix + "weight_g"]
    b = p[prefix + "bias"]
    w = _r(v * (g / v.norm()), prefix + "w")
    y = torch.matmul(_r(x, prefix + "x"), w.t()) + b
    if act == "ReLU":
        if _MASKS[0] is not None and prefix in _MASKS[0]:
            y = _r(y * _MASKS[0][prefix].reshape(y.shape).to(y.dtype), prefix + "y")
        else:
            y = _r(torch.relu(y), prefix + "y")      # fused layers store their (post-ReLU) output in bf16
    elif act != "":
        raise ValueError("oracle covers act in {'ReLU',''} only")
    return y


def fcnet(x: torch.Tensor, p: Params, prefix: str, act: str = "ReLU", dropout: float = 0.2) -> torch.Tensor:
    """Single-layer FCNet in eval mode (src/fc.py:33-34). ``prefix`` ends with '.'
    and names the FCNet module (e.g. ``'v_tucker.'``)."""
    li = fcnet_layer_index(dropout)
    return wn_linear(x, p, f"{prefix}main.{li}.", act)


# --------------------------------------------------------------------------- #
# n-mode products (reference src/Tensor.py:3-19) and the T_eff permutation
# --------------------------------------------------------------------------- #
def mode_product3(core: torch.Tensor, m1: torch.Tensor, m2: torch.Tensor, m3: torch.Tensor) -> torch.Tensor:
    """Three successive mode products of ``core`` (1,d1,d2,d3,G,1) with
    m1 (B,K,d1), m2 (B,Q,d2), m3 (B,A,d3) -> (B,K,Q,A,G), in the reference's
    order of operations (src/Tensor.py:6-19), *including* its mode-1 quirk:
    the core is flattened in (d3,d2,G) order (``transpose(3,2)`` then
    ``view``, :6) but the product is re-viewed as (G,d3,d2) (:8), which
    scrambles the glimpse axis for G>1.
    """
    _, d1, d2, d3, G, _ = core.shape
    B, K = m1.shape[0], m1.shape[1]
    # mode 1 (:6-8): rows of the unfolding are d1; columns are flat (l, j, g).
    unfold1 = core[0, ..., 0].permute(0, 2, 1, 3).reshape(1, d1, d3 * d2 * G)
    prod1 = torch.matmul(m1, unfold1)                       # (B,K,d3*d2*G)
    # the reference now reads that flat axis as (g', l', j')  -> t1[b,k,j',l',g']
    t1 = prod1.reshape(B, K, G, d3, d2).permute(0, 1, 4, 3, 2)
    # mode 2 (:11-13): contract j' with m2 -> t2[b,k,q,l',g']
    t2 = torch.einsum("bqj,bkjlg->bkqlg", m2, t1.float())
    # mode 3 (:17-19): contract l' with m3 -> (B,K,Q,A,G)
    return torch.einsum("bal,bkqlg->bkqag", m3, t2.float())


def teff_from_tg(t_g: torch.Tensor) -> torch.Tensor:
    """Effective core of the trilinear map, (R,d,d,d,G), such that
    logits = sum_r einsum(T_eff[r], Vc_r, Qc_r, Ac_r) equals the reference's
    rank loop (src/tc.py:46-50 with src/Tensor.py:6-8).  Identity for G == 1."""
    _, R, d1, d2, d3, G, ho = t_g.shape
    assert ho == 1, "h_out' > 1 is unusable in the reference (ModeProduct view fails)"
    t = t_g[0, ..., 0]                                               # (R,i,j,l,g)
    flat = t.permute(0, 1, 3, 2, 4).reshape(R, d1, G, d3, d2)        # (l,j,g) re-read as (g',l',j')
    return flat.permute(0, 1, 4, 3, 2).contiguous()                  # (R,i,j',l',g')


def teff_index_map(R: int, d: int, G: int) -> torch.Tensor:
    """int64 tensor of shape (R,d,d,d,G): T_eff.flat[n] = T_g.flat[map.flat[n]]."""
    ar = torch.arange(R * d * d * d * G, dtype=torch.float64).view(1, R, d, d, d, G, 1)
    return teff_from_tg(ar).round().long()


# --------------------------------------------------------------------------- #
# TCNet / TriAttention (reference src/tc.py, src/attention.py:43-59)
# --------------------------------------------------------------------------- #
def tcnet_projections(v, q, a, p: Params, prefix: str, rank: int):
    """Tucker projections then the R per-rank projections (src/tc.py:43-49).
    Returns Vc (B,K,R,d), Qc (B,Q,R,d), Ac (B,A,R,d)."""
    vt = fcnet(v, p, prefix + "v_tucker.", dropout=0.5)
    qt = fcnet(q, p, prefix + "q_tucker.", dropout=0.2)
    at = fcnet(a, p, prefix + "a_tucker.", dropout=0.2)
    vc = torch.stack([fcnet(vt, p, f"{prefix}v_net.{r}.", dropout=0.5) for r in range(rank)], 2)
    qc = torch.stack([fcnet(qt, p, f"{prefix}q_net.{r}.", dropout=0.2) for r in range(rank)], 2)
    ac = torch.stack([fcnet(at, p, f"{prefix}a_net.{r}.", dropout=0.2) for r in range(rank)], 2)
    return vc, qc, ac


def tcnet_logits(v, q, a, p: Params, prefix: str = "") -> torch.Tensor:
    """``TCNet.forward`` (src/tc.py:41-52) in the reference's own order:
    for each rank three tiny WN-linears and a 3-mode product, accumulated."""
    t_g = p[prefix + "T_g"]
    rank = t_g.shape[1]
    vt = fcnet(v, p, prefix + "v_tucker.", dropout=0.5)
    qt = fcnet(q, p, prefix + "q_tucker.", dropout=0.2)
    at = fcnet(a, p, prefix + "a_tucker.", dropout=0.2)
    out = 0
    for r in range(rank):
        v_ = fcnet(vt, p, f"{prefix}v_net.{r}.", dropout=0.5)
        q_ = fcnet(qt, p, f"{prefix}q_net.{r}.", dropout=0.2)
        a_ = fcnet(at, p, f"{prefix}a_net.{r}.", dropout=0.2)
        out = mode_product3(t_g[:, r], v_, q_, a_) + out
    return out


def tcnet_logits_closed(v, q, a, p: Params, prefix: str = "") -> torch.Tensor:
    """Closed form of ``TCNet.forward``."""
    t_g = p[prefix + "T_g"]
    vc, qc, ac = tcnet_projections(v, q, a, p, prefix, t_g.shape[1])
    return trilinear_closed(vc, qc, ac, teff_from_tg(t_g))


def trilinear_closed(vc, qc, ac, t_eff) -> torch.Tensor:
    """logits[b,k,q,a,g] = sum_{r,i,j,l} T_eff[r,i,j,l,g] Vc[b,k,r,i] Qc[b,q,r,j] Ac[b,a,r,l],
    contracted a -> q -> v (the minimal-FLOP order)."""
    n1 = _r(torch.einsum("balr,rijlg->barijg", ac.permute(0, 1, 3, 2), _r(t_eff, "core")), "n1")
    m = _r(torch.einsum("bqrj,barijg->briqag", qc, n1), "m")
    return torch.einsum("bkri,briqag->bkqag", vc, m)


def zero_row_mask(v: torch.Tensor) -> torch.Tensor:
    """mask[b,k] = (sum_c |v[b,k,c]| == 0)  (src/attention.py:55, :36)."""
    return v.abs().sum(2) == 0


def tri_attention(v, q, a, p: Params, prefix: str = "TriAtt.") -> Tuple[torch.Tensor, torch.Tensor]:
    """``TriAttention.forward`` (src/attention.py:49-59): logits, zero-row mask to
    -inf, softmax over the flattened (k,q,a) axis per (b,g). Returns (p, logits)."""
    # with rounding points on, follow the kernels' contraction order (a -> q -> v) so they apply
    logits = tcnet_logits_closed(v, q, a, p, prefix) if _ROUND[0] else tcnet_logits(v, q, a, p, prefix)
    B, K, Q, A, G = logits.shape
    logits = logits.masked_fill(zero_row_mask(v)[:, :, None, None, None], float("-inf"))
    att = torch.softmax(logits.reshape(B, K * Q * A, G), 1).view(B, K, Q, A, G)
    return att, logits


def tcnet_pool(v, q, a, w, p: Params, prefix: str = "") -> torch.Tensor:
    """``TCNet.forward_with_weights`` (src/tc.py:54-61):
    out[b,c] = sum_{k,q,a} V[b,k,c] w[b,k,q,a] Q[b,q,c] A[b,a,c]."""
    vp = fcnet(v, p, prefix + "v_tucker.", dropout=0.5)
    qp = fcnet(q, p, prefix + "q_tucker.", dropout=0.2)
    ap = fcnet(a, p, prefix + "a_tucker.", dropout=0.2)
    return trilinear_pool(vp, qp, ap, w)


def trilinear_pool(vp, qp, ap, w) -> torch.Tensor:
    return torch.einsum("bkc,bkqa,bqc,bac->bc", vp, _r(w, "att_w"), qp, ap)


# --------------------------------------------------------------------------- #
# BCNet / BiAttention (reference src/bc.py, src/attention.py:14-40)
# --------------------------------------------------------------------------- #
def bcnet_logits(v, q, p: Params, prefix: str = "", h_mat: torch.Tensor | None = None) -> torch.Tensor:
    """``BCNet.forward`` in the ``h_out <= 32`` branch (src/bc.py:52-58):
    logits[b,g,k,q] = sum_c Vb[b,k,c] h_mat[g,c] Qb[b,q,c] + h_bias[g]."""
    vb = fcnet(v, p, prefix + "v_net.", dropout=0.2)
    qb = fcnet(q, p, prefix + "q_net.", dropout=0.2)
    if h_mat is None:
        if prefix + "h_mat" in p:
            h_mat = p[prefix + "h_mat"]
        else:  # weight_norm(name='h_mat', dim=None), src/attention.py:19-20
            hv = p[prefix + "h_mat_v"]
            h_mat = hv * (p[prefix + "h_mat_g"] / hv.norm())
    h_bias = p[prefix + "h_bias"]
    return bilinear_closed(vb, qb, h_mat, h_bias)


def bilinear_closed(vb, qb, h_mat, h_bias) -> torch.Tensor:
    if _ROUND[0]:                                    # the kernel folds h_mat into the question operand
        hq = _r(qb.unsqueeze(1) * h_mat, "hq")       # (B,G,Q,C)
        return torch.matmul(vb.unsqueeze(1), hq.transpose(2, 3)) + h_bias
    hv = vb.unsqueeze(1) * h_mat                     # (B,G,K,C)
    return torch.matmul(hv, qb.unsqueeze(1).transpose(2, 3)) + h_bias


def bi_attention(v, q, p: Params, prefix: str = "logits.", v_mask: bool = True):
    """``BiAttention.forward_all`` (src/attention.py:30-40). Returns (p, logits)."""
    logits = bcnet_logits(v, q, p, prefix)
    B, G, K, Q = logits.shape
    if v_mask:
        logits = logits.masked_fill(zero_row_mask(v)[:, None, :, None], float("-inf"))
    att = torch.softmax(logits.reshape(B, G, K * Q), 2).view(B, G, K, Q)
    return att, logits


def bcnet_pool(v, q, w, p: Params, prefix: str = "", k: int = 1) -> torch.Tensor:
    """``BCNet.forward_with_weights`` (src/bc.py:70-78):
    out[b,c] = sum_{k,q} V[b,k,c] w[b,k,q] Q[b,q,c]; for k > 1 channel groups of
    k are sum-pooled (AvgPool1d(k) * k, :75-77)."""
    vp = fcnet(v, p, prefix + "v_net.", dropout=0.2)
    qp = fcnet(q, p, prefix + "q_net.", dropout=0.2)
    out = torch.einsum("bkc,bkq,bqc->bc", vp, _r(w, "att_w"), qp)
    if k > 1:
        out = out.view(out.shape[0], -1, k).sum(2)
    return out


# --------------------------------------------------------------------------- #
# Caller glue that fixes call order and shapes
# --------------------------------------------------------------------------- #
def cti_hot_path(v, q_emb, a_emb, p: Params, glimpse: int):
    """The hot-path slice of ``TanModel.forward`` / ``CTIModel.forward``
    (src/MC/base_model.py:143-150, src/FFOE/base_model.py:127-134):
    attention once, then per glimpse pooling + residual q/a updates.
    Keys: ``v_att.TriAtt.*``, ``t_net.{g}.*``, ``q_prj.{g}.*``, ``a_prj.{g}.*``.
    Returns (joint (B,1024), att, logits)."""
    att, logits = tri_attention(v, q_emb, a_emb, p, "v_att.TriAtt.")
    for g in range(glimpse):
        b_emb = tcnet_pool(v, q_emb, a_emb, att[:, :, :, :, g], p, f"t_net.{g}.")
        q_emb = fcnet(b_emb.unsqueeze(1), p, f"q_prj.{g}.", act="", dropout=0.2) + q_emb
        a_emb = fcnet(b_emb.unsqueeze(1), p, f"a_prj.{g}.", act="", dropout=0.2) + a_emb
    joint = q_emb.sum(1) + a_emb.sum(1)
    return joint, att, logits


def ban_hot_path(v, q_emb, p: Params, glimpse: int):
    """Hot-path slice of the FFOE ``BanModel.forward`` without the counter
    (src/FFOE/base_model.py:50-66). Returns (joint (B,1024), att, logits)."""
    att, logits = bi_attention(v, q_emb, p, "v_att.logits.")
    q_list = []
    for g in range(glimpse):
        b_emb = bcnet_pool(v, q_emb, att[:, g], p, f"b_net.{g}.")
        q_emb = fcnet(b_emb.unsqueeze(1), p, f"q_prj.{g}.", act="", dropout=0.2) + q_emb
        q_list.append(q_emb)
    joint = torch.stack(q_list, 1).sum(1).sum(1)
    return joint, att, logits



# --------------------------------------------------------------------------- #
# Right after the hot path: classifier and the trainer's update tail
# --------------------------------------------------------------------------- #
def simple_classifier(x: torch.Tensor, p: Params, prefix: str = "classifier.") -> torch.Tensor:
    """``SimpleClassifier.forward`` in eval mode with ``activation='relu'`` (src/classifier.py:19-28):
    weight_norm(Linear) -> ReLU -> Dropout (identity in eval) -> weight_norm(Linear); modules main.0 and main.3."""
    h = wn_linear(x, p, prefix + "main.0.", "ReLU")
    return wn_linear(h, p, prefix + "main.3.", "")


def gru_forward_all(x: torch.Tensor, p: Params, prefix: str = "rnn.") -> torch.Tensor:
    """``QuestionEmbedding.forward_all`` (src/language_model.py:93-98): ``nn.GRU(in, H, 1, batch_first=True)`` from a zero
    initial state, every hidden state returned, (B, T, in) -> (B, T, H).  The cell arithmetic is third-party
    (torch.nn.GRU; requirements.txt pins torch==1.1.0), restated from its documented equations, gate order r, z, n:
    r = s(W_ir x + b_ir + W_hr h + b_hr), z = s(W_iz x + b_iz + W_hz h + b_hz),
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) n + z h.  ``forward`` (:80-91) is ``[:, -1]`` of this."""
    w_ih, w_hh = p[prefix + "weight_ih_l0"], p[prefix + "weight_hh_l0"]
    b_ih, b_hh = p[prefix + "bias_ih_l0"], p[prefix + "bias_hh_l0"]
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gx = torch.matmul(_r(x, prefix + "x"), _r(w_ih, prefix + "w").t()) + b_ih
    h = x.new_zeros(B, H)
    outs = []
    for t in range(T):
        gh = torch.matmul(_r(h, prefix + "h"), _r(w_hh, prefix + "w").t()) + b_hh
        r = torch.sigmoid(gx[:, t, :H] + gh[:, :H])
        z = torch.sigmoid(gx[:, t, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gx[:, t, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, 1)


def word_embedding(tokens: torch.Tensor, p: Params, prefix: str = "w_emb.") -> torch.Tensor:
    """``WordEmbedding.forward`` with op 'c' in eval mode (src/language_model.py:42-47): the trainable table and the
    frozen one, concatenated -> (B, T, 600)."""
    emb = torch.nn.functional.embedding(tokens, p[prefix + "emb.weight"])
    if prefix + "emb_.weight" in p:
        emb = torch.cat((emb, torch.nn.functional.embedding(tokens, p[prefix + "emb_.weight"])), 2)
    return emb


def mc_model_logits(v, q_tok, a_tok, p: Params, glimpse: int):
    """``TanModel.forward`` (src/MC/base_model.py:134-152): word embeddings -> the two GRUs -> the hot path ->
    classifier.  Keys as in the reference's ``state_dict()``.  Returns (class logits (B, 2), att)."""
    q_emb = gru_forward_all(word_embedding(q_tok, p, "w_emb."), p, "q_emb.rnn.")
    a_emb = gru_forward_all(word_embedding(a_tok, p, "wa_emb."), p, "ans_emb.rnn.")
    joint, att, _ = cti_hot_path(v, q_emb, a_emb, p, glimpse)
    return simple_classifier(joint, p, "classifier."), att


def mc_answers(logits: torch.Tensor) -> torch.Tensor:
    """The chosen candidate of each question: argmax over its 4 rows of softmax(logits)[:, 0]
    (``compute_score_mc``, src/MC/trainer.py:292-299)."""
    return torch.softmax(logits, 1)[:, 0].view(-1, 4).argmax(1)


def trainer_update(params, grads, exp_avg, exp_inf, step: int, lr: float, grad_denom: float, clip_norm: float,
                   beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> float:
    """One update of the reference trainer's tail, in place on lists of fp32 tensors; returns the pre-clip norm.

    ``Trainer._all_reduce_and_rescale`` (src/MC/trainer.py:208-219): the flat gradient is divided by ``grad_denom``
    and clipped by its global L2 norm, ``coef = clip_norm / (norm + 1e-6)`` applied only if ``norm > clip_norm > 0``
    (``clip_grad_norm_``, src/utils.py:323-328).  Then ``torch.optim.Adamax`` (src/MC/train.py:32; third-party,
    requirements.txt pins torch==1.1.0; Kingma & Ba 2015, algorithm 2): ``m = b1 m + (1-b1) g``,
    ``u = max(b2 u, |g| + eps)``, ``p -= lr / (1 - b1^t) * m / u`` with t = ``step`` counted from 1."""
    flat = torch.cat([g.reshape(-1) for g in grads]) / grad_denom
    norm = float(flat.norm())
    coef = clip_norm / (norm + 1e-6) if norm > clip_norm > 0 else 1.0
    clr = lr / (1.0 - beta1 ** step)
    for prm, g, m, u in zip(params, grads, exp_avg, exp_inf):
        gg = g / grad_denom * coef
        m.mul_(beta1).add_(gg, alpha=1.0 - beta1)
        torch.maximum(u * beta2, gg.abs() + eps, out=u)
        prm.addcdiv_(m, u, value=-clr)
    return norm


def distillation_loss(x, teacher, target, T: float, alpha: float) -> torch.Tensor:
    """``Distillation_Loss.forward`` (src/loss_function.py:20-25)."""
    logp = torch.log_softmax(x / T, dim=1)
    pt = torch.softmax(teacher / T, dim=1)
    kl = (pt * (torch.log(pt.clamp_min(1e-45)) - logp)).sum(1).mean() * (alpha * T * T)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(x, target, reduction="sum") / x.size(0)
    return kl + bce * (1.0 - alpha)


# --------------------------------------------------------------------------- #
# Synthetic inputs and random parameters with the
# reference constructors' shapes -- used where no reference import is possible
# (the GPU box).
# --------------------------------------------------------------------------- #
def synthetic_inputs(B, K, Q, A, v_dim=2048, q_dim=1024, a_dim=1024, seed=1204, min_boxes=10,
                     device="cpu"):
    g = torch.Generator().manual_seed(seed)
    v = torch.relu(torch.randn(B, K, v_dim, generator=g))
    nb = torch.randint(min(min_boxes, K), K + 1, (B,), generator=g)
    v = v * (torch.arange(K)[None, :] < nb[:, None]).float()[:, :, None]
    q = torch.tanh(torch.randn(B, Q, q_dim, generator=g))
    a = torch.tanh(torch.randn(B, A, a_dim, generator=g)) if A else None
    return (v.to(device), q.to(device), None if a is None else a.to(device))


def _init_wn_linear(p: Params, prefix: str, fin: int, fout: int, g: torch.Generator):
    bound = 1.0 / math.sqrt(fin)          # nn.Linear default init (kaiming_uniform a=sqrt(5))
    v = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound
    p[prefix + "bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound
    p[prefix + "weight_g"] = v.norm().clone()
    p[prefix + "weight_v"] = v


def random_tcnet_params(prefix, v_dim, q_dim, a_dim, h_dim, rank, glimpse, k, g, p=None) -> Params:
    """Same tensors, shapes and key order as ``TCNet.__init__`` (src/tc.py:10-38)."""
    p = {} if p is None else p
    H = h_dim * k
    d = h_dim // rank
    if H < 1024:
        p[prefix + "T_g"] = torch.randn(1, rank, d, d, d, glimpse, 1, generator=g)
    _init_wn_linear(p, prefix + "v_tucker.main.1.", v_dim, H, g)
    _init_wn_linear(p, prefix + "q_tucker.main.1.", q_dim, H, g)
    _init_wn_linear(p, prefix + "a_tucker.main.1.", a_dim, H, g)
    if H < 1024:
        for name in ("v_net", "q_net", "a_net"):
            for r in range(rank):
                _init_wn_linear(p, f"{prefix}{name}.{r}.main.1.", H, d, g)
    return p


def random_cti_params(v_dim=2048, num_hid=1024, h_mm=512, rank=32, glimpse=2, seed=1204) -> Params:
    """Hot-path parameters of ``build_cti`` (src/MC/base_model.py:196-205)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    random_tcnet_params("v_att.TriAtt.", v_dim, num_hid, num_hid, h_mm, rank, glimpse, 1, g, p)
    for i in range(glimpse):
        random_tcnet_params(f"t_net.{i}.", v_dim, num_hid, num_hid, h_mm, rank, 1, 2, g, p)
    for i in range(glimpse):
        _init_wn_linear(p, f"q_prj.{i}.main.1.", num_hid, num_hid, g)
        _init_wn_linear(p, f"a_prj.{i}.main.1.", num_hid, num_hid, g)
    return p


def random_ban_params(v_dim=2048, num_hid=1024, glimpse=2, seed=1204) -> Params:
    """Hot-path parameters of FFOE ``build_ban`` (src/FFOE/base_model.py:139-159)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    pre = "v_att.logits."
    p[pre + "h_bias"] = torch.randn(1, glimpse, 1, 1, generator=g)
    _init_wn_linear(p, pre + "v_net.main.1.", v_dim, num_hid * 3, g)
    _init_wn_linear(p, pre + "q_net.main.1.", num_hid, num_hid * 3, g)
    hv = torch.randn(1, glimpse, 1, num_hid * 3, generator=g)
    p[pre + "h_mat_g"] = hv.norm().clone()
    p[pre + "h_mat_v"] = hv
    for i in range(glimpse):
        _init_wn_linear(p, f"b_net.{i}.v_net.main.1.", v_dim, num_hid, g)
        _init_wn_linear(p, f"b_net.{i}.q_net.main.1.", num_hid, num_hid, g)
    for i in range(glimpse):
        _init_wn_linear(p, f"q_prj.{i}.main.1.", num_hid, num_hid, g)
    return p
