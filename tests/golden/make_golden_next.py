"""Golden vectors for the components right after the hot path, made by running the
REFERENCE's own code.  Build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_next.py

Writes tests/golden/next_golden.pt:
  * ``clf_*``   -- ``SimpleClassifier`` (src/classifier.py:11-29) forward + backward in eval mode,
  * ``gru_*``   -- ``QuestionEmbedding.forward_all`` / ``forward`` (src/language_model.py:50-98; nn.GRU on the CPU) forward +
                   backward, including a 300-wide input (not a multiple of 8),
  * ``trainer`` -- three update steps of the reference trainer's tail: gradients / grad_denom, global-norm clip with
                   ``src.utils.clip_grad_norm_`` (src/utils.py:323-328, called from src/MC/trainer.py:208-219) and
                   ``torch.optim.Adamax`` (src/MC/train.py:32) on a handful of odd-sized parameters.
"""
import collections
import collections.abc
import os
import sys
import types
import warnings

import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = os.environ.get("CTI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
six = types.ModuleType("torch._six")
six.string_classes = (str, bytes)
sys.modules.setdefault("torch._six", six)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
collections.Mapping, collections.Sequence = collections.abc.Mapping, collections.abc.Sequence

from src.classifier import SimpleClassifier  # noqa: E402
from src.language_model import QuestionEmbedding  # noqa: E402
import src.utils as ref_utils  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "next_golden.pt")


def main():
    out = {}
    torch.manual_seed(5)
    args = types.SimpleNamespace(activation="relu", dropout=0.5)
    for name, (i, h, o, rows) in {"clf_mc": (64, 128, 2, (6,)), "clf_ffoe": (48, 96, 40, (3, 5))}.items():
        m = SimpleClassifier(i, h, o, args).eval()
        m.main[2].inplace = False             # in-place dropout after ReLU trips torch-2 autograd checks; same numbers
        x = torch.randn(*rows, i, requires_grad=True)
        y = m(x)
        c = torch.randn(y.shape)
        (y * c).sum().backward()
        out[name] = {"sd": {k: t.detach().clone() for k, t in m.state_dict().items()}, "x": x.detach().clone(),
                     "y": y.detach().clone(), "cot": c, "dx": x.grad.clone(),
                     "grads": {k: p.grad.detach().clone() for k, p in m.named_parameters()}, "dims": (i, h, o)}

    # ---- trainer tail: rescale, clip, Adamax ------------------------------------------------------
    g = torch.Generator().manual_seed(9)
    shapes = [(7, 33), (), (129,), (16, 16, 3), (1, 5, 1, 64), (4097,)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    opt = torch.optim.Adamax(params, lr=7e-4)                      # src/MC/train.py:32 (lr_default = 7e-4 there too)
    clip_norm, steps = 0.25, []
    p0 = [p.detach().clone() for p in params]
    for step, (scale, denom) in enumerate([(3.0, 64.0), (0.01, 4.0), (50.0, 256.0)]):
        grads = [scale * torch.randn(s, generator=g) for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        # Trainer._all_reduce_and_rescale (src/MC/trainer.py:208-219) on the flat buffer
        flat = torch.cat([p.grad.view(-1) for p in params])
        flat.div_(denom)
        norm = ref_utils.clip_grad_norm_(flat, clip_norm)
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        opt.step()
        steps.append({"grads": grads, "denom": denom, "norm": float(norm),
                      "params": [p.detach().clone() for p in params]})
    out["trainer"] = {"shapes": shapes, "p0": p0, "lr": 7e-4, "clip_norm": clip_norm, "steps": steps,
                      "betas": (0.9, 0.999), "eps": 1e-8}
    # ---- QuestionEmbedding (appended last so the cases above keep their RNG streams) -------------------
    torch.manual_seed(21)
    for name, (din, hid, rows, T) in {"gru_small": (24, 32, 3, 5), "gru_300": (300, 64, 4, 7)}.items():
        m = QuestionEmbedding(din, hid, 1, False, .0).eval()
        x = torch.randn(rows, T, din, requires_grad=True)
        y = m.forward_all(x)
        last = m.forward(x)
        c = torch.randn(y.shape)
        (y * c).sum().backward()
        out[name] = {"sd": {k: t.detach().clone() for k, t in m.state_dict().items()}, "x": x.detach().clone(),
                     "y": y.detach().clone(), "last": last.detach().clone(), "cot": c, "dx": x.grad.clone(),
                     "grads": {k: p.grad.detach().clone() for k, p in m.named_parameters()}, "dims": (din, hid)}
    # ---- the whole multiple-choice model, small but with the real structure (d = h_mm / rank = 16, k = 2 pooling) ------
    import src.MC.base_model as mc
    torch.manual_seed(33)
    ds = types.SimpleNamespace(dictionary=types.SimpleNamespace(ntoken=50), v_dim=64, num_ans_candidates=4)
    margs = types.SimpleNamespace(op="c", num_hid=128, gamma=2, h_mm=64, h_out=1, rank=4, k=1, activation="relu",
                                  dropout=0.5, use_counter=False, num_stacks=2)
    model = mc.build_cti(margs, ds).eval()
    model.classifier.main[2].inplace = False
    with torch.no_grad():
        model.w_emb.emb_.weight.normal_()                 # the frozen table is filled from GloVe in the reference
        model.wa_emb.emb_.weight.normal_()
    gq = torch.Generator().manual_seed(34)
    nq, K = 3, 12
    v = torch.relu(torch.randn(nq, K, 64, generator=gq))
    nb = torch.randint(6, K + 1, (nq,), generator=gq)
    v = v * (torch.arange(K)[None, :] < nb[:, None]).float()[:, :, None]
    v4 = v.unsqueeze(1).expand(nq, 4, K, 64).contiguous().view(nq * 4, K, 64)      # src/MC/train.py:75-76
    q_tok = torch.randint(0, 50, (nq, 12), generator=gq)
    q4 = q_tok.unsqueeze(1).expand(nq, 4, 12).contiguous().view(nq * 4, 12)
    a_tok = torch.randint(0, 51, (nq * 4, 6), generator=gq)                          # 50 = padding id
    labels = torch.zeros(nq * 4, 2)
    labels[torch.arange(nq * 4), torch.randint(0, 2, (nq * 4,), generator=gq)] = 1.0
    logits, att = model(v4, None, q4, a_tok)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction="sum") / labels.size(0)
    loss.backward()
    out["mc_model"] = {"args": dict(ntoken=50, v_dim=64, num_hid=128, h_mm=64, rank=4, gamma=2),
                       "sd": {k: t.detach().clone() for k, t in model.state_dict().items()},
                       "v": v, "q_tok": q4, "a_tok": a_tok, "labels": labels, "logits": logits.detach().clone(),
                       "att": att.detach().clone().contiguous(), "loss": loss.detach().clone(),
                       "grads": {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}}
    # ---- the BAN distillation student (BASELINE config 3 in miniature): build_ban + Distillation_Loss(T=5, alpha=0.005) ----
    import src.FFOE.base_model as ff
    from src.loss_function import Distillation_Loss
    torch.manual_seed(44)
    n_ans = 37
    ds = types.SimpleNamespace(dictionary=types.SimpleNamespace(ntoken=50), v_dim=64, num_ans_candidates=n_ans)
    bargs = types.SimpleNamespace(op="c", num_hid=128, gamma=2, h_mm=64, h_out=1, rank=4, k=1, activation="relu",
                                  dropout=0.5, use_counter=False, num_stacks=2)
    ban = ff.build_ban(bargs, ds).eval()
    ban.classifier.main[2].inplace = False
    with torch.no_grad():
        ban.w_emb.emb_.weight.normal_()
    gb = torch.Generator().manual_seed(45)
    Bn, K = 5, 12
    v = torch.relu(torch.randn(Bn, K, 64, generator=gb))
    nb = torch.randint(6, K + 1, (Bn,), generator=gb)
    v = v * (torch.arange(K)[None, :] < nb[:, None]).float()[:, :, None]
    boxes = torch.rand(Bn, K, 6, generator=gb)
    q_tok = torch.randint(0, 51, (Bn, 12), generator=gb)
    teacher = torch.randn(Bn, n_ans, generator=gb).half().float()             # fp16 teacher logits (src/FFOE/test.py:129)
    target = torch.zeros(Bn, n_ans)
    target[torch.arange(Bn), torch.randint(0, n_ans, (Bn,), generator=gb)] = 1.0
    target[torch.arange(Bn), torch.randint(0, n_ans, (Bn,), generator=gb)] += 0.3  # soft scores (tools/compute_softscore.py:86-96)
    target.clamp_(max=1.0)
    logits, att = ban(v, boxes, q_tok, None)
    loss = Distillation_Loss(T=5, alpha=0.005)(logits, teacher, target)
    loss.backward()
    out["ban_model"] = {"args": dict(ntoken=50, v_dim=64, num_hid=128, gamma=2, n_ans=n_ans),
                        "sd": {k: t.detach().clone() for k, t in ban.state_dict().items()},
                        "v": v, "boxes": boxes, "q_tok": q_tok, "teacher": teacher, "target": target,
                        "logits": logits.detach().clone(), "att": att.detach().clone().contiguous(),
                        "loss": loss.detach().clone(),
                        "grads": {k: p.grad.detach().clone().bfloat16() for k, p in ban.named_parameters()
                                  if p.grad is not None}}
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
