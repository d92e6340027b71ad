"""Generate golden vectors by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/cti_golden.pt (a dict of small fp32 tensors).  The file is
committed; this script is the record of how it was made.  Nothing here is
imported by the product.
"""
import os
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
sys.dont_write_bytecode = True
REF = os.environ.get("CTI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from src.attention import BiAttention, TriAttention  # noqa: E402
from src.bc import BCNet  # noqa: E402
from src.fc import FCNet  # noqa: E402
from src.tc import TCNet  # noqa: E402
import src.Tensor as RefTensor  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cti_golden.pt")


def inputs(B, K, Q, A, vd, qd, seed):
    g = torch.Generator().manual_seed(seed)
    v = torch.relu(torch.randn(B, K, vd, generator=g))
    nb = torch.randint(max(1, K // 2), K + 1, (B,), generator=g)
    v = v * (torch.arange(K)[None, :] < nb[:, None]).float()[:, :, None]
    q = torch.tanh(torch.randn(B, Q, qd, generator=g))
    a = torch.tanh(torch.randn(B, A, qd, generator=g))
    return v, q, a


def sd(m):
    return {k: t.detach().clone() for k, t in m.state_dict().items()}


def grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def tri_case(name, B, K, Q, A, vd, qd, h_mm, rank, G, seed, out):
    torch.manual_seed(seed)
    att = TriAttention(vd, qd, qd, h_mm, 1, rank, G, 1).eval()
    pools = [TCNet(vd, qd, qd, h_mm, 1, rank, 1, k=2).eval() for _ in range(G)]
    v, q, a = inputs(B, K, Q, A, vd, qd, seed + 1)
    q.requires_grad_(True)
    a.requires_grad_(True)
    p, logits = att(v, q, a)
    pooled = [pools[g].forward_with_weights(v, q, a, p[:, :, :, :, g]) for g in range(G)]
    gen = torch.Generator().manual_seed(seed + 2)
    cot = [torch.randn(pooled[0].shape, generator=gen) for _ in range(G)]
    loss = sum((o * c).sum() for o, c in zip(pooled, cot))
    loss.backward()
    out[name] = {
        "cfg": dict(B=B, K=K, Q=Q, A=A, v_dim=vd, q_dim=qd, h_mm=h_mm, rank=rank, G=G,
                    k_pool=2),
        "v": v, "q": q.detach().clone(), "a": a.detach().clone(),
        "att_sd": sd(att), "pool_sd": [sd(m) for m in pools],
        "p": p.detach().clone(), "logits": logits.detach().clone().contiguous(),
        "pooled": [o.detach().clone() for o in pooled], "cot": cot,
        "dq": q.grad.clone(), "da": a.grad.clone(),
        "att_grads": grads(att), "pool_grads": [grads(m) for m in pools],
    }


def bi_case(name, B, K, Q, vd, qd, hid, G, seed, out):
    torch.manual_seed(seed)
    att = BiAttention(vd, qd, hid, G).eval()
    pools = [BCNet(vd, qd, hid, None, k=1).eval() for _ in range(G)]
    v, q, _ = inputs(B, K, Q, 1, vd, qd, seed + 1)
    q.requires_grad_(True)
    p, logits = att.forward_all(v, q)
    pooled = [pools[g].forward_with_weights(v, q, p[:, g]) for g in range(G)]
    gen = torch.Generator().manual_seed(seed + 2)
    cot = [torch.randn(pooled[0].shape, generator=gen) for _ in range(G)]
    loss = sum((o * c).sum() for o, c in zip(pooled, cot))
    loss.backward()
    out[name] = {
        "cfg": dict(B=B, K=K, Q=Q, v_dim=vd, q_dim=qd, hid=hid, G=G),
        "v": v, "q": q.detach().clone(),
        "att_sd": sd(att), "pool_sd": [sd(m) for m in pools],
        "p": p.detach().clone(), "logits": logits.detach().clone(),
        "pooled": [o.detach().clone() for o in pooled], "cot": cot,
        "dq": q.grad.clone(),
        "att_grads": grads(att), "pool_grads": [grads(m) for m in pools],
    }


def main():
    out = {}
    # Kolda & Bader n-mode product example held in src/Tensor.py:31-32
    X = torch.tensor([[[1, 13], [4, 16], [7, 19], [10, 22]], [[2, 14], [5, 17], [8, 20], [11, 23]],
                      [[3, 15], [6, 18], [9, 21], [12, 24]]], dtype=torch.float32)
    U1 = torch.tensor([[1, 3, 5], [2, 4, 6]], dtype=torch.float32).unsqueeze(0)
    Y = RefTensor.ModeProduct(X.unsqueeze(0).unsqueeze(4), U1, torch.eye(4).unsqueeze(0),
                              torch.eye(2).unsqueeze(0), None)
    out["kolda_bader"] = {"X": X, "U1": U1, "Y": Y.detach().clone().contiguous()}
    # FCNet with and without activation / dropout module
    torch.manual_seed(7)
    for nm, act, dr in (("fc_relu", "ReLU", 0.2), ("fc_lin", "", 0.2), ("fc_nodrop", "ReLU", 0)):
        m = FCNet([24, 40], act, dr).eval()
        x = torch.randn(5, 3, 24, requires_grad=True)
        y = m(x)
        c = torch.randn(y.shape)
        (y * c).sum().backward()
        out[nm] = {"sd": sd(m), "x": x.detach().clone(), "y": y.detach().clone(), "cot": c,
                   "dx": x.grad.clone(), "grads": grads(m), "act": act, "dropout": dr}
    tri_case("tri_small_g2", 3, 5, 3, 2, 32, 16, 16, 4, 2, 11, out)
    tri_case("tri_small_g3", 2, 7, 4, 3, 32, 16, 16, 4, 3, 21, out)
    tri_case("tri_d16", 2, 10, 12, 6, 64, 48, 64, 4, 2, 31, out)   # d = 16 like the real model
    bi_case("bi_small", 3, 6, 4, 32, 16, 24, 2, 41, out)
    bi_case("bi_c128", 3, 10, 12, 64, 48, 128, 2, 51, out)        # channel counts the CUDA kernels tile (x128)
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
