"""The loader wire format on the device: bf16 features + zero-row mask in, bit-identical results out."""
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


def test_bf16_features_give_bit_identical_results_and_skip_the_cast():
    torch.manual_seed(3)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, 2, 1).to(DEV).eval()
    pool = cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2).to(DEV).eval()
    v, q, a = (t.to(DEV) for t in O.synthetic_inputs(6, 50, 12, 6, seed=5))

    def run(vv):
        qq = q.clone().requires_grad_(True)
        p, logits = att(vv, qq, a)
        out = pool.forward_with_weights(vv, qq, a, p[:, :, :, :, 0])
        out.square().sum().backward()
        return p.detach(), logits.detach(), out.detach(), qq.grad
    ref = run(v)
    v16 = v.to(torch.bfloat16)
    mask = (v.abs().sum(2) == 0).to(torch.uint8).reshape(-1)
    assert torch.equal(K_.rowmask_bf16(v16.view(-1, 2048)), mask)            # the device-side mask pass
    K_.STATS.launches = 0
    got_unprimed = run(v16.clone())                                          # mask computed on the device
    n_unprimed = K_.STATS.launches
    K_.STATS.launches = 0
    got = run(cti_b200.prime_features(v16, mask))                            # mask from the loader
    assert K_.STATS.launches == n_unprimed - 1
    for x, y, z in zip(ref, got, got_unprimed):
        assert torch.equal(x, y) and torch.equal(x, z)
    # the pinned batch object: gather from a store, copy, prime
    import numpy as np
    feats = v.cpu().numpy().reshape(-1, 2048)
    pos = np.stack([np.arange(6) * 50, np.arange(6) * 50 + 50], 1)
    store = cti_b200.FeatureStoreBF16.from_reference_arrays(feats, pos, max_boxes=50)
    fb = cti_b200.FeatureBatch(6, 50, 2048, torch.device(DEV)).fill(store, list(range(6)))
    vv = fb.to_device()
    assert torch.equal(vv, v16)
    for x, y in zip(ref, run(vv)):
        assert torch.equal(x, y)
    # train mode: the dropout kernels take bf16 features too
    att.train()
    p, _ = att(vv, q, a)
    assert torch.isfinite(p).all()


def test_bf16_token_embeddings_skip_the_cast_and_get_bf16_gradients():
    """Question / answer token embeddings that already travel as bf16 are consumed as they are (no cast launch); the
    forward is bit-identical to the fp32 call on the same (bf16-representable) values and dq / da come back in bf16,
    equal to the rounded fp32 gradients."""
    torch.manual_seed(4)
    att = cti_b200.TriAttention(2048, 1024, 1024, 512, 1, 32, 2, 1).to(DEV).eval()
    pool = cti_b200.TCNet(2048, 1024, 1024, 512, 1, 32, 1, k=2).to(DEV).eval()
    v, q, a = (t.to(DEV) for t in O.synthetic_inputs(5, 50, 12, 6, seed=6))
    q16, a16 = q.to(torch.bfloat16), a.to(torch.bfloat16)

    def run(qq, aa):
        qq, aa = qq.clone().requires_grad_(True), aa.clone().requires_grad_(True)
        K_.STATS.launches = 0
        p, logits = att(v, qq, aa)
        out = pool.forward_with_weights(v, qq, aa, p[:, :, :, :, 0])
        n = K_.STATS.launches
        out.square().sum().backward()
        return (p.detach(), logits.detach(), out.detach()), (qq.grad, aa.grad), n
    run(q16.float(), a16.float())                           # builds the lazily cached weight packs
    ref, gref, n32 = run(q16.float(), a16.float())
    got, g16, n16 = run(q16, a16)
    assert n16 == n32 - 2                                   # the two token casts are gone
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
    for g, gr in zip(g16, gref):
        assert g.dtype == torch.bfloat16
        assert (g.float() - gr).abs().max().item() <= 1e-2 * gr.abs().max().item()
