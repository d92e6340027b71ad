"""Loader-side wire formats -- host code, CPU only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cti_b200  # noqa: E402
from cti_b200.loader import _to_bf16_bits  # noqa: E402


def _reference_store(n_images, v_dim, seed):
    """Arrays laid out like the reference's HDF5 feature store (tools/adaptive_detection_features_converter.py:9-15)."""
    rng = np.random.default_rng(seed)
    n_boxes = rng.integers(10, 101, n_images)                         # adaptive 10-100 boxes per image
    pos = np.zeros((n_images, 2), dtype=np.int64)
    pos[:, 1] = np.cumsum(n_boxes)
    pos[1:, 0] = pos[:-1, 1]
    feats = np.maximum(rng.standard_normal((int(pos[-1, 1]), v_dim)).astype(np.float32), 0)
    feats[pos[3, 0] + 2] = 0                                          # a genuine all-zero box
    return feats, pos


def test_bf16_bits_match_torch_rounding():
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 7
    x[:4] = [0.0, -0.0, 1.0000001, 3.3895314e38]
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(_to_bf16_bits(x), want)


def test_feature_store_equals_trim_collate_padding(tmp_path):
    feats, pos = _reference_store(12, 64, 1)
    store = cti_b200.FeatureStoreBF16.from_reference_arrays(feats, pos, max_boxes=50)
    assert len(store) == 12 and store.feats.shape == (12, 50, 64)
    store.save(str(tmp_path / "s"))
    store = cti_b200.FeatureStoreBF16.load(str(tmp_path / "s"))
    idx = [5, 0, 3, 11]
    # what the reference feeds the model: first <= 50 boxes of each image (src/MC/dataset.py:252-256), zero-padded to the
    # longest of the batch (trim_collate, src/utils.py:127-136) -- here always padded to max_boxes, which only adds zero rows
    batch = []
    for i in idx:
        rows = torch.from_numpy(feats[pos[i, 0]:min(pos[i, 1], pos[i, 0] + 50)])
        batch.append(torch.nn.functional.pad(rows, (0, 0, 0, 50 - rows.shape[0])))
    want = torch.stack(batch, 0)
    fb = cti_b200.FeatureBatch.__new__(cti_b200.FeatureBatch)         # no pinned memory needed on a CPU-only box
    fb.host = torch.empty((4, 50, 64), dtype=torch.bfloat16)
    fb.host_mask = torch.empty((4, 50), dtype=torch.uint8)
    fb.fill(store, idx)
    assert torch.equal(fb.host, want.to(torch.bfloat16))
    assert torch.equal(fb.host_mask.bool(), want.abs().sum(2) == 0)   # the mask of src/attention.py:55
    assert fb.host_mask[2, 2] == 1 and fb.host_mask[2, 1] == 0        # the genuine zero box of image 3


def test_teacher_logits_matrix_equals_reference_dict():
    rng = np.random.default_rng(2)
    d = {int(q): np.float16(rng.standard_normal(37)) for q in rng.permutation(1000)[:50]}     # make_json_with_logits
    t = cti_b200.TeacherLogits.from_reference_dict(d)
    assert t.logits.dtype == torch.float16 and t.logits.shape == (50, 37)
    ask = list(d)[:7][::-1]
    got = t.batch(ask)
    for row, q in zip(got, ask):
        assert np.array_equal(row.numpy(), d[q])
