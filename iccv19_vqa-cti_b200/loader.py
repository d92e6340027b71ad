"""Loader-side wire formats for the hot path.

The reference keeps the bottom-up image features as an HDF5 store -- ``image_features`` (total boxes, 2048) float32 plus
``pos_boxes`` (images, 2) start / end rows (tools/adaptive_detection_features_converter.py:9-15) -- slices a variable
number of boxes per image (src/MC/dataset.py:252-256, capped at ``max_boxes``), pads every batch to its longest image
with zero rows in ``trim_collate`` (src/utils.py:127-136) and ships fp32 to the GPU, where TriAttention re-derives the
padding mask from the zero rows (src/attention.py:55).  Teacher logits for distillation travel as a pickled dict
``{question id: float16 array}`` (src/FFOE/test.py:125-130,184-187).

Here the same information is laid out for the device once, at conversion time:

* ``FeatureStoreBF16`` -- every image padded to ``max_boxes`` rows, **bf16**, one contiguous (images, K, v_dim) array
  (memory-mappable) plus the zero-row mask (images, K) uint8.  A batch is a gather of whole images: no per-batch
  padding, half the host-to-device bytes, and neither the fp32 -> bf16 cast nor the mask pass runs on the device.
* ``FeatureBatch`` -- a pinned host batch and its device twin; ``to_device`` copies features + mask and primes the
  modules' feature cache (``fc.cast_features``), so ``TriAttention`` / ``TCNet`` / ``BCNet`` take the bf16 tensor as their
  ordinary ``v`` argument.  bf16 rounding on the host is the same round-to-nearest-even the device cast performs: the
  modules' outputs are bit-identical to the fp32 path.
* ``TeacherLogits`` -- the per-question dict as ONE fp16 matrix + an index, which ``Distillation_Loss`` consumes as is.

Host code (numpy / torch CPU); the device work is two plain copies.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from . import fc as _fc


def _to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (uint16), round to nearest even -- what ``cti_cast_rows_mask`` does on the device."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    out = (rounded >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.uint16(0x7FC0)
    return out


class FeatureStoreBF16:
    """(images, max_boxes, v_dim) bf16 features + (images, max_boxes) zero-row mask (1 = padding / all-zero row)."""

    def __init__(self, feats_u16: np.ndarray, mask: np.ndarray):
        assert feats_u16.dtype == np.uint16 and feats_u16.ndim == 3 and mask.shape == feats_u16.shape[:2]
        self.feats = feats_u16
        self.mask = mask

    @classmethod
    def from_reference_arrays(cls, image_features: np.ndarray, pos_boxes: np.ndarray, max_boxes: int = 50,
                              chunk: int = 256) -> "FeatureStoreBF16":
        """image_features (total boxes, v_dim) float32 and pos_boxes (images, 2) as in the reference's HDF5 store.
        Images with more than ``max_boxes`` boxes keep their first ``max_boxes`` (src/MC/dataset.py:252-256)."""
        n, d = int(pos_boxes.shape[0]), int(image_features.shape[1])
        feats = np.zeros((n, max_boxes, d), dtype=np.uint16)
        mask = np.ones((n, max_boxes), dtype=np.uint8)
        for lo in range(0, n, chunk):
            for i in range(lo, min(n, lo + chunk)):
                s, e = int(pos_boxes[i, 0]), int(pos_boxes[i, 1])
                e = min(e, s + max_boxes)
                rows = np.asarray(image_features[s:e], dtype=np.float32)
                feats[i, :e - s] = _to_bf16_bits(rows)
                mask[i, :e - s] = (np.abs(rows).sum(1) == 0).astype(np.uint8)      # a genuine all-zero box is masked too
        return cls(feats, mask)

    def save(self, path: str) -> None:
        np.save(path + ".feat.npy", self.feats)
        np.save(path + ".mask.npy", self.mask)

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "FeatureStoreBF16":
        return cls(np.load(path + ".feat.npy", mmap_mode="r" if mmap else None), np.load(path + ".mask.npy"))

    def __len__(self) -> int:
        return self.feats.shape[0]


class FeatureBatch:
    """Pinned host buffers for one batch of ``FeatureStoreBF16`` images and (optionally) a persistent device twin."""

    def __init__(self, batch: int, max_boxes: int, v_dim: int, device: Optional[torch.device] = None):
        self.host = torch.empty((batch, max_boxes, v_dim), dtype=torch.bfloat16).pin_memory()
        self.host_mask = torch.empty((batch, max_boxes), dtype=torch.uint8).pin_memory()
        self.dev = self.dev_mask = None
        if device is not None:
            self.dev = torch.empty((batch, max_boxes, v_dim), dtype=torch.bfloat16, device=device)
            self.dev_mask = torch.empty((batch * max_boxes,), dtype=torch.uint8, device=device)

    def fill(self, store: FeatureStoreBF16, indices: Sequence[int]) -> "FeatureBatch":
        """Gather whole images into the pinned buffers (the collate step: no padding work, no dtype conversion)."""
        idx = np.asarray(indices)
        self.host.view(torch.int16).numpy()[:] = store.feats[idx].view(np.int16)
        self.host_mask.numpy()[:] = store.mask[idx]
        return self

    def to_device(self, non_blocking: bool = True):
        """Copy features + mask to the device twin and prime the modules' feature cache.  Returns the bf16 tensor to
        pass as ``v``."""
        self.dev.copy_(self.host, non_blocking=non_blocking)
        self.dev_mask.copy_(self.host_mask.view(-1), non_blocking=non_blocking)
        return prime_features(self.dev, self.dev_mask)


def prime_features(v_bf16: torch.Tensor, rowmask: torch.Tensor) -> torch.Tensor:
    """Attach the loader's zero-row mask to a device bf16 feature tensor (batch, K, v_dim): the modules then use the
    tensor as is -- no cast, no mask pass (``fc.cast_features`` finds the entry in its per-tensor cache)."""
    if v_bf16.dtype != torch.bfloat16 or v_bf16.dim() != 3 or not v_bf16.is_contiguous():
        raise RuntimeError("prime_features: expected a contiguous bf16 (batch, regions, dim) tensor")
    if rowmask.dtype != torch.uint8 or rowmask.numel() != v_bf16.shape[0] * v_bf16.shape[1]:
        raise RuntimeError("prime_features: the mask must be uint8 with one entry per (sample, region)")
    key = (v_bf16._version, v_bf16.data_ptr())
    setattr(v_bf16, _fc._FEAT_ATTR, (key, v_bf16.view(-1, v_bf16.shape[-1]), rowmask.reshape(-1)))
    return v_bf16


class TeacherLogits:
    """The teacher's class logits as ONE fp16 matrix (questions, classes) + a question-id index, instead of the
    reference's pickled ``{qid: float16 array}`` (src/FFOE/test.py:125-130)."""

    def __init__(self, logits_f16: torch.Tensor, qids: Iterable[int]):
        assert logits_f16.dtype == torch.float16 and logits_f16.dim() == 2
        self.logits = logits_f16
        self.qids = [int(q) for q in qids]
        self.row: Dict[int, int] = {q: i for i, q in enumerate(self.qids)}

    @classmethod
    def from_reference_dict(cls, d: Dict[int, np.ndarray]) -> "TeacherLogits":
        qids = sorted(int(k) for k in d)
        mat = np.stack([np.asarray(d[q], dtype=np.float16) for q in qids], 0)
        return cls(torch.from_numpy(mat), qids)

    def batch(self, qids: Iterable[int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp16 (len(qids), classes) rows in the order asked for (into a pinned ``out`` if given)."""
        rows = torch.as_tensor([self.row[int(q)] for q in qids])
        if out is None:
            return self.logits.index_select(0, rows)
        torch.index_select(self.logits, 0, rows, out=out)
        return out
