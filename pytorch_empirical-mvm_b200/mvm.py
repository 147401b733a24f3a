"""MVM (masked visual modelling) pieces on either side of the Swin encoder (reference main_pretrain.py).

* ``sample_block_masks``  -- the "bm" blockwise mask sampler of ``Agent_Pretrain.masking`` (main_pretrain.py:309-321) as
  vectorised host code: the same ``np.random.randint`` calls in the same order (so a seeded run draws the same blocks), the
  python set / triple loop replaced by slice assignment into a (T,h,w) coverage grid.
* ``apply_block_mask``    -- ``img[i] *= 1 - cov`` and the full-resolution ``mvm_mask`` (main_pretrain.py:355-362): one CUDA
  pass over the clip (``vsw_block_mask_apply``); the canonical mask stays the (B,T,h,w) uint8 coverage grid.
* ``mvm_3d_feature_loss`` -- the ``3d_feature`` target (main_pretrain.py:508-524): masked L1 between the prediction and the
  TEACHER Swin's output tokens, read straight from the channels-last buffer the encoder returns a view of (no permute
  copy); ``max_pool2d(mvm_mask, 32).sum(1) / 3`` is the coverage grid itself.

CUDA only, no fallback (the sampler is host integer code, as in the reference).
"""
from __future__ import annotations

import numpy as np
import torch

from . import functional as VF

__all__ = ["sample_block_masks", "apply_block_mask", "mvm_3d_feature_loss"]


def sample_block_masks(B: int, T: int, h: int, w: int, rng=np.random) -> np.ndarray:
    """(B,T,h,w) uint8 coverage: per sample, T random (t,h,w) blocks (main_pretrain.py:312-321).  ``rng`` needs
    ``randint(low, high)`` with numpy's half-open semantics; the default is the global numpy stream the reference uses."""
    cov = np.zeros((B, T, h, w), dtype=np.uint8)
    for i in range(B):
        for _ in range(T):
            bt = rng.randint(1, T) if T > 1 else 1
            bh = rng.randint(1, h * 2 // 3)
            bw = rng.randint(1, w * 2 // 3)
            t1 = rng.randint(0, T - bt + 1)
            h1 = rng.randint(0, h - bh + 1)
            w1 = rng.randint(0, w - bw + 1)
            cov[i, t1:t1 + bt, h1:h1 + bh, w1:w1 + bw] = 1
    return cov


def apply_block_mask(img: torch.Tensor, cov, patch_size: int = 32, inplace: bool = False, want_mask: bool = True):
    """img (B,T,3,H,W) on the GPU, cov (B,T,h,w) {0,1} (numpy or tensor).  Returns (masked clip, mvm_mask fp32 or None)."""
    cov_t = torch.as_tensor(cov).to(device=img.device, dtype=torch.uint8)
    return VF.block_mask_apply(img, cov_t, patch_size, inplace=inplace, want_mask=want_mask)


def mvm_3d_feature_loss(pred: torch.Tensor, teacher_out: torch.Tensor, cov, in_c: int = 3) -> torch.Tensor:
    """pred (B,T,h*w,C): ``fc_mvm`` of the non-class fusion outputs; teacher_out (B,C,T,h,w): the teacher
    ``SwinTransformer3D`` output as returned (a permuted view of the (B,T,h,w,C) buffer); cov (B,T,h,w) {0,1}.
    Returns the fp32 scalar ``sum(|pred - target| * m) / (sum(m) + 1e-5) / in_c`` (main_pretrain.py:520-522)."""
    B, Tn, hw, C = pred.shape
    target = teacher_out.detach().permute(0, 2, 3, 4, 1)          # (B,T,h,w,C): contiguous for our encoder's output
    if not target.is_contiguous():
        target = target.contiguous()
    if target.dtype not in (pred.dtype, torch.float32):
        target = target.to(pred.dtype)
    m = torch.as_tensor(cov).to(device=pred.device, dtype=torch.float32).reshape(B * Tn * hw)
    return VF.masked_l1(pred.reshape(B * Tn * hw, C), target.reshape(B * Tn * hw, C), m, float(in_c))
