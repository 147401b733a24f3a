"""CPU oracle for the MVM masking / 3d_feature loss around the encoder -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in plain numpy / torch on the CPU, three pieces of the reference's ``Agent_Pretrain`` (main_pretrain.py,
tsujuifu/pytorch_empirical-mvm):
  block_cells        the "bm" block sampler, main_pretrain.py:309-321 (same np.random.randint calls, same order)
  apply_block_mask   coverage grid -> pixel zeroing + full-resolution mvm_mask, main_pretrain.py:349-362
  feature_loss       the ``3d_feature`` MVM loss, main_pretrain.py:508-524
Only ``tests/`` may import it, as the checker.  It is a restatement, not a copy: masking is expressed on a (T,h,w)
coverage grid with ``repeat_interleave`` instead of the reference's 6-D ``expand`` / ``flatten``; the loss takes the
coverage grid directly (``max_pool2d(mvm_mask, ps).sum(1) / 3`` of an expanded {0,1} grid is the grid).

Parity status: PINNED against the unmodified reference methods.  ``main_pretrain.py`` cannot be imported (utils/lib.py needs
easydict / skimage / fairscale / toolz), so ``tests/golden/make_golden_mvm.py`` cuts the ``masking`` and
``calc_mvm_loss`` method definitions out of the file with ``ast`` and runs exactly that source against a stub ``self``; the
inputs, RNG seeds and outputs live in ``tests/golden/mvm.pt`` and ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def block_cells(T: int, h: int, w: int, rng=np.random):
    """one sample's masked cells as the reference builds them: T blocks, each a set of (t,h,w) tuples (:312-321)"""
    cells = set()
    for _ in range(T):
        bt = rng.randint(1, T) if T > 1 else 1
        bh = rng.randint(1, h * 2 // 3)
        bw = rng.randint(1, w * 2 // 3)
        t1, h1, w1 = rng.randint(0, T - bt + 1), rng.randint(0, h - bh + 1), rng.randint(0, w - bw + 1)
        for it in range(t1, t1 + bt):
            for ih in range(h1, h1 + bh):
                for iw in range(w1, w1 + bw):
                    cells.add((it, ih, iw))
    return cells


def cover_grid(cells, T: int, h: int, w: int) -> torch.Tensor:
    cov = torch.zeros(T, h, w)
    for it, ih, iw in cells:
        cov[it, ih, iw] = 1.0                                                      # :350-352
    return cov


def apply_block_mask(img: torch.Tensor, cov: torch.Tensor, ps: int = 32):
    """img (B,T,Cin,H,W), cov (B,T,h,w) float {0,1} -> (img * (1 - mask), mask (B,T,Cin,H,W) float)   (:355-362)"""
    full = cov.repeat_interleave(ps, dim=-2).repeat_interleave(ps, dim=-1)        # (B,T,H,W)
    mask = full.unsqueeze(2).expand(-1, -1, img.shape[2], -1, -1).contiguous()
    return img * (1.0 - mask).to(img.dtype), mask


def feature_loss(pred: torch.Tensor, teacher_out: torch.Tensor, cov: torch.Tensor, in_c: int = 3) -> torch.Tensor:
    """pred (B,T,hw,C); teacher_out (B,C,T,h,w) as the Swin returns it; cov (B,T,h,w)   (:516-522)"""
    B, Tn, hw, C = pred.shape
    target = teacher_out.permute(0, 2, 3, 4, 1).reshape(B, Tn, hw, C)
    m = cov.reshape(B, Tn, hw, 1).float()
    return (F.l1_loss(pred, target, reduction="none").float() * m).sum() / (m.sum() + 1e-5) / in_c
