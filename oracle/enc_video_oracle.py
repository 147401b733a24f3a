"""CPU oracle for the EncVideo tail -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, as one stateless function over plain tensors, what the reference's ``EncVideo.forward`` (model.py:32-78,
tsujuifu/pytorch_empirical-mvm) does AFTER the Swin backbone: token view of the backbone output (model.py:39-40), ``fc``
(:42), class row (:57), position embedding (:58), frame-length / frame-order embedding (:60-67), LayerNorm (:69) and
the token mask (:71-76).  Only ``tests/`` and ``__graft_entry__.smoke()`` may import it, as the checker.

Not a copy: the reference builds the result with ``cat`` / ``expand`` / per-sample python loops over in-place adds; this
restatement writes the closed form ``pre[b,t,p] = (p == 0 ? cls : fc(tok)[b,t,p-1]) + pos[p] + (odr[b,t] == t ? len[t] :
odr_emb)`` with index tensors, which is also the form the CUDA kernel implements (csrc/enc_video.cu).

Parity status: PINNED against the unmodified reference class.  ``model.py`` cannot be imported here (it star-imports
utils/lib.py, which needs easydict / skimage / fairscale / toolz -- not installed), so
``tests/golden/make_golden_enc_video.py`` extracts the ``EncVideo`` class statement from /root/reference/model.py with
``ast``, executes exactly that source with ``T = torch`` and a stand-in backbone, and stores inputs, outputs and
gradients in ``tests/golden/enc_video.pt``; ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def enc_video_tail(swin_out: torch.Tensor, p: Dict[str, torch.Tensor], odr=None, vt_mask=None, eps: float = 1e-5):
    """swin_out (B, L, T, h, w): the backbone output exactly as ``SwinTransformer3D.forward`` returns it.
    p: ``fc.weight`` (hidden, L) / ``fc.bias`` (optional pair), ``emb_cls`` (1,1,1,hidden), ``emb_pos``
    (1,1,1+max_patch^2,hidden), ``emb_len`` (1,max_frame,1,hidden), ``emb_odr`` (1,1,1,hidden), ``norm.weight``,
    ``norm.bias``.  odr: (B,T) ints or None.  vt_mask: broadcastable to (B,T,1+h*w) or None.
    Returns (f_img (B, T*(1+h*w), hidden), m_img (B, T*(1+h*w)) int64)."""
    B, L, T, h, w = swin_out.shape
    hw, P = h * w, 1 + h * w
    tok = swin_out.permute(0, 2, 3, 4, 1).reshape(B, T, hw, L)                 # model.py:39-40
    if "fc.weight" in p:
        tok = F.linear(tok, p["fc.weight"], p.get("fc.bias"))                   # model.py:42
    hid = tok.shape[-1]
    dt = torch.promote_types(tok.dtype, p["emb_cls"].dtype)
    pre = torch.empty(B, T, P, hid, dtype=dt)
    pre[:, :, 0] = p["emb_cls"].reshape(hid)                                    # model.py:57
    pre[:, :, 1:] = tok
    pre = pre + p["emb_pos"].reshape(-1, hid)[:P].view(1, 1, P, hid)            # model.py:58
    emb_len = p["emb_len"].reshape(-1, hid)
    if emb_len.shape[0] < T:
        raise ValueError(f"emb_len holds {emb_len.shape[0]} frames, clip has {T}")   # the reference's add fails to broadcast
    frame = emb_len[:T].unsqueeze(0).expand(B, T, hid)                          # model.py:67
    if odr is not None:                                                         # model.py:60-66
        keep = torch.as_tensor(odr).reshape(B, T) == torch.arange(T).view(1, T)
        frame = torch.where(keep.unsqueeze(-1), frame, p["emb_odr"].reshape(1, 1, hid).expand(B, T, hid))
    pre = pre + frame.unsqueeze(2)
    out = F.layer_norm(pre, (hid,), p["norm.weight"], p["norm.bias"], eps).reshape(B, T * P, hid)   # model.py:69
    m = torch.ones(B, T, P, dtype=torch.int64)                                  # model.py:71-73
    if vt_mask is not None:
        m = m * torch.as_tensor(vt_mask)                                        # model.py:75
    return out, m.reshape(B, T * P)
