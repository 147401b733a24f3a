"""EncVideo -- the VIOLET video encoder head that consumes the Swin output every step (reference model.py:7-78).

Drop-in for the reference class: same constructor ``EncVideo(args, hidden_size)``, same parameter names
(``swin.*``, ``fc.{weight,bias}``, ``emb_cls``, ``emb_pos``, ``emb_len``, ``emb_odr``, ``norm.{weight,bias}``; VIOLET
checkpoints store them under ``enc_img.*``, model.py:363) and shapes, same ``forward(img, odr=None, vt_mask=None) ->
(f_img (B, T*(1+h*w), hidden), m_img (B, T*(1+h*w)) int64)``.

Everything after the backbone is two launches through the C ABI (include/vsw.h): ``fc`` is a ``vsw_linear_fwd`` on the
channels-last Swin tokens (no permute copy: the backbone already returns a view of that buffer) and the class row +
embedding adds + LayerNorm + mask are ``vsw_enc_video_tail_fwd``.  CUDA only, no fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as VF
from .video_swin import _cast, _compute_dtype, _require_cuda, get_vidswin_model

__all__ = ["EncVideo"]


class EncVideo(nn.Module):
    def __init__(self, args, hidden_size, swin: nn.Module = None):
        """``swin``: optional pre-built backbone (otherwise ``get_vidswin_model(args)``, model.py:11)."""
        super().__init__()
        self.swin = swin if swin is not None else get_vidswin_model(args)
        self.latent_feat_size = self.swin.norm.normalized_shape[0]
        self.img_feature_dim = hidden_size
        self.swinbert = getattr(args, "swinbert", False)
        self.max_size_frame = getattr(args, "max_size_frame", 6)
        self.max_size_patch = getattr(args, "max_size_patch", 14)
        if self.swinbert:
            raise NotImplementedError("EncVideo(swinbert=True) (model.py:27-29, 44-55) is not part of the B200 path")
        if self.latent_feat_size != self.img_feature_dim:
            self.fc = nn.Linear(self.latent_feat_size, self.img_feature_dim)
        else:
            self.fc = None
        d = self.img_feature_dim
        self.emb_cls = nn.Parameter(0.02 * torch.randn(1, 1, 1, d))
        self.emb_pos = nn.Parameter(0.02 * torch.randn(1, 1, 1 + self.max_size_patch ** 2, d))
        self.emb_len = nn.Parameter(0.02 * torch.randn(1, self.max_size_frame, 1, d))
        self.emb_odr = nn.Parameter(0.02 * torch.randn(1, 1, 1, d))
        self.norm = nn.LayerNorm(d)
        self.transform_normalize = None

    def forward(self, img, odr=None, vt_mask=None):
        _require_cuda(img)
        _B, _T, _C, _H, _W = img.shape
        _h, _w = _H // 32, _W // 32
        if self.transform_normalize is not None:
            img = self.transform_normalize(img)
        with torch.cuda.device(img.device):
            y = self.swin(img.transpose(1, 2))                       # (B, 8E, T, h, w): a view of the (B,T,h,w,8E) buffer
            tok = y.permute(0, 2, 3, 4, 1).view([_B, _T, _h * _w, self.latent_feat_size])   # model.py:39-40, no copy
            cd = _compute_dtype(self.norm.weight)
            tok = _cast(tok, cd)
            if self.fc is not None:
                tok = VF.linear(tok.view(-1, self.latent_feat_size), _cast(self.fc.weight, cd),
                                _cast(self.fc.bias, cd)).view(_B, _T, _h * _w, self.img_feature_dim)
            odr_t = None
            if odr is not None:   # model.py:60-66: frame i keeps emb_len[i] iff odr[b][i] == i, else emb_odr
                odr_t = torch.as_tensor(odr, device=img.device).to(torch.int32).reshape(_B, _T).contiguous()
            vt = None
            if vt_mask is not None:   # model.py:75 (broadcast multiply into the all-ones mask)
                vt = torch.broadcast_to(torch.as_tensor(vt_mask, device=img.device).to(torch.int64),
                                        (_B, _T, 1 + _h * _w)).contiguous()
            # the cat with the fp32 emb_cls promotes to fp32 under autocast, and LayerNorm returns fp32 there
            out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else cd
            f_img, m_img = VF.enc_video_tail(tok, self.emb_cls, self.emb_pos, self.emb_len, self.emb_odr,
                                             self.norm.weight, self.norm.bias, odr_t, vt, out_dtype)
        return f_img, m_img
