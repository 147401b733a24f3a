"""Drop-in B200-native replacement for the reference's ``visbackbone/video_swin.py``.

Same public surface (constructor arguments, ``forward`` signatures, ``state_dict`` keys, helper
functions) as tsujuifu/pytorch_empirical-mvm ``visbackbone/video_swin.py`` -- citations in the
docstrings are file:line of that file -- but every tensor op of the hot path runs in the
hand-written sm_100a kernels of ``libvsw_b200.so`` (see ``include/vsw.h``).  Activations stay
channels-last tokens ``(B, D, H, W, C)`` from the patch embedding to the final norm: there is no
``roll`` / ``window_partition`` / ``permute().contiguous()`` copy anywhere.

CUDA only.  There is no CPU path and no fallback: CPU tensors raise.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.utils.checkpoint as checkpoint

from . import _lib as L
from . import functional as VF
from .functional import get_window_size  # noqa: F401  (re-exported, video_swin.py:95)

__all__ = ["SwinTransformer3D", "SwinTransformerBlock3D", "WindowAttention3D", "Mlp", "PatchMerging", "PatchEmbed3D",
           "BasicLayer", "DropPath", "drop_path", "window_partition", "window_reverse", "get_window_size",
           "compute_mask", "get_vidswin_model", "load_checkpoint_3d", "trunc_normal_"]


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Truncated normal init, same sampling recipe as video_swin.py:18-43 (uniform -> erfinv)."""
    def cdf(v):
        return (1. + math.erf(v / math.sqrt(2.))) / 2.
    with torch.no_grad():
        lo, hi = cdf((a - mean) / std), cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_().mul_(std * math.sqrt(2.)).add_(mean).clamp_(min=a, max=b)
    return tensor


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _compute_dtype(*params) -> torch.dtype:
    """dtype the kernels run in: the autocast dtype when CUDA autocast is on, else the parameters' dtype."""
    if torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    for p in params:
        if p is not None:
            return p.dtype
    return torch.float32


def _autocast_on() -> bool:
    return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") != torch.float32


def _cast(t: Optional[torch.Tensor], dtype):
    if t is None or t.dtype == dtype:
        return t
    return t.to(dtype)


def _require_cuda(x: torch.Tensor):
    if not x.is_cuda:
        raise L.VswError("pytorch_empirical-mvm_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")


def drop_path(x, drop_prob: float = 0., training: bool = False):
    """Stochastic depth (video_swin.py:46-54): x / keep * floor(keep + U[0,1)) per sample."""
    if drop_prob == 0. or not training:
        return x
    return x * _drop_path_scale(x, drop_prob).to(x.dtype).view((x.shape[0],) + (1,) * (x.ndim - 1))


def _drop_path_scale(x, drop_prob: float) -> torch.Tensor:
    """(B,) fp32 factors floor(keep+U)/keep, drawing the same ``torch.rand`` the reference draws."""
    keep = 1 - drop_prob
    u = torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
    if not u.is_cuda or u.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return ((keep + u).floor_().reshape(-1).float() / keep).contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib().vsw_drop_path_scale(L.ptr(u), float(keep), L.ptr(out), x.shape[0], L.dt(u), L.stream()),
                "vsw_drop_path_scale")
    return out


class DropPath(nn.Module):
    """video_swin.py:57-63"""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)

    def scale(self, x) -> Optional[torch.Tensor]:
        if not self.training or not self.drop_prob:
            return None
        return _drop_path_scale(x, self.drop_prob)


def window_partition(x, window_size):
    """(B,D,H,W,C) -> (B*nW, N, C) (video_swin.py:84-88); an index-map gather, not a permute copy."""
    B, D, H, W, C = x.shape
    plan = VF.window_plan((D, H, W), tuple(window_size), (0, 0, 0), x.device)
    return x.reshape(B, D * H * W, C)[:, plan.gather.long()].reshape(B * plan.nW, plan.N, C)


def window_reverse(windows, window_size, B, D, H, W):
    """inverse of window_partition (video_swin.py:90-93)"""
    plan = VF.window_plan((D, H, W), tuple(window_size), (0, 0, 0), windows.device)
    C = windows.shape[-1]
    out = torch.empty(B, D * H * W, C, dtype=windows.dtype, device=windows.device)
    out[:, plan.gather.long()] = windows.reshape(B, plan.nW * plan.N, C)
    return out.view(B, D, H, W, C)


def compute_mask(D, H, W, window_size, shift_size, device):
    """(nW,N,N) fp32 additive mask of {0,-100} (video_swin.py:292-307), built on the GPU from the
    closed-form region ids.  The blocks themselves never read this dense tensor: they consume the
    (nW,N) uint8 region ids and compare them in registers."""
    device = torch.device(device)
    if device.type != "cuda":
        raise L.VswError("compute_mask: CUDA only")
    ws, ss = tuple(window_size), tuple(shift_size)
    plan = VF.window_plan((D, H, W), ws, ss, device)
    if plan.region is None:
        return torch.zeros(plan.nW, plan.N, plan.N, device=device)
    return VF.shift_mask_from_region(plan.region, plan.nW, plan.N)


class _RegionMask:
    """What BasicLayer hands to its blocks instead of the dense (nW,N,N) mask tensor: a token that
    says "use the canonical shift mask of this geometry" (the kernels derive it from region ids)."""

    def __init__(self, grid, window, shift):
        self.grid, self.window, self.shift = grid, window, shift

    def to(self, *a, **k):  # BasicLayer casts the mask per block (video_swin.py:363)
        return self


# ----------------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------------
class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (video_swin.py:65-81).  Dropout must be 0 (the reference hard-codes it)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError("vsw Mlp: only nn.GELU (erf) is implemented in the fused epilogue")
        if drop != 0.:
            raise NotImplementedError("vsw Mlp: dropout > 0 is not implemented (reference uses drop=0)")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        _require_cuda(x)
        cd = _compute_dtype(self.fc1.weight)
        shp = x.shape
        y = VF.mlp(_cast(x, cd).reshape(-1, shp[-1]), _cast(self.fc1.weight, cd), _cast(self.fc1.bias, cd),
                   _cast(self.fc2.weight, cd), _cast(self.fc2.bias, cd))
        return y.view(*shp[:-1], y.shape[-1])


class WindowAttention3D(nn.Module):
    """Window multi-head self-attention with relative position bias (video_swin.py:111-172)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("vsw WindowAttention3D: dropout > 0 is not implemented (reference uses 0)")
        self.dim = dim
        self.window_size = tuple(window_size)
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        wd, wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), num_heads))
        # closed form of video_swin.py:123-137 (checked bit-exactly against the reference in tests)
        n = torch.arange(wd * wh * ww)
        d, h, w = n // (wh * ww), (n // ww) % wh, n % ww
        rel = ((d[:, None] - d[None, :] + wd - 1) * ((2 * wh - 1) * (2 * ww - 1))
               + (h[:, None] - h[None, :] + wh - 1) * (2 * ww - 1) + (w[:, None] - w[None, :] + ww - 1))
        self.register_buffer("relative_position_index", rel.to(torch.int64))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self._codes = {}

    def bias_codes(self, N: int):
        """cached (rowcode, colcode) int32 device vectors for the [:N,:N] slice of the index buffer"""
        buf = self.relative_position_index
        key = (N, buf.device, buf._version, buf.data_ptr())
        hit = self._codes.get(key)
        if hit is None:
            self._codes.clear()
            hit = VF.bias_codes(buf, N)
            self._codes[key] = hit
        return hit

    def params(self, cd):
        return (_cast(self.qkv.weight, cd), _cast(self.qkv.bias, cd), _cast(self.relative_position_bias_table, cd),
                _cast(self.proj.weight, cd), _cast(self.proj.bias, cd))

    def forward(self, x, mask=None):
        """x (B_, N, C); mask (nW, N, N) additive or None (video_swin.py:147-172)."""
        _require_cuda(x)
        cd = _compute_dtype(self.qkv.weight)
        B_, N, C = x.shape
        rowcode, colcode = self.bias_codes(N)
        nW = 1
        if mask is not None:
            nW = mask.shape[0]
            mask = _cast(mask, cd).contiguous()
        wq, bq, tab, wp, bp = self.params(cd)
        return VF.window_attention(_cast(x, cd), wq, bq, tab, wp, bp, rowcode, colcode, mask, nW, self.num_heads,
                                   self.scale, cfg_window=self.window_size)


class SwinTransformerBlock3D(nn.Module):
    """video_swin.py:175-263.  ``forward(x[B,D,H,W,C], mask_matrix)``; ``mask_matrix`` may be the dense
    (nW,N,N) tensor the reference passes or the region-id token BasicLayer passes."""

    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("vsw SwinTransformerBlock3D: only nn.LayerNorm is implemented")
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = tuple(window_size)
        self.shift_size = tuple(shift_size)
        self.mlp_ratio = mlp_ratio
        self.use_checkpoint = use_checkpoint
        for a in range(3):
            assert 0 <= self.shift_size[a] < self.window_size[a], "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                      qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _dp_scale(self, x):
        return self.drop_path.scale(x) if isinstance(self.drop_path, DropPath) else None

    def _part1(self, x, mask_matrix, cd):
        B, D, H, W, C = x.shape
        plan = VF.window_plan((D, H, W), self.window_size, self.shift_size, x.device)
        dense = None
        if plan.shifted and isinstance(mask_matrix, torch.Tensor):
            # a caller-supplied dense mask is honoured verbatim (reference semantics, :220-227)
            if tuple(mask_matrix.shape) != (plan.nW, plan.N, plan.N):
                raise ValueError(f"mask_matrix has shape {tuple(mask_matrix.shape)}, this block's windows need "
                                 f"{(plan.nW, plan.N, plan.N)}")
            dense = _cast(mask_matrix, cd).contiguous()
        # BasicLayer passes the _RegionMask token (= the canonical shift mask, derived from region ids in the kernels); an
        # explicit None means NO mask even on a shifted block, as in the reference (video_swin.py:231-235)
        use_region = isinstance(mask_matrix, _RegionMask)
        rowcode, colcode = self.attn.bias_codes(plan.N)
        wq, bq, tab, wp, bp = self.attn.params(cd)
        y = VF.attn_branch(x.view(B, D * H * W, C), _cast(self.norm1.weight, x.dtype), _cast(self.norm1.bias, x.dtype), wq, bq,
                           tab, wp, bp, self._dp_scale(x), plan, rowcode, colcode, dense, self.num_heads,
                           self.attn.scale, cfg_window=self.attn.window_size, use_region=use_region)
        return y.view(B, D, H, W, C)

    def _part2(self, x, cd):
        B, D, H, W, C = x.shape
        m = self.mlp
        y = VF.mlp_branch(x.view(B, D * H * W, C), _cast(self.norm2.weight, x.dtype), _cast(self.norm2.bias, x.dtype),
                          _cast(m.fc1.weight, cd), _cast(m.fc1.bias, cd), _cast(m.fc2.weight, cd),
                          _cast(m.fc2.bias, cd), self._dp_scale(x))
        return y.view(B, D, H, W, C)

    def forward(self, x, mask_matrix=None):
        _require_cuda(x)
        cd = _compute_dtype(self.norm1.weight)
        # Under autocast an fp32 residual stream stays fp32, as in the reference (LayerNorm reads and `shortcut +
        # drop_path(x)` produces fp32 there: stage 0, behind the fp32 patch norm); a 16-bit stream (behind a PatchMerging
        # Linear) stays 16-bit.  Otherwise the stream is in the parameters' dtype.
        if not (_autocast_on() and x.dtype == torch.float32):
            x = _cast(x, cd)
        x = x.contiguous()
        # residual adds and drop-path are fused into the proj / fc2 epilogues (video_swin.py:256, 261)
        if self.use_checkpoint and torch.is_grad_enabled():
            x = checkpoint.checkpoint(self._part1, x, mask_matrix, cd, use_reentrant=False)
            x = checkpoint.checkpoint(self._part2, x, cd, use_reentrant=False)
        else:
            x = self._part1(x, mask_matrix, cd)
            x = self._part2(x, cd)
        return x


class PatchMerging(nn.Module):
    """2x2 spatial merge + LN(4C) + Linear(4C->2C, no bias) (video_swin.py:266-289)."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("vsw PatchMerging: only nn.LayerNorm is implemented")
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        _require_cuda(x)
        cd = _compute_dtype(self.reduction.weight)
        B, D, H, W, C = x.shape
        y = VF.patch_merge(_cast(x, cd).reshape(B, D * H * W, C), _cast(self.norm.weight, cd),
                           _cast(self.norm.bias, cd), _cast(self.reduction.weight, cd), (D, H, W))
        return y.view(B, D, (H + 1) // 2, (W + 1) // 2, 2 * C)


class BasicLayer(nn.Module):
    """One Swin stage (video_swin.py:310-370).  ``forward`` keeps the reference's channels-first
    contract; SwinTransformer3D uses ``forward_tokens`` to stay channels-last."""

    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(
                dim=dim, num_heads=num_heads, window_size=self.window_size,
                shift_size=(0, 0, 0) if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer,
                use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.downsample = downsample
        if self.downsample is not None:
            self.downsample = downsample(dim=dim, norm_layer=norm_layer)

    def forward_tokens(self, x):
        """x (B,D,H,W,C) channels-last -> (B,D,H',W',C') channels-last"""
        B, D, H, W, C = x.shape
        token = _RegionMask((D, H, W), self.window_size, self.shift_size)
        for blk in self.blocks:
            x = blk(x, token)
        if self.downsample is not None:
            x = self.downsample(x)
        return x

    def forward(self, x):
        """x (B,C,D,H,W) -> (B,C',D,H',W') like the reference (video_swin.py:352-370)"""
        y = self.forward_tokens(x.permute(0, 2, 3, 4, 1))
        return y.permute(0, 4, 1, 2, 3)


class PatchEmbed3D(nn.Module):
    """video_swin.py:373-407: pad, Conv3d(k=patch, stride=(1,ph,pw)) after appending one zero frame, LN."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        # a real Conv3d module keeps the reference's parameter names, shapes and default init
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=self.patch_size, stride=(1, 4, 4))
        if norm_layer is not None:
            if norm_layer is not nn.LayerNorm:
                raise NotImplementedError("vsw PatchEmbed3D: only nn.LayerNorm is implemented")
            self.norm = norm_layer(embed_dim)
        else:
            self.norm = None
        if self.patch_size[1:] != (4, 4):
            raise NotImplementedError("vsw PatchEmbed3D: the reference hard-codes spatial stride (4,4)")

    def forward_tokens(self, x):
        """x (B,Cin,D,H,W) -> (B,D',H',W',E) channels-last tokens in the compute dtype"""
        _require_cuda(x)
        cd = _compute_dtype(self.proj.weight)
        B, Cin, D, H, W = x.shape
        pd, ph, pw = self.patch_size
        g = _cast(self.norm.weight, cd) if self.norm is not None else None
        b = _cast(self.norm.bias, cd) if self.norm is not None else None
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x = x.float()
        y = VF.patch_embed(x, _cast(self.proj.weight, cd), _cast(self.proj.bias, cd), g, b, self.patch_size,
                           out_dtype=torch.float32 if (_autocast_on() and self.norm is not None) else None)
        return y.view(B, D + 2 - pd, -(-H // ph), -(-W // pw), self.embed_dim)

    def forward(self, x):
        """(B,Cin,D,H,W) -> (B,E,D,H/4,W/4) like the reference (a permuted view of the token buffer)"""
        return self.forward_tokens(x).permute(0, 4, 1, 2, 3)


class SwinTransformer3D(nn.Module):
    """Video Swin Transformer backbone (video_swin.py:410-570), same constructor and state_dict."""

    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(2, 4, 4), in_chans=3, embed_dim=128,
                 depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=(8, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=True, frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if drop_rate != 0.:
            raise NotImplementedError("vsw SwinTransformer3D: drop_rate > 0 is not implemented (reference uses 0)")
        self.pretrained = pretrained
        self.pretrained2d = pretrained2d
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.frozen_stages = frozen_stages
        self.window_size = tuple(window_size)
        self.patch_size = tuple(patch_size)
        self.patch_embed = PatchEmbed3D(patch_size=self.patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                        norm_layer=norm_layer if self.patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]  # video_swin.py:447
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=self.window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = norm_layer(self.num_features)

    def forward(self, x):
        """x (B,3,D,H,W) -> (B, 8E, D, H/32, W/32), a permuted view of the channels-last buffer
        exactly like the reference returns (video_swin.py:470-482)."""
        _require_cuda(x)
        with torch.cuda.device(x.device):
            t = self.patch_embed.forward_tokens(x)
            for layer in self.layers:
                t = layer.forward_tokens(t)
            B, D, H, W, C = t.shape
            cd = t.dtype
            # F.layer_norm runs in fp32 under autocast, so the reference's output is fp32 there
            out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else cd
            y = VF.layer_norm(t.view(B, D * H * W, C), _cast(self.norm.weight, cd), _cast(self.norm.bias, cd),
                              out_dtype)
            return y.view(B, D, H, W, C).permute(0, 4, 1, 2, 3)

    # ---- weight init / loading: host-side state_dict surgery (video_swin.py:484-570) -------------
    def inflate_weights(self):
        """2-D Swin checkpoint -> 3-D (video_swin.py:484-535): conv weight repeated over time / pd, bias
        tables bicubic-resized to (2wh-1, 2ww-1) when needed then tiled 2wd-1 times."""
        ckpt = torch.load(self.pretrained, map_location="cpu")
        sd = ckpt["model"]
        for k in [k for k in sd if "relative_position_index" in k or "attn_mask" in k]:
            del sd[k]
        pd = self.patch_size[0]
        sd["patch_embed.proj.weight"] = sd["patch_embed.proj.weight"].unsqueeze(2).repeat(1, 1, pd, 1, 1) / pd
        wd, wh, ww = self.window_size
        own = self.state_dict()
        for k in [k for k in sd if "relative_position_bias_table" in k]:
            tab = sd[k]
            L1, nH1 = tab.size()
            nH2 = own[k].size(1)
            L2 = (2 * wh - 1) * (2 * ww - 1)
            if nH1 != nH2:
                print(f"Error in loading {k}, passing")
            elif L1 != L2:
                S1 = int(L1 ** 0.5)
                tab = torch.nn.functional.interpolate(tab.permute(1, 0).view(1, nH1, S1, S1),
                                                      size=(2 * wh - 1, 2 * ww - 1), mode="bicubic")
                tab = tab.view(nH2, L2).permute(1, 0)
            sd[k] = tab.repeat(2 * wd - 1, 1)
        msg = self.load_state_dict(sd, strict=False)
        print(msg)
        print(f"=> loaded successfully '{self.pretrained}'")
        del ckpt

    def init_weights(self, pretrained=None):
        """video_swin.py:537-570: Linear trunc-normal(.02)/bias 0, LayerNorm 1/0, Conv3d default."""
        def _init(m):
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str):
            self.apply(_init)
            print(f"load model from: {self.pretrained}")
            if self.pretrained2d:
                print("Inflate 2D model into 3D model.")
                self.inflate_weights()
            else:
                raise NotImplementedError("Directly load 3D model, not supported")
        elif self.pretrained is None:
            self.apply(_init)
        else:
            raise TypeError("pretrained must be a str or None")


# ----------------------------------------------------------------------------------------------
# factory (video_swin.py:573-659)
# ----------------------------------------------------------------------------------------------
def load_checkpoint_3d(model_path):
    """strip the ``backbone.`` prefix of a Kinetics Video-Swin checkpoint (video_swin.py:653-659)"""
    sd = torch.load(model_path, map_location="cpu")["state_dict"]
    return {k.replace("backbone.", ""): v for k, v in sd.items()}


def get_vidswin_model(args):
    """Same selection logic and side effects on ``args`` as video_swin.py:573-650; the backbone config
    is read with ``config_loader`` (``_base_`` inheritance) from the reference's ``visbackbone/`` files
    when they exist, else from the built-in presets of the same names."""
    from .config_loader import load_backbone_cfg
    size = args.vis_backbone_size
    if int(args.size_img) == 384 and size == "large":
        config_path = "./visbackbone/swin_%s_384_patch244_window81212_kinetics600_22k.py" % size
        model_path = ("./models/swin_transformer/swin_%s_patch4_window12_384_22k.pth" % size
                      if args.vis_backbone_init == "2d" else
                      "./models/video_swin_transformer/swin_%s_384_patch244_window81212_kinetics%s_22k.pth"
                      % (size, args.kinetics))
    elif size not in ["tiny", "violet"]:
        config_path = "./visbackbone/swin_%s_patch244_window877_kinetics400_22k.py" % size
        model_path = ("./models/swin_transformer/swin_%s_patch4_window7_224_22k.pth" % size
                      if args.vis_backbone_init == "2d" else
                      "./models/video_swin_transformer/swin_%s_patch244_window877_kinetics%s_22k.pth"
                      % (size, args.kinetics))
    elif size == "tiny":
        assert int(args.size_img) == 224
        config_path = "./visbackbone/swin_%s_patch244_window877_kinetics400_1k.py" % size
        model_path = ("./models/swin_transformer/swin_%s_patch4_window7_224.pth" % size
                      if args.vis_backbone_init == "2d" else
                      "./models/video_swin_transformer/swin_tiny_patch244_window877_kinetics400_1k.pth")
    else:
        assert size == "violet"
        config_path = "videoswin/swin_violet_patch244_window877.py"
        model_path = None
        args.vis_backbone_init = "random"
    print(f"video swin (config path): {config_path}")
    bb = load_backbone_cfg(config_path)
    if args.vis_backbone_init == "2d":
        print(f"video swin with pre-trained 2d (model path): {model_path}")
        pretrained2d = model_path
        args.vis_backbone_pretrained_weight = model_path
    elif args.vis_backbone_init == "random":
        print("video swin random initialized")
        pretrained2d = None
        model_path = None
        args.vis_backbone_pretrained_weight = None
    else:
        print(f"video swin with pre-trained 3d (model path): {model_path}")
        pretrained2d = None
        args.vis_backbone_pretrained_weight = model_path
    video_swin = SwinTransformer3D(
        pretrained=pretrained2d, pretrained2d=True, patch_size=bb["patch_size"], in_chans=3,
        embed_dim=bb["embed_dim"], depths=bb["depths"], num_heads=bb["num_heads"], window_size=bb["window_size"],
        mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2,
        norm_layer=torch.nn.LayerNorm, patch_norm=bb["patch_norm"], frozen_stages=-1, use_checkpoint=False)
    if args.vis_backbone_init == "3d" and model_path is not None:
        sd = load_checkpoint_3d(model_path)
        missing, unexpected = video_swin.load_state_dict(sd, strict=False)
        print(f"Missing keys in loaded video_swin_transformerr: {missing}")
        print(f"Unexpected keys in loaded video_swin_transformer: {unexpected}")
    else:
        video_swin.init_weights()
    return video_swin
