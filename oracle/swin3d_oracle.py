"""CPU oracle for the Video-Swin 3D hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, in plain fp32/fp64 PyTorch on the CPU, the algorithm of the reference's
``visbackbone/video_swin.py`` (tsujuifu/pytorch_empirical-mvm).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker (or as the timed CPU baseline) -- never on the product path.

It is deliberately *not* a copy of the reference: it is a stateless functional implementation over a
``state_dict`` that replaces the reference's ``roll``/``view``/``permute`` plumbing with explicit
closed-form index maps, so that the same closed forms can be checked
bit-exactly against the CUDA index kernels.

Parity status: PINNED.  The reference has no tests of its own, so the oracle
is pinned against outputs of the reference module itself, imported from /root/reference in the build
container by ``tests/golden/make_golden.py``; the resulting vectors live in ``tests/golden/`` and
``tests/test_oracle_golden.py`` replays them (index maps bit-exact, activations/gradients to fp32
round-off).

Reference citations (file:line relative to /root/reference):
  get_window_size        visbackbone/video_swin.py:95-108
  window gather/scatter  visbackbone/video_swin.py:84-93, 220-241
  compute_mask           visbackbone/video_swin.py:292-307
  relative_position_index visbackbone/video_swin.py:123-137
  WindowAttention3D.forward visbackbone/video_swin.py:147-172
  Mlp                    visbackbone/video_swin.py:65-81
  SwinTransformerBlock3D visbackbone/video_swin.py:206-263
  PatchMerging           visbackbone/video_swin.py:273-289
  PatchEmbed3D           visbackbone/video_swin.py:390-407
  BasicLayer / SwinTransformer3D.forward visbackbone/video_swin.py:352-370, 470-482
  drop_path              visbackbone/video_swin.py:46-54
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Triple = Tuple[int, int, int]


# --------------------------------------------------------------------------------------
# integer geometry (bit-exact domain)
# --------------------------------------------------------------------------------------
def effective_window(grid: Triple, window: Triple, shift: Optional[Triple] = None):
    """video_swin.py:95-108 -- clamp the window to the grid (``<=``) and zero the shift there."""
    ws = [grid[a] if grid[a] <= window[a] else window[a] for a in range(3)]
    if shift is None:
        return tuple(ws)
    ss = [0 if grid[a] <= window[a] else shift[a] for a in range(3)]
    return tuple(ws), tuple(ss)


def padded_grid(grid: Triple, ws: Triple) -> Triple:
    """video_swin.py:213-218 / 357-359 -- ceil each axis to a multiple of the effective window."""
    return tuple(-(-grid[a] // ws[a]) * ws[a] for a in range(3))


def window_gather_map(pgrid: Triple, ws: Triple, ss: Triple) -> np.ndarray:
    """(nW, N) int64: flat index (row-major over the PADDED grid) of the token that lands in
    window ``w`` slot ``n`` after ``roll(-ss)`` + ``window_partition`` (video_swin.py:84-88, 220-229).

    Closed form: window (a,b,c), slot (i,j,k) reads
    ``x[(a*wd+i+sd) % Dp, (b*wh+j+sh) % Hp, (c*ww+k+sw) % Wp]``."""
    Dp, Hp, Wp = pgrid
    wd, wh, ww = ws
    sd, sh, sw = ss
    a = np.arange(Dp // wd)[:, None, None, None, None, None]
    b = np.arange(Hp // wh)[None, :, None, None, None, None]
    c = np.arange(Wp // ww)[None, None, :, None, None, None]
    i = np.arange(wd)[None, None, None, :, None, None]
    j = np.arange(wh)[None, None, None, None, :, None]
    k = np.arange(ww)[None, None, None, None, None, :]
    d = (a * wd + i + sd) % Dp
    h = (b * wh + j + sh) % Hp
    w = (c * ww + k + sw) % Wp
    flat = (d * Hp + h) * Wp + w
    return flat.reshape(-1, wd * wh * ww).astype(np.int64)


def _axis_region(S: int, w: int, s: int) -> np.ndarray:
    """Region id along one axis of the shifted frame (video_swin.py:296-298).

    The three slices are ``[:-w]``, ``[-w:-s]``, ``[-s:]`` assigned in that order; with ``s == 0``
    the last slice is ``[0:]`` and overwrites the whole axis with id 2."""
    p = np.arange(S)
    if s == 0:
        return np.full(S, 2, dtype=np.int64)
    rid = np.zeros(S, dtype=np.int64)
    rid[(p >= S - w) & (p < S - s)] = 1
    rid[p >= S - s] = 2
    return rid


def window_region_ids(pgrid: Triple, ws: Triple, ss: Triple) -> np.ndarray:
    """(nW, N) int64 region counter ``cnt = 9*rd + 3*rh + rw`` of every window slot
    (video_swin.py:294-302).  The counter lives in the SHIFTED frame, i.e. it is partitioned
    without a roll."""
    Dp, Hp, Wp = pgrid
    rd = _axis_region(Dp, ws[0], ss[0])[:, None, None]
    rh = _axis_region(Hp, ws[1], ss[1])[None, :, None]
    rw = _axis_region(Wp, ws[2], ss[2])[None, None, :]
    cnt = 9 * rd + 3 * rh + rw  # (Dp,Hp,Wp)
    plain = window_gather_map(pgrid, ws, (0, 0, 0))
    return cnt.reshape(-1)[plain]


def shift_mask(pgrid: Triple, ws: Triple, ss: Triple) -> torch.Tensor:
    """(nW, N, N) fp32 of {0, -100} (video_swin.py:303-307)."""
    rid = torch.from_numpy(window_region_ids(pgrid, ws, ss))
    diff = rid[:, None, :] != rid[:, :, None]
    return torch.where(diff, torch.tensor(-100.0), torch.tensor(0.0))


def relative_position_index(window: Triple) -> np.ndarray:
    """(N, N) int64 (video_swin.py:123-137): row-major slot (d,h,w);
    ``idx[i,j] = (di-dj+wd-1)(2wh-1)(2ww-1) + (hi-hj+wh-1)(2ww-1) + (wi-wj+ww-1)``."""
    wd, wh, ww = window
    n = np.arange(wd * wh * ww)
    d, h, w = n // (wh * ww), (n // ww) % wh, n % ww
    rel = ((d[:, None] - d[None, :] + wd - 1) * (2 * wh - 1) * (2 * ww - 1)
           + (h[:, None] - h[None, :] + wh - 1) * (2 * ww - 1)
           + (w[:, None] - w[None, :] + ww - 1))
    return rel.astype(np.int64)


def merge_gather_map(grid: Triple) -> np.ndarray:
    """(D*H2*W2, 4) int64 source token (flat over the UNPADDED (D,H,W) grid, -1 = zero pad) of
    the four concatenated channel groups of PatchMerging (video_swin.py:276-284): order
    (dh,dw) = (0,0), (1,0), (0,1), (1,1)."""
    D, H, W = grid
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    d = np.arange(D)[:, None, None]
    h2 = np.arange(H2)[None, :, None]
    w2 = np.arange(W2)[None, None, :]
    out = np.empty((D, H2, W2, 4), dtype=np.int64)
    for g, (dh, dw) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        h, w = 2 * h2 + dh, 2 * w2 + dw
        flat = (d * H + h) * W + w
        valid = (h < H) & (w < W)
        out[..., g] = np.where(valid & np.ones_like(d, bool), flat, -1)
    return out.reshape(-1, 4)


# --------------------------------------------------------------------------------------
# floating-point path
# --------------------------------------------------------------------------------------
@dataclass
class SwinCfg:
    patch_size: Triple = (2, 4, 4)
    in_chans: int = 3
    embed_dim: int = 128
    depths: Sequence[int] = (2, 2, 18, 2)
    num_heads: Sequence[int] = (4, 8, 16, 32)
    window_size: Triple = (8, 7, 7)
    mlp_ratio: float = 4.0
    qk_scale: Optional[float] = None
    patch_norm: bool = True
    drop_path_rate: float = 0.0  # oracle runs eval-mode unless drop-path keep masks are given

    @staticmethod
    def swin_b(**kw):
        return SwinCfg(embed_dim=128, num_heads=(4, 8, 16, 32), **kw)

    @staticmethod
    def violet(**kw):  # visbackbone/swin_violet.py:6-10
        return SwinCfg(embed_dim=96, num_heads=(3, 6, 12, 24), **kw)

    @staticmethod
    def swin_l_384(**kw):  # visbackbone/swin_large.py + ..._window81212_...py
        return SwinCfg(embed_dim=192, num_heads=(6, 12, 24, 48), window_size=(8, 12, 12), **kw)


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def patch_embed(x, sd, prefix, cfg: SwinCfg):
    """video_swin.py:390-407.  x (B,3,D,H,W) -> tokens (B,D,H',W',E) channels-last."""
    pd, ph, pw = cfg.patch_size
    B, Cin, D, H, W = x.shape
    x = F.pad(x, (0, (-W) % pw, 0, (-H) % ph, 0, 1))  # right/bottom pad to x4, +1 zero frame
    Hp, Wp = x.shape[3] // ph, x.shape[4] // pw
    Dout = x.shape[2] - pd + 1
    wt = sd[prefix + "proj.weight"]  # (E,Cin,pd,ph,pw)
    E = wt.shape[0]
    # explicit im2col: temporal stride 1, spatial stride = patch
    cols = []
    for dt in range(pd):
        xs = x[:, :, dt:dt + Dout]  # (B,Cin,Dout,H,W)
        xs = xs.reshape(B, Cin, Dout, Hp, ph, Wp, pw).permute(0, 2, 3, 5, 1, 4, 6)  # B,D,Hp,Wp,Cin,ph,pw
        cols.append(xs)
    col = torch.stack(cols, dim=5)  # B,D,Hp,Wp,Cin,pd,ph,pw
    col = col.reshape(B, Dout, Hp, Wp, Cin * pd * ph * pw)
    y = col @ wt.reshape(E, -1).t() + sd[prefix + "proj.bias"]
    if cfg.patch_norm:
        y = _ln(y, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return y


def window_attention(xw, sd, prefix, num_heads, mask, scale, probs_out=None):
    """video_swin.py:147-172.  xw (B_, N, C); mask (nW,N,N) or None."""
    B_, N, C = xw.shape
    hd = C // num_heads
    qkv = xw @ sd[prefix + "qkv.weight"].t()
    if (prefix + "qkv.bias") in sd:
        qkv = qkv + sd[prefix + "qkv.bias"]
    qkv = qkv.reshape(B_, N, 3, num_heads, hd)
    q = qkv[:, :, 0].transpose(1, 2) * scale
    k = qkv[:, :, 1].transpose(1, 2)
    v = qkv[:, :, 2].transpose(1, 2)
    s = q @ k.transpose(-1, -2)  # (B_,nH,N,N)
    idx = sd[prefix + "relative_position_index"][:N, :N]  # slice quirk, video_swin.py:155
    bias = sd[prefix + "relative_position_bias_table"][idx.reshape(-1)].reshape(N, N, num_heads)
    s = s + bias.permute(2, 0, 1)[None]
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]).reshape(B_, num_heads, N, N)
    p = torch.softmax(s, dim=-1)
    if probs_out is not None:
        probs_out.append(p)
    o = (p @ v).transpose(1, 2).reshape(B_, N, C)
    return o @ sd[prefix + "proj.weight"].t() + sd[prefix + "proj.bias"]


def mlp(x, sd, prefix):
    """video_swin.py:75-81 (erf GELU, dropout p=0)."""
    h = F.gelu(x @ sd[prefix + "fc1.weight"].t() + sd[prefix + "fc1.bias"])
    return h @ sd[prefix + "fc2.weight"].t() + sd[prefix + "fc2.bias"]


def swin_block(x, sd, prefix, num_heads, window, shift, scale, keep1=None, keep2=None):
    """video_swin.py:206-263.  x (B,D,H,W,C).  ``keepX``: optional (B,) drop-path factors
    (``floor(keep+U)/keep``) multiplying each branch (video_swin.py:46-54)."""
    B, D, H, W, C = x.shape
    ws, ss = effective_window((D, H, W), window, shift)
    pg = padded_grid((D, H, W), ws)
    n1 = _ln(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    n1 = F.pad(n1, (0, 0, 0, pg[2] - W, 0, pg[1] - H, 0, pg[0] - D))  # zeros AFTER the norm
    gmap = torch.from_numpy(window_gather_map(pg, ws, ss))  # (nW,N)
    nW, N = gmap.shape
    flat = n1.reshape(B, pg[0] * pg[1] * pg[2], C)
    xw = flat[:, gmap.reshape(-1)].reshape(B * nW, N, C)
    mask = shift_mask(pg, ws, ss).to(x.dtype) if any(s > 0 for s in ss) else None
    aw = window_attention(xw, sd, prefix + "attn.", num_heads, mask, scale)
    back = torch.empty_like(flat)
    back[:, gmap.reshape(-1)] = aw.reshape(B, nW * N, C)  # gmap is a permutation
    a = back.reshape(B, *pg, C)[:, :D, :H, :W]
    if keep1 is not None:
        a = a * keep1.reshape(B, 1, 1, 1, 1)
    x = x + a
    m = mlp(_ln(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"]), sd, prefix + "mlp.")
    if keep2 is not None:
        m = m * keep2.reshape(B, 1, 1, 1, 1)
    return x + m


def patch_merge(x, sd, prefix):
    """video_swin.py:273-289."""
    B, D, H, W, C = x.shape
    gm = torch.from_numpy(merge_gather_map((D, H, W)))  # (T2,4)
    flat = torch.cat([x.reshape(B, D * H * W, C), x.new_zeros(B, 1, C)], dim=1)  # slot -1 -> zeros
    cat = flat[:, gm.reshape(-1)].reshape(B, D, (H + 1) // 2, (W + 1) // 2, 4 * C)
    cat = _ln(cat, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return cat @ sd[prefix + "reduction.weight"].t()


def drop_path_rates(cfg: SwinCfg) -> List[float]:
    """video_swin.py:447."""
    return [v.item() for v in torch.linspace(0, cfg.drop_path_rate, sum(cfg.depths))]


def swin_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: SwinCfg,
                 keeps: Optional[List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]] = None,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """SwinTransformer3D.forward (video_swin.py:470-482).  Returns (B, 8E, D, H/32, W/32) as the same
    permuted view of a channels-last buffer that the reference returns."""
    t = patch_embed(x, sd, "patch_embed.", cfg)
    if taps is not None:
        taps["patch_embed"] = t
    shift = tuple(w // 2 for w in cfg.window_size)
    blk_id = 0
    for li, depth in enumerate(cfg.depths):
        C = t.shape[-1]
        nH = cfg.num_heads[li]
        scale = cfg.qk_scale or (C // nH) ** -0.5
        for bi in range(depth):
            k1, k2 = keeps[blk_id] if keeps is not None else (None, None)
            t = swin_block(t, sd, f"layers.{li}.blocks.{bi}.", nH, tuple(cfg.window_size),
                           (0, 0, 0) if bi % 2 == 0 else shift, scale, k1, k2)
            blk_id += 1
        if taps is not None:
            taps[f"stage{li}"] = t
        if li < len(cfg.depths) - 1:
            t = patch_merge(t, sd, f"layers.{li}.downsample.")
    t = _ln(t, sd["norm.weight"], sd["norm.bias"])
    return t.permute(0, 4, 1, 2, 3)


# --------------------------------------------------------------------------------------
# state_dict construction (shapes/keys of SURVEY section 8b) -- random init like init_weights()
# --------------------------------------------------------------------------------------
def _trunc_normal(shape, std, gen):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0, generator=gen)
    return t


def make_state_dict(cfg: SwinCfg, seed: int = 0, ln_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random state_dict with the reference's 351-key layout (for depths [2,2,18,2]).
    ``ln_jitter`` perturbs LayerNorm weights/biases and Linear biases away from the 1/0 init so
    tests see them (the reference initialises them to exactly 1/0, video_swin.py:544-551)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    E = cfg.embed_dim
    kvol = cfg.in_chans * cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
    bound = 1.0 / math.sqrt(kvol)
    sd["patch_embed.proj.weight"] = (torch.rand((E, cfg.in_chans, *cfg.patch_size), generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand(E, generator=g) * 2 - 1) * bound

    def ln(prefix, n):
        sd[prefix + "weight"] = torch.ones(n) + ln_jitter * torch.randn(n, generator=g)
        sd[prefix + "bias"] = ln_jitter * torch.randn(n, generator=g)

    def lin(prefix, o, i, bias=True):
        sd[prefix + "weight"] = _trunc_normal((o, i), 0.02, g)
        if bias:
            sd[prefix + "bias"] = ln_jitter * torch.randn(o, generator=g)

    if cfg.patch_norm:
        ln("patch_embed.norm.", E)
    wd, wh, ww = cfg.window_size
    rpi = torch.from_numpy(relative_position_index(tuple(cfg.window_size)))
    for li, depth in enumerate(cfg.depths):
        C = E * 2 ** li
        nH = cfg.num_heads[li]
        for bi in range(depth):
            p = f"layers.{li}.blocks.{bi}."
            ln(p + "norm1.", C)
            sd[p + "attn.relative_position_bias_table"] = _trunc_normal(
                ((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), nH), 0.02, g)
            sd[p + "attn.relative_position_index"] = rpi.clone()
            lin(p + "attn.qkv.", 3 * C, C)
            lin(p + "attn.proj.", C, C)
            ln(p + "norm2.", C)
            lin(p + "mlp.fc1.", int(C * cfg.mlp_ratio), C)
            lin(p + "mlp.fc2.", C, int(C * cfg.mlp_ratio))
        if li < len(cfg.depths) - 1:
            p = f"layers.{li}.downsample."
            lin(p + "reduction.", 2 * C, 4 * C, bias=False)
            ln(p + "norm.", 4 * C)
    ln("norm.", E * 2 ** (len(cfg.depths) - 1))
    return sd


def forward_backward(sd, x, cfg: SwinCfg, R: torch.Tensor, keeps=None, dtype=torch.float32):
    """Loss = sum(y * R) (SURVEY 8c pitfall 1: never mean(y^2)).  Returns (y, grads dict)."""
    params = {k: (v.detach().to(dtype).requires_grad_(True) if v.is_floating_point() else v)
              for k, v in sd.items()}
    y = swin_forward(params, x.to(dtype), cfg, keeps)
    loss = (y * R.to(dtype)).sum()
    names = [k for k, v in params.items() if v.is_floating_point()]
    grads = torch.autograd.grad(loss, [params[k] for k in names])
    return y.detach(), dict(zip(names, grads))
