"""Autograd functions over the vsw C ABI -- the host side of the Video-Swin 3D hot path.

Every tensor op on the path is a hand-written CUDA kernel reached through ``_lib`` (ctypes);
PyTorch only owns memory, streams and the autograd graph.  Tensors are channels-last tokens
``(B, T, C)``; windowed tensors are ``(B*nW*N, C)``.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib as L

Triple = Tuple[int, int, int]
LN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# geometry (host ints) + cached device index maps
# ----------------------------------------------------------------------------------------------
def get_window_size(x_size, window_size, shift_size=None):
    """Clamp the window to the grid and zero the shift on clamped axes (video_swin.py:95-108)."""
    ws = tuple(int(x_size[i]) if x_size[i] <= window_size[i] else int(window_size[i]) for i in range(len(x_size)))
    if shift_size is None:
        return ws
    ss = tuple(0 if x_size[i] <= window_size[i] else int(shift_size[i]) for i in range(len(x_size)))
    return ws, ss


@dataclass(frozen=True)
class WindowPlan:
    grid: Triple          # (D,H,W) unpadded
    pgrid: Triple         # padded to window multiples
    ws: Triple            # effective window
    ss: Triple            # effective shift
    nW: int
    N: int
    gather: torch.Tensor  # (nW*N,) int32, token index or -1
    region: Optional[torch.Tensor]  # (nW*N,) uint8 region ids; None when the block is unshifted

    @property
    def shifted(self) -> bool:
        return any(s > 0 for s in self.ss)


@functools.lru_cache(maxsize=256)
def _window_plan(grid: Triple, window: Triple, shift: Triple, device_index: int) -> WindowPlan:
    ws, ss = get_window_size(grid, window, shift)
    pg = tuple(-(-grid[a] // ws[a]) * ws[a] for a in range(3))
    nW = (pg[0] // ws[0]) * (pg[1] // ws[1]) * (pg[2] // ws[2])
    N = ws[0] * ws[1] * ws[2]
    dev = torch.device("cuda", device_index)
    with torch.cuda.device(dev):
        gather = torch.empty(nW * N, dtype=torch.int32, device=dev)
        region = torch.empty(nW * N, dtype=torch.uint8, device=dev) if any(ss) else None
        L.check(L.lib().vsw_window_maps(*grid, *ws, *ss, L.ptr(gather), L.ptr(region), L.stream()), "vsw_window_maps")
    return WindowPlan(grid, pg, ws, ss, nW, N, gather, region)


def window_plan(grid, window, shift, device) -> WindowPlan:
    device = torch.device(device)
    if device.type != "cuda":
        raise L.VswError("vsw kernels are CUDA-only (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return _window_plan(tuple(int(g) for g in grid), tuple(int(w) for w in window), tuple(int(s) for s in shift), idx)


@functools.lru_cache(maxsize=64)
def _merge_map(grid: Triple, device_index: int) -> torch.Tensor:
    D, H, W = grid
    dev = torch.device("cuda", device_index)
    with torch.cuda.device(dev):
        out = torch.empty(D * ((H + 1) // 2) * ((W + 1) // 2) * 4, dtype=torch.int32, device=dev)
        L.check(L.lib().vsw_merge_map(D, H, W, L.ptr(out), L.stream()), "vsw_merge_map")
    return out


def merge_map(grid, device) -> torch.Tensor:
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return _merge_map(tuple(int(g) for g in grid), idx)


def rel_pos_index(window: Triple, device) -> torch.Tensor:
    """relative_position_index buffer (N,N) int64 (video_swin.py:123-137), built on the device."""
    N = window[0] * window[1] * window[2]
    out = torch.empty(N, N, dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        L.check(L.lib().vsw_rel_pos_index(*window, L.ptr(out), L.stream()), "vsw_rel_pos_index")
    return out


def shift_mask_from_region(region: torch.Tensor, nW: int, N: int, dtype=torch.float32) -> torch.Tensor:
    out = torch.empty(nW, N, N, dtype=dtype, device=region.device)
    with torch.cuda.device(out.device):
        L.check(L.lib().vsw_shift_mask(L.ptr(region), nW, N, L.ptr(out), L.dt(dtype), L.stream()), "vsw_shift_mask")
    return out


def bias_codes(rel_index: torch.Tensor, N: int):
    """rowcode[i] = index[i,0], colcode[j] = index[0,j] - index[0,0] over the [:N,:N] slice
    (video_swin.py:155); index[i,j] == rowcode[i] + colcode[j] because the index is a function of
    coordinate differences only."""
    idx = rel_index[:N, :N]
    rowcode = idx[:, 0].to(torch.int32).contiguous()
    colcode = (idx[0, :] - idx[0, 0]).to(torch.int32).contiguous()
    return rowcode, colcode


# ----------------------------------------------------------------------------------------------
# optional per-kernel CUDA-event timing (bench.py's roofline leg); off by default
# ----------------------------------------------------------------------------------------------
class KernelTimer:
    """Records a CUDA-event pair on the launching stream around every call of the wrapped kernel
    families, with the ALGORITHMIC flops / bytes of that call."""

    def __init__(self):
        self.records = {}  # name -> [(start, end, flops, bytes)]

    def begin(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        return ev

    def end(self, name, start, flops, nbytes):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        self.records.setdefault(name, []).append((start, ev, flops, nbytes))

    def summary(self):
        out = {}
        for name, recs in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
            out[name] = dict(launches=len(recs), ms=ms, flops=sum(r[2] for r in recs), bytes=sum(r[3] for r in recs))
        return out


PROFILER: Optional[KernelTimer] = None


def _esz(t):
    return t.element_size()


# ----------------------------------------------------------------------------------------------
# thin kernel wrappers (no autograd)
# ----------------------------------------------------------------------------------------------
def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


def ln_fwd(x, gamma, beta, gmap, B, Tin, Tout, C, out_dtype=None, want_stats=True):
    out_dtype = out_dtype or x.dtype
    y = _empty((B, Tout, C), out_dtype, x.device)
    mean = _empty((B * Tout,), torch.float32, x.device) if want_stats else None
    rstd = _empty((B * Tout,), torch.float32, x.device) if want_stats else None
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_ln_fwd(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(gmap), L.ptr(y), L.ptr(mean), L.ptr(rstd),
                               B, Tin, Tout, C, LN_EPS, L.dt(x), L.dt(out_dtype), L.stream()), "vsw_ln_fwd")
    if t0 is not None:
        PROFILER.end("ln_fwd", t0, 0.0, B * C * (Tin * _esz(x) + Tout * y.element_size()))
    return y, mean, rstd


def ln_bwd(dy, x, gamma, mean, rstd, gmap, dres, B, Tin, Tout, C, need_dx=True, need_dparams=True, grad_dtype=None,
           g_key=0, b_key=0):
    """grad_dtype: dtype dgamma / dbeta are written in (default fp32); g_key / b_key: gradient-sink keys of the two parameters
    (with a sink installed the kernel writes straight into their slots of the flat gradient buffer)"""
    dx = _empty((B, Tin, C), x.dtype, x.device) if need_dx else None
    gdt = grad_dtype or torch.float32
    dg = db = None
    if need_dparams:
        dg = _sink_view(g_key, (C,), gdt)
        db = _sink_view(b_key, (C,), gdt)
        dg = dg if dg is not None else _empty((C,), gdt, x.device)
        db = db if db is not None else _empty((C,), gdt, x.device)
    wsb = int(L.lib().vsw_ln_bwd_workspace(C)) if need_dparams else 0
    ws = _empty((max(wsb, 4),), torch.uint8, x.device) if need_dparams else None
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_ln_bwd_ex(L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(gmap), L.ptr(dres),
                                  L.ptr(dx), L.ptr(dg), L.ptr(db), B, Tin, Tout, C, L.dt(x), L.dt(dy), L.dt(gdt), L.ptr(ws), wsb,
                                  L.stream()), "vsw_ln_bwd")
    if t0 is not None:
        PROFILER.end("ln_bwd", t0, 0.0, B * C * (Tout * dy.element_size() + Tin * _esz(x) * (2 + (dres is not None))))
    return dx, dg, db


def residual_add(x, y, rowmap, rowscale, B, R, T, C):
    """out[b, map[r]] = x[b, map[r]] + rowscale[b] * y[b, r]  with x / out in x.dtype (the wider residual stream) and the
    branch output y in its own 16-bit dtype"""
    out = _empty((B, T, C), x.dtype, x.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_residual_add(L.ptr(x), L.ptr(y), L.ptr(rowmap), L.ptr(rowscale), L.ptr(out), B, R, T, C, L.dt(x),
                                     L.dt(y), L.stream()), "vsw_residual_add")
    if t0 is not None:
        PROFILER.end("residual_add", t0, 0.0, B * C * (2 * T * _esz(x) + R * _esz(y)))
    return out


def linear_fwd(x2d, w, bias, M, N, K, epi=L.EPI_BIAS, out=None, aux_out=None, res=None, rowmap=None, rowscale=None,
               rows_per_batch=0, dst_rows_per_batch=0, out_rows=None):
    y = out if out is not None else _empty((out_rows if out_rows is not None else M, N), x2d.dtype, x2d.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_linear_fwd(L.ptr(x2d), L.ptr(w), L.ptr(bias), L.ptr(y), M, N, K, epi, L.ptr(aux_out), L.ptr(res),
                                   L.ptr(rowmap), L.ptr(rowscale), rows_per_batch, dst_rows_per_batch, L.dt(x2d),
                                   L.stream()), "vsw_linear_fwd")
    if t0 is not None:
        e = _esz(x2d)
        nb = (M * K + N * K + M * N * (1 + (aux_out is not None) + (res is not None))) * e
        PROFILER.end("linear_fwd", t0, 2.0 * M * N * K, nb)
    return y


def linear_dgrad(dy, w, M, N, K, a_rowmap=None, a_rowscale=None, rows_per_batch=0, src_rows_per_batch=0, a_out=None,
                 gelu_pre=None, mul=None):
    """dx = (A w) [* gelu'(gelu_pre) | * mul];  `mul` is the gelu' tensor written by the EPI_GELU_GRAD forward epilogue."""
    dx = _empty((M, K), dy.dtype, dy.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    if mul is not None:
        L.check(L.lib().vsw_linear_dgrad_mul(L.ptr(dy), L.ptr(w), L.ptr(dx), M, N, K, L.ptr(a_rowmap), L.ptr(a_rowscale),
                                             rows_per_batch, src_rows_per_batch, L.ptr(a_out), L.ptr(mul), L.dt(dy),
                                             L.stream()), "vsw_linear_dgrad_mul")
    else:
        L.check(L.lib().vsw_linear_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), M, N, K, L.ptr(a_rowmap), L.ptr(a_rowscale),
                                         rows_per_batch, src_rows_per_batch, L.ptr(a_out), L.ptr(gelu_pre), L.dt(dy),
                                         L.stream()), "vsw_linear_dgrad")
    if t0 is not None:
        e = _esz(dy)
        nb = (M * N * (1 + (a_out is not None)) + N * K + M * K * (1 + (gelu_pre is not None or mul is not None))) * e
        PROFILER.end("linear_dgrad", t0, 2.0 * M * N * K, nb)
    return dx


# Gradient sink (dp.FlatGradReducer): when set, maps the data_ptr of a parameter to a fresh view of its slot in one flat
# gradient buffer, so the wgrad kernels write where the all-reduce reads (no per-parameter copy).  Returns None for unknown
# parameters, wrong shape / dtype, or a slot already written in this step (autograd then accumulates the usual way).
GRAD_SINK = None


def _sink_view(key, shape, dtype):
    if GRAD_SINK is None or not key:
        return None
    return GRAD_SINK(key, tuple(shape), dtype)


def _key(t):
    return t.data_ptr() if t is not None else 0


def linear_wgrad(dy, x2d, M, N, K, need_bias=True, grad_dtype=None, w_key=0, b_key=0):
    grad_dtype = grad_dtype or dy.dtype
    dw = _sink_view(w_key, (N, K), grad_dtype)
    if dw is None:
        dw = _empty((N, K), grad_dtype, dy.device)
    db = None
    if need_bias:
        db = _sink_view(b_key, (N,), grad_dtype)
        if db is None:
            db = _empty((N,), grad_dtype, dy.device)
    wsb = int(L.lib().vsw_linear_wgrad_workspace(M, N, K))
    ws = _empty((max(wsb, 4),), torch.uint8, dy.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_linear_wgrad(L.ptr(dy), L.ptr(x2d), L.ptr(dw), L.ptr(db), M, N, K, L.dt(dy), L.dt(grad_dtype),
                                     L.ptr(ws), wsb, L.stream()), "vsw_linear_wgrad")
    if t0 is not None:
        PROFILER.end("linear_wgrad", t0, 2.0 * M * N * K, (M * N + M * K + N * K) * _esz(dy))
    return dw, db


def window_dims(planes=0, window=None):
    """packed layout hint of the attention ABI: effective window depth | configured window height << 8 | width << 16"""
    wh, ww = (int(window[1]), int(window[2])) if window is not None else (0, 0)
    return (int(planes) & 0xFF) | ((wh & 0xFF) << 8) | ((ww & 0xFF) << 16)


def attn_fwd(qkv, table, rowcode, colcode, region, dense_mask, B_, nW, N, nH, hd, scale, window=None):
    C = nH * hd
    out = _empty((B_ * N, C), qkv.dtype, qkv.device)
    lse = _empty((B_, nH, N), torch.float32, qkv.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_window_attn_fwd(L.ptr(qkv), L.ptr(table), L.ptr(rowcode), L.ptr(colcode), L.ptr(region),
                                        L.ptr(dense_mask), L.ptr(out), L.ptr(lse), B_, nW, N, nH, hd, table.shape[0],
                                        float(scale), window_dims(0, window), L.dt(qkv), L.stream()), "vsw_window_attn_fwd")
    if t0 is not None:  # QK^T + PV = 4*N*N*hd flops per (window, head); q,k,v read + o written once
        PROFILER.end("window_attn_fwd", t0, 4.0 * B_ * nH * N * N * hd, 4 * B_ * N * C * _esz(qkv))
    return out, lse


def attn_bwd(qkv, out, dout, lse, table, rowcode, colcode, region, dense_mask, B_, nW, N, nH, hd, scale, planes=0,
             window=None):
    dqkv = torch.empty_like(qkv)
    Lt = table.shape[0]
    dtable = _empty((Lt, nH), torch.float32, qkv.device)
    wsb = int(L.lib().vsw_window_attn_bwd_workspace(B_, N, nH, hd, Lt))
    ws = _empty((max(wsb, 4),), torch.uint8, qkv.device)
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_window_attn_bwd(L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(table), L.ptr(rowcode),
                                        L.ptr(colcode), L.ptr(region), L.ptr(dense_mask), L.ptr(dqkv), L.ptr(dtable),
                                        B_, nW, N, nH, hd, Lt, float(scale), window_dims(planes, window), L.dt(qkv), L.ptr(ws), wsb,
                                        L.stream()),
            "vsw_window_attn_bwd")
    if t0 is not None:  # dV, dP, dQ, dK = 8*N*N*hd useful flops (the S/P recompute is not counted)
        PROFILER.end("window_attn_bwd", t0, 8.0 * B_ * nH * N * N * hd, 8 * B_ * N * nH * hd * _esz(qkv))
    return dqkv, dtable


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def _grad_to(g, like, key=0):
    """fp32 kernel output -> gradient in the parameter's dtype (into its flat-buffer slot when a sink is installed)"""
    if g is None:
        return None
    v = _sink_view(key, g.shape, like.dtype)
    if v is not None:
        v.copy_(g)
        return v
    return g.to(like.dtype)


# ----------------------------------------------------------------------------------------------
# autograd: attention half of a block   x1 = x + dp * proj(attn(qkv(LN1(x) windows)))
# (video_swin.py:206-245, 256)
# ----------------------------------------------------------------------------------------------
class _AttnBranch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g1, b1, wqkv, bqkv, table, wproj, bproj, rowscale, plan: WindowPlan, rowcode, colcode,
                dense_mask, nH, scale, cfg_window=None, use_region=True):
        B, T, C = x.shape
        nW, N = plan.nW, plan.N
        hd = C // nH
        R = nW * N
        x = _c(x)
        # wide = the residual stream is kept in a wider dtype (fp32) than the branch computes in (autocast, stage 0 of the
        # reference: LN reads fp32, `shortcut + drop_path(x)` promotes back to fp32); g1 / b1 then come in x.dtype
        wide = x.dtype != wqkv.dtype
        xw, mean, rstd = ln_fwd(x, g1, b1, plan.gather, B, T, R, C, out_dtype=wqkv.dtype)
        qkv = linear_fwd(xw.view(B * R, C), wqkv, bqkv, B * R, 3 * C, C)
        region = plan.region if (plan.shifted and dense_mask is None and use_region) else None
        o, lse = attn_fwd(qkv, table, rowcode, colcode, region, dense_mask, B * nW, nW, N, nH, hd, scale, window=cfg_window)
        if wide:
            y = linear_fwd(o, wproj, bproj, B * R, C, C)
            x1 = residual_add(x, y, plan.gather, rowscale, B, R, T, C)
        else:
            x1 = linear_fwd(o, wproj, bproj, B * R, C, C, epi=L.EPI_RESIDUAL, res=x, rowmap=plan.gather,
                            rowscale=rowscale, rows_per_batch=R, dst_rows_per_batch=T, out_rows=B * T).view(B, T, C)
        ctx.save_for_backward(x, g1, wqkv, table, wproj, rowscale, xw, mean, rstd, qkv, o, lse, rowcode, colcode,
                              dense_mask)
        ctx.plan, ctx.nH, ctx.scale, ctx.cfg_window = plan, nH, scale, cfg_window
        ctx.use_region = use_region
        ctx.has_qkv_bias = bqkv is not None
        ctx.keys = (_key(g1), _key(b1), _key(bqkv), _key(bproj))
        return x1

    @staticmethod
    def backward(ctx, dx1):
        (x, g1, wqkv, table, wproj, rowscale, xw, mean, rstd, qkv, o, lse, rowcode, colcode, dense_mask) = ctx.saved_tensors
        plan, nH, scale = ctx.plan, ctx.nH, ctx.scale
        B, T, C = x.shape
        nW, N = plan.nW, plan.N
        hd = C // nH
        R = nW * N
        M = B * R
        dx1 = _c(dx1)
        cd = wproj.dtype
        dx1c = dx1 if dx1.dtype == cd else dx1.to(cd)     # wide residual stream: the branch sees the 16-bit rounding of dx1
        # proj: A = gather(dx1) * rowscale ; dO = A Wp ; dWp = A^T O ; dbp = sum A
        a_buf = _empty((M, C), cd, x.device)
        dO = linear_dgrad(dx1c, wproj, M, C, C, a_rowmap=plan.gather, a_rowscale=rowscale, rows_per_batch=R,
                          src_rows_per_batch=T, a_out=a_buf)
        del dx1c
        kg1, kb1, kbq, kbp = ctx.keys
        dwp, dbp = linear_wgrad(a_buf, o, M, C, C, w_key=_key(wproj), b_key=kbp)
        del a_buf
        region = plan.region if (plan.shifted and dense_mask is None and ctx.use_region) else None
        dqkv, dtable = attn_bwd(qkv, o, dO, lse, table, rowcode, colcode, region, dense_mask, B * nW, nW, N, nH, hd, scale,
                                planes=plan.ws[0], window=ctx.cfg_window)
        del dO
        dxw = linear_dgrad(dqkv, wqkv, M, 3 * C, C)
        dwq, dbq = linear_wgrad(dqkv, xw.view(M, C), M, 3 * C, C, need_bias=ctx.has_qkv_bias, w_key=_key(wqkv), b_key=kbq)
        del dqkv
        dx, dg1, db1 = ln_bwd(dxw, x, g1, mean, rstd, plan.gather, dx1, B, T, R, C, grad_dtype=g1.dtype, g_key=kg1, b_key=kb1)
        return (dx, dg1, db1, dwq, dbq, _grad_to(dtable, table, _key(table)), dwp, dbp,
                None, None, None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------------
# autograd: MLP half of a block   out = x + dp * fc2(gelu(fc1(LN2(x))))   (video_swin.py:247-248, 261)
# ----------------------------------------------------------------------------------------------
class _MlpBranch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g2, b2, w1, bb1, w2, bb2, rowscale):
        B, T, C = x.shape
        Hd = w1.shape[0]
        M = B * T
        x = _c(x)
        wide = x.dtype != w1.dtype       # fp32 residual stream around a 16-bit branch (see _AttnBranch)
        n2, mean, rstd = ln_fwd(x, g2, b2, None, B, T, T, C, out_dtype=w1.dtype)
        need_grad = any(ctx.needs_input_grad)
        u = _empty((M, Hd), w1.dtype, x.device) if need_grad else None   # gelu'(fc1 pre-activation), consumed by the backward
        g = linear_fwd(n2.view(M, C), w1, bb1, M, Hd, C, epi=L.EPI_GELU_GRAD if need_grad else L.EPI_GELU, aux_out=u)
        if wide:
            y = linear_fwd(g, w2, bb2, M, C, Hd)
            out = residual_add(x, y, None, rowscale, B, T, T, C)
        else:
            out = linear_fwd(g, w2, bb2, M, C, Hd, epi=L.EPI_RESIDUAL, res=x, rowscale=rowscale, rows_per_batch=T,
                             dst_rows_per_batch=T).view(B, T, C)
        if need_grad:
            ctx.save_for_backward(x, g2, w1, w2, rowscale, n2, mean, rstd, u, g)
            ctx.keys = (_key(g2), _key(b2), _key(bb1), _key(bb2))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, g2, w1, w2, rowscale, n2, mean, rstd, u, g = ctx.saved_tensors
        B, T, C = x.shape
        Hd = w1.shape[0]
        M = B * T
        dout = _c(dout)
        cd = w2.dtype
        doutc = dout if dout.dtype == cd else dout.to(cd)
        a_buf = _empty((M, C), cd, x.device) if rowscale is not None else None
        du = linear_dgrad(doutc.view(M, C), w2, M, C, Hd, a_rowscale=rowscale, rows_per_batch=T, src_rows_per_batch=T,
                          a_out=a_buf, mul=u)
        kg2, kb2, kbb1, kbb2 = ctx.keys
        dw2, db2 = linear_wgrad(a_buf if a_buf is not None else doutc.view(M, C), g, M, C, Hd, w_key=_key(w2), b_key=kbb2)
        del a_buf, doutc
        dn2 = linear_dgrad(du, w1, M, Hd, C)
        dw1, db1 = linear_wgrad(du, n2.view(M, C), M, Hd, C, w_key=_key(w1), b_key=kbb1)
        del du
        dx, dg2, dbeta2 = ln_bwd(dn2, x, g2, mean, rstd, None, dout, B, T, T, C, grad_dtype=g2.dtype, g_key=kg2, b_key=kb2)
        return dx, dg2, dbeta2, dw1, db1, dw2, db2, None


# ----------------------------------------------------------------------------------------------
# autograd: standalone WindowAttention3D.forward on pre-partitioned windows (video_swin.py:147-172)
# ----------------------------------------------------------------------------------------------
class _WindowAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xw, wqkv, bqkv, table, wproj, bproj, rowcode, colcode, dense_mask, nW, nH, scale, cfg_window=None):
        B_, N, C = xw.shape
        hd = C // nH
        M = B_ * N
        xw = _c(xw)
        qkv = linear_fwd(xw.view(M, C), wqkv, bqkv, M, 3 * C, C)
        o, lse = attn_fwd(qkv, table, rowcode, colcode, None, dense_mask, B_, nW, N, nH, hd, scale, window=cfg_window)
        y = linear_fwd(o, wproj, bproj, M, C, C).view(B_, N, C)
        ctx.save_for_backward(xw, wqkv, table, wproj, qkv, o, lse, rowcode, colcode, dense_mask)
        ctx.nW, ctx.nH, ctx.scale, ctx.has_qkv_bias, ctx.cfg_window = nW, nH, scale, bqkv is not None, cfg_window
        return y

    @staticmethod
    def backward(ctx, dy):
        xw, wqkv, table, wproj, qkv, o, lse, rowcode, colcode, dense_mask = ctx.saved_tensors
        B_, N, C = xw.shape
        nH = ctx.nH
        hd = C // nH
        M = B_ * N
        dy = _c(dy).view(M, C)
        dO = linear_dgrad(dy, wproj, M, C, C)
        dwp, dbp = linear_wgrad(dy, o, M, C, C)
        dqkv, dtable = attn_bwd(qkv, o, dO, lse, table, rowcode, colcode, None, dense_mask, B_, ctx.nW, N, nH, hd, ctx.scale,
                                window=ctx.cfg_window)
        dxw = linear_dgrad(dqkv, wqkv, M, 3 * C, C).view(B_, N, C)
        dwq, dbq = linear_wgrad(dqkv, xw.view(M, C), M, 3 * C, C, need_bias=ctx.has_qkv_bias)
        return dxw, dwq, dbq, _grad_to(dtable, table), dwp, dbp, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# autograd: Mlp on arbitrary (..., C) input, LayerNorm, Linear  (stand-alone module use)
# ----------------------------------------------------------------------------------------------
class _Mlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, w1, b1, w2, b2):
        M, C = x2d.shape
        Hd, Co = w1.shape[0], w2.shape[0]
        x2d = _c(x2d)
        u = _empty((M, Hd), x2d.dtype, x2d.device)   # gelu'(fc1 pre-activation)
        g = linear_fwd(x2d, w1, b1, M, Hd, C, epi=L.EPI_GELU_GRAD, aux_out=u)
        y = linear_fwd(g, w2, b2, M, Co, Hd)
        ctx.save_for_backward(x2d, w1, w2, u, g)
        ctx.has_bias = (b1 is not None, b2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, w1, w2, u, g = ctx.saved_tensors
        M, C = x2d.shape
        Hd, Co = w1.shape[0], w2.shape[0]
        dy = _c(dy)
        du = linear_dgrad(dy, w2, M, Co, Hd, mul=u)
        dw2, db2 = linear_wgrad(dy, g, M, Co, Hd, need_bias=ctx.has_bias[1])
        dx = linear_dgrad(du, w1, M, Hd, C)
        dw1, db1 = linear_wgrad(du, x2d, M, Hd, C, need_bias=ctx.has_bias[0])
        return dx, dw1, db1, dw2, db2


class _LayerNorm(torch.autograd.Function):
    """LayerNorm over the last dim of (B,T,C); ``out_dtype`` lets the final norm return fp32 under autocast."""

    @staticmethod
    def forward(ctx, x, gamma, beta, out_dtype):
        B, T, C = x.shape
        x = _c(x)
        y, mean, rstd = ln_fwd(x, gamma, beta, None, B, T, T, C, out_dtype=out_dtype)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        B, T, C = x.shape
        dy = _c(dy)
        if dy.dtype not in (x.dtype, torch.float32):
            dy = dy.to(x.dtype)
        dx, dg, db = ln_bwd(dy, x, gamma, mean, rstd, None, None, B, T, T, C, need_dx=ctx.needs_input_grad[0])
        return dx, _grad_to(dg, gamma), _grad_to(db, gamma), None


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, w, b):
        M, K = x2d.shape
        x2d = _c(x2d)
        y = linear_fwd(x2d, w, b, M, w.shape[0], K)
        ctx.save_for_backward(x2d, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, w = ctx.saved_tensors
        M, K = x2d.shape
        N = w.shape[0]
        dy = _c(dy)
        dx = linear_dgrad(dy, w, M, N, K) if ctx.needs_input_grad[0] else None
        dw, db = linear_wgrad(dy, x2d, M, N, K, need_bias=ctx.has_bias)
        return dx, dw, db


# ----------------------------------------------------------------------------------------------
# autograd: PatchMerging (video_swin.py:273-289) and PatchEmbed3D (video_swin.py:390-407)
# ----------------------------------------------------------------------------------------------
class _PatchMerge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, wred, grid: Triple):
        B, T, C = x.shape
        D, H, W = grid
        T2 = D * ((H + 1) // 2) * ((W + 1) // 2)
        x = _c(x)
        m4 = merge_map(grid, x.device)
        y4 = _empty((B, T2, 4 * C), x.dtype, x.device)
        mean = _empty((B * T2,), torch.float32, x.device)
        rstd = _empty((B * T2,), torch.float32, x.device)
        L.check(L.lib().vsw_merge_ln_fwd(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(m4), L.ptr(y4), L.ptr(mean),
                                         L.ptr(rstd), B, T, T2, C, LN_EPS, L.dt(x), L.stream()), "vsw_merge_ln_fwd")
        y = linear_fwd(y4.view(B * T2, 4 * C), wred, None, B * T2, 2 * C, 4 * C).view(B, T2, 2 * C)
        ctx.save_for_backward(x, gamma, wred, y4, mean, rstd, m4)
        ctx.grid = grid
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, wred, y4, mean, rstd, m4 = ctx.saved_tensors
        B, T, C = x.shape
        T2 = y4.shape[1]
        M = B * T2
        dy = _c(dy).view(M, 2 * C)
        dy4 = linear_dgrad(dy, wred, M, 2 * C, 4 * C)
        dw, _ = linear_wgrad(dy, y4.view(M, 4 * C), M, 2 * C, 4 * C, need_bias=False)
        # odd H/W: padded input tokens do not exist, every real token is written exactly once
        dx = _empty((B, T, C), x.dtype, x.device)
        dg = _empty((4 * C,), torch.float32, x.device)
        db = _empty((4 * C,), torch.float32, x.device)
        wsb = int(L.lib().vsw_ln_bwd_workspace(4 * C))
        ws = _empty((wsb,), torch.uint8, x.device)
        L.check(L.lib().vsw_merge_ln_bwd(L.ptr(dy4), L.ptr(x), L.ptr(gamma), L.ptr(mean), L.ptr(rstd), L.ptr(m4),
                                         L.ptr(dx), L.ptr(dg), L.ptr(db), B, T, T2, C, L.dt(x), L.ptr(ws), wsb,
                                         L.stream()), "vsw_merge_ln_bwd")
        return dx, _grad_to(dg, gamma), _grad_to(db, gamma), dw, None


class _PatchEmbed(torch.autograd.Function):
    """x (B,Cin,D,H,W) any float dtype -> tokens (B, Dout*Hp*Wp, E) in the compute dtype of ``w``."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, patch: Triple, out_dtype=None):
        B, Cin, D, H, W = x.shape
        pd, ph, pw = patch
        E = w.shape[0]
        Dout, Hp, Wp = D + 2 - pd, -(-H // ph), -(-W // pw)
        T = Dout * Hp * Wp
        Kv = Cin * pd * ph * pw
        x = _c(x)
        cdt = w.dtype
        col = _empty((B * T, Kv), cdt, x.device)
        L.check(L.lib().vsw_patch_im2col(L.ptr(x), L.ptr(col), B, Cin, D, H, W, pd, ph, pw, L.dt(x), L.dt(cdt),
                                         L.stream()), "vsw_patch_im2col")
        w2d = w.reshape(E, Kv)
        y = linear_fwd(col, w2d, b, B * T, E, Kv).view(B, T, E)
        if gamma is not None:
            # out_dtype fp32: the patch norm's output under autocast (LayerNorm is an fp32 op there, video_swin.py:401-405)
            yn, mean, rstd = ln_fwd(y, gamma, beta, None, B, T, T, E, out_dtype=out_dtype)
        else:
            yn, mean, rstd = y, None, None
        ctx.save_for_backward(col, w2d, y if gamma is not None else None, gamma, mean, rstd)
        ctx.meta = (x.shape, x.dtype, patch, w.shape)
        return yn

    @staticmethod
    def backward(ctx, dyn):
        col, w2d, y, gamma, mean, rstd = ctx.saved_tensors
        xshape, xdtype, patch, wshape = ctx.meta
        B, Cin, D, H, W = xshape
        pd, ph, pw = patch
        E, Kv = w2d.shape
        M = col.shape[0]
        T = M // B
        dyn = _c(dyn)
        dg = dbeta = None
        if gamma is not None:
            dy, dg, dbeta = ln_bwd(dyn, y, gamma, mean, rstd, None, None, B, T, T, E)
        else:
            dy = dyn
        dy2 = dy.view(M, E)
        dw, db = linear_wgrad(dy2, col, M, E, Kv)
        dx = None
        if ctx.needs_input_grad[0]:
            dcol = linear_dgrad(dy2, w2d, M, E, Kv)
            dx = _empty(xshape, xdtype, col.device)
            L.check(L.lib().vsw_patch_col2im(L.ptr(dcol), L.ptr(dx), B, Cin, D, H, W, pd, ph, pw, L.dt(xdtype),
                                             L.dt(col), L.stream()), "vsw_patch_col2im")
        return (dx, dw.view(wshape), db, _grad_to(dg, gamma) if gamma is not None else None,
                _grad_to(dbeta, gamma) if gamma is not None else None, None, None)


# ----------------------------------------------------------------------------------------------
# autograd: EncVideo tail -- class row + position / frame-length|order embeddings + LayerNorm + mask
# (model.py:57-76)
# ----------------------------------------------------------------------------------------------
class _EncVideoTail(torch.autograd.Function):
    """f (B,Tn,hw,C) -> (out (B, Tn*(1+hw), C), m_img (B, Tn*(1+hw)) int64).  The six small parameter tensors go to the
    kernels as fp32 (their gradients come back in the parameters' own dtype)."""

    @staticmethod
    def forward(ctx, f, emb_cls, emb_pos, emb_len, emb_odr, gamma, beta, odr, vt_mask, out_dtype):
        B, Tn, hw, C = f.shape
        P = hw + 1
        f = _c(f)
        params = [t.detach().reshape(-1, C).float().contiguous() for t in (emb_cls, emb_pos, emb_len, emb_odr, gamma, beta)]
        cls32, pos32, len32, odr32, g32, b32 = params
        pos_rows, len_rows = pos32.shape[0], len32.shape[0]
        out_dtype = out_dtype or f.dtype
        out = _empty((B, Tn * P, C), out_dtype, f.device)
        m_img = _empty((B, Tn * P), torch.int64, f.device)
        mean = _empty((B * Tn * P,), torch.float32, f.device)
        rstd = _empty((B * Tn * P,), torch.float32, f.device)
        t0 = PROFILER.begin() if PROFILER is not None else None
        L.check(L.lib().vsw_enc_video_tail_fwd(L.ptr(f), L.ptr(cls32), L.ptr(pos32), L.ptr(len32), L.ptr(odr32), L.ptr(odr),
                                               L.ptr(g32), L.ptr(b32), L.ptr(vt_mask), L.ptr(out), L.ptr(m_img), L.ptr(mean),
                                               L.ptr(rstd), B, Tn, hw, C, pos_rows, len_rows, LN_EPS, L.dt(f),
                                               L.dt(out_dtype), L.stream()), "vsw_enc_video_tail_fwd")
        if t0 is not None:
            PROFILER.end("enc_video_tail_fwd", t0, 0.0, B * Tn * C * (hw * _esz(f) + P * out.element_size()))
        ctx.save_for_backward(f, cls32, pos32, len32, odr32, g32, odr, mean, rstd)
        ctx.param_meta = [(t.shape, t.dtype) for t in (emb_cls, emb_pos, emb_len, emb_odr, gamma, beta)]
        ctx.mark_non_differentiable(m_img)
        return out, m_img

    @staticmethod
    def backward(ctx, dout, _dm):
        f, cls32, pos32, len32, odr32, g32, odr, mean, rstd = ctx.saved_tensors
        B, Tn, hw, C = f.shape
        P = hw + 1
        pos_rows, len_rows = pos32.shape[0], len32.shape[0]
        dout = _c(dout)
        if dout.dtype not in (f.dtype, torch.float32):
            dout = dout.to(f.dtype)
        df = torch.empty_like(f) if ctx.needs_input_grad[0] else None
        dev = f.device
        grads = [_empty(s, torch.float32, dev) for s in ((C,), (pos_rows, C), (len_rows, C), (C,), (C,), (C,))]
        wsb = int(L.lib().vsw_enc_video_tail_bwd_workspace(B, Tn, hw, C))
        ws = _empty((max(wsb, 4),), torch.uint8, dev)
        t0 = PROFILER.begin() if PROFILER is not None else None
        L.check(L.lib().vsw_enc_video_tail_bwd(L.ptr(dout), L.ptr(f), L.ptr(cls32), L.ptr(pos32), L.ptr(len32), L.ptr(odr32),
                                               L.ptr(odr), L.ptr(g32), L.ptr(mean), L.ptr(rstd), L.ptr(df),
                                               *[L.ptr(g) for g in grads], B, Tn, hw, C, pos_rows, len_rows, L.dt(f),
                                               L.dt(dout), L.ptr(ws), wsb, L.stream()), "vsw_enc_video_tail_bwd")
        if t0 is not None:
            PROFILER.end("enc_video_tail_bwd", t0, 0.0, B * Tn * C * (P * dout.element_size() + 2 * hw * _esz(f)))
        pg = [g.view(shape).to(dtype) for g, (shape, dtype) in zip(grads, ctx.param_meta)]
        if odr is None:
            pg[3] = None   # emb_odr took no part (model.py:67): leave its .grad None like the reference, so optimizers skip it
        return (df, *pg, None, None, None)


# ----------------------------------------------------------------------------------------------
# MVM masking / loss on either side of the encoder (main_pretrain.py:355-362, 520-522)
# ----------------------------------------------------------------------------------------------
def block_mask_apply(img, cov, patch_size, inplace=False, want_mask=False):
    """img (B,T,Cin,H,W) *= 1 - cov[b,t,y/ps,x/ps]; cov (B,T,h,w) uint8 on the device.  Returns (masked clip,
    mvm_mask (B,T,Cin,H,W) fp32 or None)."""
    B, Tn, Cin, H, W = img.shape
    img = _c(img)
    cov = _c(cov)
    if cov.dtype != torch.uint8 or tuple(cov.shape) != (B, Tn, H // patch_size, W // patch_size):
        raise L.VswError(f"block_mask_apply: cov must be uint8 {(B, Tn, H // patch_size, W // patch_size)}, "
                         f"got {cov.dtype} {tuple(cov.shape)}")
    out = img if inplace else torch.empty_like(img)
    mask = _empty(img.shape, torch.float32, img.device) if want_mask else None
    t0 = PROFILER.begin() if PROFILER is not None else None
    L.check(L.lib().vsw_block_mask_apply(L.ptr(img), L.ptr(cov), L.ptr(out), L.ptr(mask), B * Tn, Cin, H, W,
                                         int(patch_size), L.dt(img), L.stream()), "vsw_block_mask_apply")
    if t0 is not None:
        PROFILER.end("block_mask_apply", t0, 0.0, img.numel() * (2 * _esz(img) + (4 if want_mask else 0)))
    return out, mask


class _MaskedL1(torch.autograd.Function):
    """loss = sum |pred - target| * w[row] / (sum w + 1e-5) / in_c  -- fp32 device scalar; gradient to pred only"""

    @staticmethod
    def forward(ctx, pred2d, target2d, row_weight, in_c):
        rows, C = pred2d.shape
        pred2d, target2d = _c(pred2d), _c(target2d)
        w = _c(row_weight.reshape(rows).float())
        out = _empty((2,), torch.float32, pred2d.device)       # [loss, sum w]
        wsb = int(L.lib().vsw_masked_l1_workspace())
        ws = _empty((wsb,), torch.uint8, pred2d.device)
        t0 = PROFILER.begin() if PROFILER is not None else None
        L.check(L.lib().vsw_masked_l1_fwd(L.ptr(pred2d), L.ptr(target2d), L.ptr(w), out.data_ptr(), out.data_ptr() + 4,
                                          rows, C, float(in_c), L.dt(pred2d), L.dt(target2d), L.ptr(ws), wsb, L.stream()),
                "vsw_masked_l1_fwd")
        if t0 is not None:
            PROFILER.end("masked_l1_fwd", t0, 0.0, rows * C * (_esz(pred2d) + _esz(target2d)))
        ctx.save_for_backward(pred2d, target2d, w, out)
        ctx.in_c = float(in_c)
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        pred2d, target2d, w, out = ctx.saved_tensors
        rows, C = pred2d.shape
        dl = _c(dloss.reshape(1).float())
        dpred = torch.empty_like(pred2d)
        L.check(L.lib().vsw_masked_l1_bwd(L.ptr(pred2d), L.ptr(target2d), L.ptr(w), out.data_ptr() + 4, L.ptr(dl),
                                          L.ptr(dpred), rows, C, ctx.in_c, L.dt(pred2d), L.dt(target2d), L.stream()),
                "vsw_masked_l1_bwd")
        return dpred, None, None, None


# public functional entry points --------------------------------------------------------------
def attn_branch(x, g1, b1, wqkv, bqkv, table, wproj, bproj, rowscale, plan, rowcode, colcode, dense_mask, nH, scale,
                cfg_window=None, use_region=True):
    """cfg_window: the module's CONFIGURED window (whose relative_position_index produced the codes) -- a layout hint.
    use_region=False: no shift mask at all (the reference's `mask_matrix=None` on a shifted block, video_swin.py:231-235)"""
    return _AttnBranch.apply(x, g1, b1, wqkv, bqkv, table, wproj, bproj, rowscale, plan, rowcode, colcode, dense_mask,
                             nH, scale, cfg_window, use_region)


def mlp_branch(x, g2, b2, w1, bb1, w2, bb2, rowscale):
    return _MlpBranch.apply(x, g2, b2, w1, bb1, w2, bb2, rowscale)


def window_attention(xw, wqkv, bqkv, table, wproj, bproj, rowcode, colcode, dense_mask, nW, nH, scale, cfg_window=None):
    return _WindowAttention.apply(xw, wqkv, bqkv, table, wproj, bproj, rowcode, colcode, dense_mask, nW, nH, scale,
                                  cfg_window)


def mlp(x2d, w1, b1, w2, b2):
    return _Mlp.apply(x2d, w1, b1, w2, b2)


def layer_norm(x3d, gamma, beta, out_dtype=None):
    return _LayerNorm.apply(x3d, gamma, beta, out_dtype)


def linear(x2d, w, b):
    return _Linear.apply(x2d, w, b)


def patch_merge(x, gamma, beta, wred, grid):
    return _PatchMerge.apply(x, gamma, beta, wred, grid)


def patch_embed(x, w, b, gamma, beta, patch, out_dtype=None):
    return _PatchEmbed.apply(x, w, b, gamma, beta, patch, out_dtype)


def enc_video_tail(f, emb_cls, emb_pos, emb_len, emb_odr, gamma, beta, odr=None, vt_mask=None, out_dtype=None):
    """odr: (B,Tn) int32 device tensor or None; vt_mask: (B,Tn,1+hw) int64 device tensor or None"""
    return _EncVideoTail.apply(f, emb_cls, emb_pos, emb_len, emb_odr, gamma, beta, odr, vt_mask, out_dtype)


def masked_l1(pred2d, target2d, row_weight, in_c=3.0):
    """main_pretrain.py:520-522: sum(|pred - target| * w) / (sum(w) + 1e-5) / in_c; target in pred's dtype or fp32"""
    return _MaskedL1.apply(pred2d, target2d, row_weight, in_c)
