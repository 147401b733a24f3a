#!/usr/bin/env python
"""Micro-benchmarks of individual vsw kernels at the Swin-B (B=32) shapes; CUDA-event timing.
usage: python scripts/kbench.py [attn|linear|ln|all] [--iters N]"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib


GRAPH = False


def timeit(fn, iters=10, warm=3):
    """ms per call; with GRAPH the calls are captured into one CUDA graph so host launch overhead is excluded."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def attn_case(B, grid, nH, shifted, iters):
    window = (8, 7, 7)
    shift = (4, 3, 3) if shifted else (0, 0, 0)
    plan = VF.window_plan(grid, window, shift, "cuda")
    nW, N = plan.nW, plan.N
    B_ = B * nW
    C = nH * 32
    qkv = torch.randn(B_ * N, 3 * C, device="cuda").bfloat16()
    table = (torch.randn(2535, nH, device="cuda") * 0.1).bfloat16()
    att = vsw.WindowAttention3D(C, window, nH).cuda()
    rc, cc = att.bias_codes(N)
    region = plan.region if plan.shifted else None
    out, lse = VF.attn_fwd(qkv, table, rc, cc, region, None, B_, nW, N, nH, 32, 32 ** -0.5, window=window)
    dout = torch.randn_like(out)
    tf = timeit(lambda: VF.attn_fwd(qkv, table, rc, cc, region, None, B_, nW, N, nH, 32, 32 ** -0.5, window=window), iters)
    tb = timeit(lambda: VF.attn_bwd(qkv, out, dout, lse, table, rc, cc, region, None, B_, nW, N, nH, 32, 32 ** -0.5, planes=8, window=window), iters)
    fl = 4.0 * B_ * nH * N * N * 32
    return dict(kernel="window_attn", B_=B_, N=N, nH=nH, shifted=shifted, fwd_ms=tf, bwd_ms=tb,
                fwd_tflops=fl / tf / 1e9, bwd_tflops=2 * fl / tb / 1e9,
                fwd_us_per_item=tf * 1e3 / (B_ * nH) * 148, bwd_us_per_item=tb * 1e3 / (B_ * nH) * 148)


def linear_case(M, N, K, iters):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.zeros(N, device="cuda").bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t_f = timeit(lambda: VF.linear_fwd(x, w, b, M, N, K), iters)
    t_g = timeit(lambda: VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU, aux_out=u), iters)
    t_gg = timeit(lambda: VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU_GRAD, aux_out=u), iters)
    pre = torch.randn(M, K, device="cuda").bfloat16()
    t_dg = timeit(lambda: VF.linear_dgrad(dy, w, M, N, K, gelu_pre=pre), iters)
    t_dm = timeit(lambda: VF.linear_dgrad(dy, w, M, N, K, mul=pre), iters)
    t_d = timeit(lambda: VF.linear_dgrad(dy, w, M, N, K), iters)
    t_w = timeit(lambda: VF.linear_wgrad(dy, x, M, N, K), iters)
    t_ref = timeit(lambda: torch.matmul(x, w.t()), iters)
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    return dict(kernel="linear", M=M, N=N, K=K, fwd_ms=t_f, gelu_ms=t_g, gelu_grad_ms=t_gg, dgrad_ms=t_d, dgrad_gelu_ms=t_dg,
                dgrad_mul_ms=t_dm, wgrad_ms=t_w, cublas_ms=t_ref,
                fwd_tflops=fl / t_f / 1e9, gelu_tflops=fl / t_g / 1e9, dgrad_tflops=fl / t_d / 1e9,
                wgrad_tflops=fl / t_w / 1e9, cublas_tflops=fl / t_ref / 1e9, fwd_gbs=by / t_f / 1e6,
                t_flop_ms=fl / 1373.6e9, t_byte_ms=by / 6545.6e6)


def ln_case(B, grid, C, shifted, iters):
    window = (8, 7, 7)
    shift = (4, 3, 3) if shifted else (0, 0, 0)
    plan = VF.window_plan(grid, window, shift, "cuda")
    T = grid[0] * grid[1] * grid[2]
    Tout = plan.nW * plan.N
    x = torch.randn(B, T, C, device="cuda").bfloat16()
    g = torch.ones(C, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda").bfloat16()
    res = {}
    for name, gmap, To in (("plain", None, T), ("gather", plan.gather, Tout)):
        y, mean, rstd = VF.ln_fwd(x, g, b, gmap, B, T, To, C)
        dy = torch.randn_like(y); dres = torch.randn_like(x)
        tf = timeit(lambda: VF.ln_fwd(x, g, b, gmap, B, T, To, C), iters)
        tb = timeit(lambda: VF.ln_bwd(dy, x, g, mean, rstd, gmap, dres, B, T, To, C), iters)
        fb = 2.0 * B * C * (T + To); bb = 2.0 * B * C * (To + 3 * T)
        res[name] = dict(fwd_ms=tf, bwd_ms=tb, fwd_gbs=fb / tf / 1e6, bwd_gbs=bb / tb / 1e6)
    return dict(kernel="ln", B=B, T=T, C=C, shifted=shifted, **{f"{k}_{m}": v for k, r in res.items() for m, v in r.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays (no host launch overhead)")
    a = ap.parse_args()
    global GRAPH
    GRAPH = a.graph
    B = a.batch
    res = []
    if a.what in ("attn", "all"):
        for grid, nH in (((8, 56, 56), 4), ((8, 28, 28), 8), ((8, 14, 14), 16), ((8, 7, 7), 32)):
            for sh in (False, True):
                if sh and grid == (8, 7, 7):
                    continue
                res.append(attn_case(B, grid, nH, sh, a.iters))
                print(json.dumps(res[-1]), flush=True)
    if a.what in ("ln", "all"):
        for grid, C in (((8, 56, 56), 128), ((8, 28, 28), 256), ((8, 14, 14), 512), ((8, 7, 7), 1024)):
            res.append(ln_case(B, grid, C, grid != (8, 7, 7), a.iters))
            print(json.dumps(res[-1]), flush=True)
    if a.what in ("linear", "all"):
        T = [25088, 6272, 1568, 392]
        for s, C in enumerate((128, 256, 512, 1024)):
            M = B * T[s]
            for (N, K) in ((3 * C, C), (C, C), (4 * C, C), (C, 4 * C)):
                res.append(linear_case(M, N, K, a.iters))
                print(json.dumps(res[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"kbench_{a.what}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
