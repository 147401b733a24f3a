"""Times the MVM kernels (csrc/mvm_kernels.cu) at the VIOLET step size: patch masking of a 32 x 8 x 3 x 224^2 clip and the
masked L1 against 32 x 8 x 49 teacher tokens of 1024 channels.  Direct C-ABI calls, L2 flushed before every launch, CUDA
events on the launching stream (see scripts/bench_enc_tail.py).  Algorithmic bytes: masking = clip read + clip written
(+ fp32 mask when materialised); loss forward = pred + target rows WITH non-zero weight only; backward = those + dpred.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
L = vsw._lib


def timed(fn, flush, iters=20):
    ev = []
    for _ in range(iters + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev[3:])
    return t[len(t) // 2]


def main():
    dev = "cuda"
    lib, st = L.lib(), L.stream()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    res = {"hbm_peak_GBps": json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs")}
    B, Tn, H, W = 32, 8, 224, 224
    np.random.seed(0)
    cov = torch.from_numpy(vsw.mvm.sample_block_masks(B, Tn, 7, 7)).to(dev)
    for dtype in (torch.float32, torch.bfloat16):
        img = torch.randn(B, Tn, 3, H, W, device=dev).to(dtype)
        out = torch.empty_like(img)
        mask = torch.empty(img.shape, dtype=torch.float32, device=dev)
        nb = img.numel() * img.element_size()
        t1 = timed(lambda: L.check(lib.vsw_block_mask_apply(L.ptr(img), L.ptr(cov), L.ptr(out), None, B * Tn, 3, H, W, 32, L.dt(dtype), st)), flush)
        t2 = timed(lambda: L.check(lib.vsw_block_mask_apply(L.ptr(img), L.ptr(cov), L.ptr(out), L.ptr(mask), B * Tn, 3, H, W, 32, L.dt(dtype), st)), flush)
        res[f"block_mask_{str(dtype)[6:]}"] = dict(clip_only_ms=round(t1, 4), clip_only_GBps=round(2 * nb / t1 / 1e6, 1),
                                                   with_mask_ms=round(t2, 4), with_mask_GBps=round((2 * nb + mask.numel() * 4) / t2 / 1e6, 1))
    rows, C = B * Tn * 49, 1024
    m = cov.reshape(-1).float()
    live = int(m.sum())
    ws = torch.empty(int(lib.vsw_masked_l1_workspace()), dtype=torch.uint8, device=dev)
    sc = torch.empty(2, device=dev)
    dl = torch.ones(1, device=dev)
    for dtype in (torch.bfloat16, torch.float32):
        pred = torch.randn(rows, C, device=dev).to(dtype)
        tgt = torch.randn(rows, C, device=dev).to(dtype)
        dp = torch.empty_like(pred)
        es = pred.element_size()
        tf = timed(lambda: L.check(lib.vsw_masked_l1_fwd(L.ptr(pred), L.ptr(tgt), L.ptr(m), sc.data_ptr(), sc.data_ptr() + 4, rows, C, 3.0,
                                                         L.dt(dtype), L.dt(dtype), L.ptr(ws), ws.numel(), st)), flush)
        tb = timed(lambda: L.check(lib.vsw_masked_l1_bwd(L.ptr(pred), L.ptr(tgt), L.ptr(m), sc.data_ptr() + 4, L.ptr(dl), L.ptr(dp), rows, C, 3.0,
                                                         L.dt(dtype), L.dt(dtype), st)), flush)
        fb, bb = 2 * live * C * es, (2 * live + rows) * C * es
        res[f"masked_l1_{str(dtype)[6:]}"] = dict(rows=rows, live_rows=live, fwd_ms=round(tf, 4), fwd_GBps=round(fb / tf / 1e6, 1),
                                                  bwd_ms=round(tb, 4), bwd_GBps=round(bb / tb / 1e6, 1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
