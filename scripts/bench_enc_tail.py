"""Times the EncVideo tail kernels (csrc/enc_video.cu) at the VIOLET step size and prints achieved HBM GB/s.

    python scripts/bench_enc_tail.py [--batch 32] [--frames 8] [--hw 49] [--hidden 768]

The C ABI is called directly (ctypes) with pre-allocated buffers; every iteration is [L2 flush, event, launch(es), event]
enqueued back to back without a host sync, so the events bracket device time only (the flush kernel is longer than the
host-side enqueue).  Algorithmic bytes (each tensor once, storage dtype): forward f + out; backward dy + f + df (the fp32
dpre workspace round trip of the backward is overhead, not algorithmic traffic).
"""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
L = vsw._lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--hw", type=int, default=49)
    ap.add_argument("--hidden", type=int, default=768)
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    B, Tn, hw, C = a.batch, a.frames, a.hw, a.hidden
    P = hw + 1
    dev = "cuda"
    torch.manual_seed(0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
    res = {"shape": dict(B=B, frames=Tn, hw=hw, hidden=C), "hbm_peak_GBps": peaks.get("hbm_gbs")}
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    lib, st = L.lib(), L.stream()
    pos_rows, len_rows = 197, max(Tn, 6)
    f32 = lambda *s: 0.02 * torch.randn(*s, device=dev)
    cls, pos, ln, od, g, b = f32(C), f32(pos_rows, C), f32(len_rows, C), f32(C), 1 + f32(C), f32(C)
    odr = torch.stack([torch.randperm(Tn) for _ in range(B)]).to(torch.int32).to(dev)
    mean = torch.empty(B * Tn * P, device=dev)
    rstd = torch.empty(B * Tn * P, device=dev)
    m_img = torch.empty(B * Tn * P, dtype=torch.int64, device=dev)
    grads = [torch.empty_like(t) for t in (cls, pos, ln, od, g, b)]
    wsb = int(lib.vsw_enc_video_tail_bwd_workspace(B, Tn, hw, C))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    for dtype in (torch.bfloat16, torch.float32):
        f = torch.randn(B, Tn, hw, C, device=dev).to(dtype)
        out = torch.empty(B, Tn * P, C, device=dev, dtype=dtype)
        dy = torch.randn(B, Tn * P, C, device=dev).to(dtype)
        df = torch.empty_like(f)
        ev = []
        for it in range(a.iters + 3):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            flush.zero_()
            e[0].record()
            L.check(lib.vsw_enc_video_tail_fwd(L.ptr(f), L.ptr(cls), L.ptr(pos), L.ptr(ln), L.ptr(od), L.ptr(odr), L.ptr(g),
                                               L.ptr(b), None, L.ptr(out), L.ptr(m_img), L.ptr(mean), L.ptr(rstd), B, Tn, hw,
                                               C, pos_rows, len_rows, 1e-5, L.dt(dtype), L.dt(dtype), st))
            e[1].record()
            flush.zero_()
            e[2].record()
            L.check(lib.vsw_enc_video_tail_bwd(L.ptr(dy), L.ptr(f), L.ptr(cls), L.ptr(pos), L.ptr(ln), L.ptr(od), L.ptr(odr),
                                               L.ptr(g), L.ptr(mean), L.ptr(rstd), L.ptr(df), *[L.ptr(t) for t in grads],
                                               B, Tn, hw, C, pos_rows, len_rows, L.dt(dtype), L.dt(dtype), L.ptr(ws), wsb, st))
            e[3].record()
            ev.append(e)
        torch.cuda.synchronize()
        tf = sorted(e[0].elapsed_time(e[1]) for e in ev[3:])
        tb = sorted(e[2].elapsed_time(e[3]) for e in ev[3:])
        es = f.element_size()
        fb = B * Tn * C * (hw + P) * es
        bb = B * Tn * C * (P + 2 * hw) * es
        mf, mb = tf[len(tf) // 2], tb[len(tb) // 2]
        res[str(dtype).replace("torch.", "")] = dict(
            fwd_ms=round(mf, 4), fwd_GBps=round(fb / mf / 1e6, 1), fwd_bytes=fb,
            bwd_ms=round(mb, 4), bwd_GBps=round(bb / mb / 1e6, 1), bwd_bytes=bb, bwd_launches=3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
