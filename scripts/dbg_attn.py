#!/usr/bin/env python
"""debug / timing driver of the fused window attention at one geometry: tcgen05 path vs the CUDA-core path.
usage: python scripts/dbg_attn.py D H W wd wh ww sd sh sw nH B [dtype] [iters]"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib
a = [int(v) for v in sys.argv[1:12]]
grid, window, shift, nH, B = tuple(a[0:3]), tuple(a[3:6]), tuple(a[6:9]), a[9], a[10]
dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[12] if len(sys.argv) > 12 else "bf16"]
iters = int(sys.argv[13]) if len(sys.argv) > 13 else 0
do_bwd = os.environ.get("DBG_BWD", "1") == "1"
plan = VF.window_plan(grid, window, shift, "cuda")
nW, N = plan.nW, plan.N
B_, C = B * nW, nH * 32
torch.manual_seed(0)
qkv = torch.randn(B_ * N, 3 * C, device="cuda").to(dtype)
Lt = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
table = (torch.randn(Lt, nH, device="cuda") * 0.5).to(dtype)
att = vsw.WindowAttention3D(C, window, nH).cuda()
rc, cc = att.bias_codes(N)
region = plan.region if plan.shifted else None
sc = 32 ** -0.5
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
L.set_gemm_backend(L.GEMM_SIMT)
o_ref, lse_ref = VF.attn_fwd(qkv.float(), table.float(), rc, cc, region, None, B_, nW, N, nH, 32, sc, window=window)
dout = torch.randn_like(o_ref).to(dtype)
if do_bwd:
    dq_ref, dt_ref = VF.attn_bwd(qkv.float(), o_ref, dout.float(), lse_ref, table.float(), rc, cc, region, None, B_, nW, N, nH, 32, sc,
                                 planes=plan.ws[0], window=window)
L.set_gemm_backend(L.GEMM_TCGEN05)
o, lse = VF.attn_fwd(qkv, table, rc, cc, region, None, B_, nW, N, nH, 32, sc, window=window)
torch.cuda.synchronize()
print(f"N={N} nW={nW} B_={B_} nH={nH} {dtype}: fwd out rel {rel(o, o_ref):.3e}  lse rel {rel(lse, lse_ref):.3e}  finite {bool(torch.isfinite(o.float()).all())}")
if do_bwd:
    dq, dt = VF.attn_bwd(qkv, o, dout, lse, table, rc, cc, region, None, B_, nW, N, nH, 32, sc, planes=plan.ws[0], window=window)
    torch.cuda.synchronize()
    d5, r5 = dq.view(B_, N, 3, nH, 32), dq_ref.view(B_, N, 3, nH, 32)
    print("   bwd: " + "  ".join(f"d{n} rel {rel(d5[:, :, i], r5[:, :, i]):.3e}" for i, n in enumerate("qkv")) + f"  dtable rel {rel(dt, dt_ref):.3e}")
if iters:
    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    tf = timeit(lambda: VF.attn_fwd(qkv, table, rc, cc, region, None, B_, nW, N, nH, 32, sc, window=window))
    items = B_ * nH
    per = tf * 1e3 / max(1, -(-items // 148))
    print(f"   fwd {tf:.4f} ms = {per:.2f} us per item per SM  ({4.0 * items * N * N * 32 / tf / 1e9:.1f} TFLOP/s)")
    if do_bwd:
        tb = timeit(lambda: VF.attn_bwd(qkv, o, dout, lse, table, rc, cc, region, None, B_, nW, N, nH, 32, sc, planes=plan.ws[0], window=window))
        print(f"   bwd {tb:.4f} ms = {tb * 1e3 / max(1, -(-items // 148)):.2f} us per item per SM  ({8.0 * items * N * N * 32 / tb / 1e9:.1f} TFLOP/s)")
