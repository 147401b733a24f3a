#!/usr/bin/env python
"""one fused-window-attention forward+backward call at the Swin-B stage-0 shape (for ncu)"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF = vsw.functional
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid, nH, window = (8, 56, 56), 4, (8, 7, 7)
plan = VF.window_plan(grid, window, (4, 3, 3), "cuda")
nW, N = plan.nW, plan.N
B_, C = B * nW, nH * 32
qkv = torch.randn(B_ * N, 3 * C, device="cuda").bfloat16()
table = (torch.randn(2535, nH, device="cuda") * 0.1).bfloat16()
att = vsw.WindowAttention3D(C, window, nH).cuda()
rc, cc = att.bias_codes(N)
for _ in range(2):
    out, lse = VF.attn_fwd(qkv, table, rc, cc, plan.region, None, B_, nW, N, nH, 32, 32 ** -0.5, window=window)
    dout = torch.randn_like(out)
    VF.attn_bwd(qkv, out, dout, lse, table, rc, cc, plan.region, None, B_, nW, N, nH, 32, 32 ** -0.5, planes=8, window=window)
torch.cuda.synchronize()
print("done")
