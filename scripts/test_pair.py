"""Quick parity + timing check of the CTA-pair GEMM variants against torch (run with VSW_GEMM_PAIR=0/1)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib
torch.manual_seed(0)
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for (M, N, K) in [(50176, 2048, 512), (50176, 512, 2048), (50176, 1536, 512), (12544, 1024, 1024), (50000, 768, 264), (200704, 1024, 256)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); dy = torch.randn(M, N, device="cuda").bfloat16()
    y = VF.linear_fwd(x, w, b, M, N, K)
    ref = (x.float() @ w.float().t() + b.float())
    dx = VF.linear_dgrad(dy, w, M, N, K)
    refdx = dy.float() @ w.float()
    torch.cuda.synchronize()
    print(M, N, K, "fwd rel", f"{rel(y, ref):.2e}", "dgrad rel", f"{rel(dx, refdx):.2e}",
          "fwd ms", f"{t(lambda: VF.linear_fwd(x, w, b, M, N, K)):.4f}", "dgrad ms", f"{t(lambda: VF.linear_dgrad(dy, w, M, N, K)):.4f}",
          "cublas ms", f"{t(lambda: torch.matmul(x, w.t())):.4f}", flush=True)
