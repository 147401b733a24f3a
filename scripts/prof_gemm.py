import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib
for (M, N, K) in [(802816, 384, 128), (802816, 512, 128), (50176, 1536, 512), (50176, 2048, 512), (50176, 512, 2048)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.zeros(N, device="cuda").bfloat16(); dy = torch.randn(M, N, device="cuda").bfloat16()
    u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2): VF.linear_fwd(x, w, b, M, N, K)
    for _ in range(2): VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU, aux_out=u)
    for _ in range(2): VF.linear_dgrad(dy, w, M, N, K)
    for _ in range(2): VF.linear_wgrad(dy, x, M, N, K)
VF.linear_fwd(x, w, b, M, N, K)
torch.cuda.synchronize()
