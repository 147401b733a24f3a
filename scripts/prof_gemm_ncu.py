"""One launch each of the stage-2 fc1 GEMMs (fwd plain, fwd+GELU, dgrad, wgrad) for an ncu capture."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vsw = importlib.import_module("pytorch_empirical-mvm_b200")
VF, L = vsw.functional, vsw._lib
M, N, K = 50176, 2048, 512
x = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
b = torch.zeros(N, device="cuda").bfloat16(); dy = torch.randn(M, N, device="cuda").bfloat16()
u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    VF.linear_fwd(x, w, b, M, N, K)
    VF.linear_fwd(x, w, b, M, N, K, epi=L.EPI_GELU, aux_out=u)
    VF.linear_dgrad(dy, w, M, N, K)
    VF.linear_wgrad(dy, x, M, N, K)
torch.cuda.synchronize()
