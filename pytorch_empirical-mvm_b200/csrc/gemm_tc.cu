// tcgen05 GEMM family -- placeholder until the TMA/TMEM kernels land (returns UNSUPPORTED so that
// VSW_GEMM_AUTO uses the CUDA-core kernels and VSW_GEMM_TCGEN05 fails loudly).
#include "gemm_tc.cuh"

namespace vsw {

int tc_linear(const TcLinearArgs&, cudaStream_t) { set_error("tcgen05 linear: not built"); return VSW_ERR_UNSUPPORTED; }
int tc_dgrad(const TcDgradArgs&, cudaStream_t) { set_error("tcgen05 dgrad: not built"); return VSW_ERR_UNSUPPORTED; }
size_t tc_wgrad_workspace(int, int, int) { return 0; }
int tc_wgrad(const void*, const void*, void*, int, int, int, int, void*, size_t, cudaStream_t) {
    set_error("tcgen05 wgrad: not built");
    return VSW_ERR_UNSUPPORTED;
}

}  // namespace vsw
