// tcgen05 window attention -- placeholder until the TMEM kernels land.
#include "attn.cuh"
namespace vsw {
int tc_attn_fwd(const void*, const void*, const int32_t*, const int32_t*, const uint8_t*, void*, float*, int, int, int,
                int, int, int, float, cudaStream_t) { return VSW_ERR_UNSUPPORTED; }
size_t tc_attn_bwd_workspace(int, int, int, int, int) { return 0; }
int tc_attn_bwd(const void*, const void*, const void*, const float*, const void*, const int32_t*, const int32_t*,
                const uint8_t*, void*, float*, int, int, int, int, int, int, float, void*, size_t, cudaStream_t) {
    return VSW_ERR_UNSUPPORTED;
}
}  // namespace vsw
