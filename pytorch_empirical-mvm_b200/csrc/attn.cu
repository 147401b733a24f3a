// C-ABI entry points of fused window attention; routes to the tcgen05 or CUDA-core family.
// Reference: WindowAttention3D.forward, visbackbone/video_swin.py:149-169.
#include <stdlib.h>
#include "attn.cuh"

namespace vsw {
int backend();
static bool attn_want_tc(int dtype, int hd, const void* dmask) {
    if (dtype != VSW_BF16 || hd != 32 || dmask) return false;
    const int b = backend();
    return b == VSW_GEMM_TCGEN05 || b == VSW_GEMM_AUTO;
}
static bool attn_want_tc2(int dtype, int hd, const void* dmask, bool bwd) {
    // which generation of the tcgen05 kernel bf16 takes: VSW_ATTN_TC2 is a bit mask read once per process (bit 0: forward,
    // bit 1: backward use the second generation; default 3 = both).  fp16 only exists in the second generation.
    static const int pref = getenv("VSW_ATTN_TC2") ? atoi(getenv("VSW_ATTN_TC2")) : 3;
    if ((dtype != VSW_BF16 && dtype != VSW_F16) || hd != 32 || dmask) return false;
    if (dtype == VSW_BF16 && !(pref & (bwd ? 2 : 1))) return false;
    const int b = backend();
    return b == VSW_GEMM_TCGEN05 || b == VSW_GEMM_AUTO;
}
}  // namespace vsw
using namespace vsw;

#define VSW_ATTN_CHECK(name)                                                                                   \
    VSW_REQUIRE(qkv && bias_table && rowcode && colcode && B_ > 0 && nW > 0 && N > 0 && nH > 0 && hd > 0 && L > 0, \
                VSW_ERR_ARG, name ": bad args");                                                               \
    VSW_REQUIRE(B_ % nW == 0, VSW_ERR_ARG, name ": B_=%d not a multiple of nW=%d", B_, nW)

extern "C" int vsw_window_attn_fwd(const void* qkv, const void* bias_table, const int32_t* rowcode,
                                   const int32_t* colcode, const uint8_t* region, const void* dense_mask, void* out,
                                   float* lse, int B_, int nW, int N, int nH, int hd, int L, float scale,
                                   int window_dims, int dtype, void* stream) {
    VSW_ATTN_CHECK("vsw_window_attn_fwd");
    VSW_REQUIRE(out && lse, VSW_ERR_ARG, "vsw_window_attn_fwd: out/lse NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (attn_want_tc2(dtype, hd, dense_mask, false)) {
        int rc = tc2_attn_fwd(qkv, bias_table, rowcode, colcode, region, out, lse, B_, nW, N, nH, hd, L, scale, window_dims, dtype, st);
        if (rc != VSW_ERR_UNSUPPORTED) return rc;
    }
    if (attn_want_tc(dtype, hd, dense_mask)) {
        int rc = tc_attn_fwd(qkv, bias_table, rowcode, colcode, region, out, lse, B_, nW, N, nH, hd, L, scale, window_dims, st);
        if (rc != VSW_ERR_UNSUPPORTED || backend() == VSW_GEMM_TCGEN05) return rc;
    }
    return simt_attn_fwd(qkv, bias_table, rowcode, colcode, region, dense_mask, out, lse, B_, nW, N, nH, hd, L, scale,
                         dtype, st);
}

extern "C" size_t vsw_window_attn_bwd_workspace(int B_, int N, int nH, int hd, int L) {
    size_t a = simt_attn_bwd_workspace(B_, N, nH, hd, L), b = tc_attn_bwd_workspace(B_, N, nH, hd, L);
    const size_t c = tc2_attn_bwd_workspace(B_, N, nH, hd, L);
    if (c > b) b = c;
    return a > b ? a : b;
}

extern "C" int vsw_window_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                                   const void* bias_table, const int32_t* rowcode, const int32_t* colcode,
                                   const uint8_t* region, const void* dense_mask, void* dqkv, float* dbias_table,
                                   int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims,
                                   int dtype, void* ws, size_t ws_bytes, void* stream) {
    VSW_ATTN_CHECK("vsw_window_attn_bwd");
    VSW_REQUIRE(out && dout && lse && dqkv && dbias_table && ws, VSW_ERR_ARG, "vsw_window_attn_bwd: NULL pointer");
    VSW_REQUIRE(ws_bytes >= vsw_window_attn_bwd_workspace(B_, N, nH, hd, L), VSW_ERR_WORKSPACE,
                "vsw_window_attn_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (attn_want_tc2(dtype, hd, dense_mask, true)) {
        int rc = tc2_attn_bwd(qkv, out, dout, lse, bias_table, rowcode, colcode, region, dqkv, dbias_table, B_, nW, N, nH, hd, L, scale,
                              window_dims, dtype, ws, ws_bytes, st);
        if (rc != VSW_ERR_UNSUPPORTED) return rc;
    }
    if (attn_want_tc(dtype, hd, dense_mask)) {
        int rc = tc_attn_bwd(qkv, out, dout, lse, bias_table, rowcode, colcode, region, dqkv, dbias_table, B_, nW, N,
                             nH, hd, L, scale, window_dims, ws, ws_bytes, st);
        if (rc != VSW_ERR_UNSUPPORTED || backend() == VSW_GEMM_TCGEN05) return rc;
    }
    return simt_attn_bwd(qkv, out, dout, lse, bias_table, rowcode, colcode, region, dense_mask, dqkv, dbias_table, B_,
                         nW, N, nH, hd, L, scale, dtype, ws, ws_bytes, st);
}
