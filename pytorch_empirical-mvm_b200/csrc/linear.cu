// C-ABI entry points of the Linear family; routes to the tcgen05 (bf16 / fp16) or CUDA-core kernels.
// Reference ops replaced: nn.Linear qkv/proj (video_swin.py:139-141,149,170), Mlp fc1/fc2 (:70-79),
// PatchMerging.reduction (:270,287) and their autograd backward.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace vsw {
int backend();

static bool want_tc(int dtype) {
    if (dtype != VSW_BF16 && dtype != VSW_F16) return false;
    const int b = backend();
    return b == VSW_GEMM_TCGEN05 || b == VSW_GEMM_AUTO;
}
}  // namespace vsw

using namespace vsw;

extern "C" int vsw_linear_fwd(const void* x, const void* w, const void* bias, void* y, int M, int N, int K,
                              int epilogue, void* aux_out, const void* res, const int32_t* rowmap,
                              const float* rowscale, int rows_per_batch, int dst_rows_per_batch, int dtype,
                              void* stream) {
    VSW_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, VSW_ERR_ARG, "vsw_linear_fwd: bad args");
    VSW_REQUIRE(epilogue >= VSW_EPI_BIAS && epilogue <= VSW_EPI_GELU_GRAD, VSW_ERR_ARG, "vsw_linear_fwd: bad epilogue");
    VSW_REQUIRE(epilogue != VSW_EPI_GELU_GRAD || aux_out, VSW_ERR_ARG, "vsw_linear_fwd: VSW_EPI_GELU_GRAD needs aux_out");
    if (epilogue == VSW_EPI_RESIDUAL) {
        VSW_REQUIRE(res, VSW_ERR_ARG, "vsw_linear_fwd: residual epilogue needs res");
        if (rows_per_batch <= 0) { rows_per_batch = M; dst_rows_per_batch = M; }
        VSW_REQUIRE(M % rows_per_batch == 0 && dst_rows_per_batch > 0, VSW_ERR_ARG,
                    "vsw_linear_fwd: M=%d not a multiple of rows_per_batch=%d", M, rows_per_batch);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (want_tc(dtype)) {
        TcLinearArgs a{};
        a.x = x; a.w = w; a.bias = bias; a.y = y; a.M = M; a.N = N; a.K = K; a.epi = epilogue;
        a.aux_out = aux_out; a.res = res; a.rowmap = rowmap; a.rowscale = rowscale;
        a.rows_per_batch = rows_per_batch; a.dst_rows_per_batch = dst_rows_per_batch; a.gelu_pre = nullptr; a.dtype = dtype;
        int rc = tc_linear(a, st);
        if (rc != VSW_ERR_UNSUPPORTED || backend() == VSW_GEMM_TCGEN05) return rc;
        // AUTO: shapes the tcgen05 tiling cannot take (K % 8 != 0 etc.) run on the CUDA-core kernel
    }
    SimtGemmParams p{};
    p.A = x; p.B = w; p.C = y; p.M = M; p.N = N; p.K = K;
    p.sam = K; p.sak = 1; p.sbn = K; p.sbk = 1; p.ldc = N;
    p.epi = epilogue == VSW_EPI_BIAS ? SE_BIAS : (epilogue == VSW_EPI_RESIDUAL ? SE_RESIDUAL : SE_GELU);
    p.aux_is_grad = epilogue == VSW_EPI_GELU_GRAD;
    p.bias = bias; p.aux_out = aux_out; p.res = res; p.rowmap = rowmap; p.rowscale = rowscale;
    p.rows_per_batch = rows_per_batch; p.dst_rows_per_batch = dst_rows_per_batch;
    return launch_simt_gemm(p, true, true, dtype, st);
}

static int linear_dgrad_impl(const void* dy, const void* w, void* dx, int M, int N, int K, const int32_t* a_rowmap,
                             const float* a_rowscale, int rows_per_batch, int src_rows_per_batch, void* a_out,
                             const void* gelu_pre, bool pre_is_grad, int dtype, void* stream) {
    VSW_REQUIRE(dy && w && dx && M > 0 && N > 0 && K > 0, VSW_ERR_ARG, "vsw_linear_dgrad: bad args");
    if (a_rowmap || a_rowscale) {
        VSW_REQUIRE(rows_per_batch > 0 && M % rows_per_batch == 0 && src_rows_per_batch > 0, VSW_ERR_ARG,
                    "vsw_linear_dgrad: bad rows_per_batch");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (want_tc(dtype)) {
        TcDgradArgs a{};
        a.dy = dy; a.w = w; a.dx = dx; a.M = M; a.N = N; a.K = K; a.a_rowmap = a_rowmap; a.a_rowscale = a_rowscale;
        a.rows_per_batch = rows_per_batch; a.src_rows_per_batch = src_rows_per_batch; a.a_out = a_out;
        a.gelu_pre = gelu_pre; a.pre_is_grad = pre_is_grad; a.dtype = dtype;
        int rc = tc_dgrad(a, st);
        if (rc != VSW_ERR_UNSUPPORTED || backend() == VSW_GEMM_TCGEN05) return rc;
    }
    SimtGemmParams p{};
    // dx[m,k] = sum_n A[m,n] w[n,k]  ->  C (M x K), reduction over N; "B"(k,n) = w[n*K + k]
    p.A = dy; p.B = w; p.C = dx; p.M = M; p.N = K; p.K = N;
    p.sam = N; p.sak = 1; p.sbn = 1; p.sbk = K; p.ldc = K;
    p.a_rowmap = a_rowmap; p.a_rowscale = a_rowscale; p.rows_per_batch = rows_per_batch;
    p.src_rows_per_batch = src_rows_per_batch; p.a_out = a_out;
    p.epi = SE_DGRAD; p.gelu_pre = gelu_pre; p.aux_is_grad = pre_is_grad ? 1 : 0;
    return launch_simt_gemm(p, true, false, dtype, st);
}

extern "C" int vsw_linear_dgrad(const void* dy, const void* w, void* dx, int M, int N, int K, const int32_t* a_rowmap,
                                const float* a_rowscale, int rows_per_batch, int src_rows_per_batch, void* a_out,
                                const void* gelu_pre, int dtype, void* stream) {
    return linear_dgrad_impl(dy, w, dx, M, N, K, a_rowmap, a_rowscale, rows_per_batch, src_rows_per_batch, a_out, gelu_pre,
                             false, dtype, stream);
}

extern "C" int vsw_linear_dgrad_mul(const void* dy, const void* w, void* dx, int M, int N, int K, const int32_t* a_rowmap,
                                    const float* a_rowscale, int rows_per_batch, int src_rows_per_batch, void* a_out,
                                    const void* mul, int dtype, void* stream) {
    VSW_REQUIRE(mul, VSW_ERR_ARG, "vsw_linear_dgrad_mul: mul is required");
    return linear_dgrad_impl(dy, w, dx, M, N, K, a_rowmap, a_rowscale, rows_per_batch, src_rows_per_batch, a_out, mul, true,
                             dtype, stream);
}

static void wgrad_split(int M, int N, int K, int* splits, int* m_per_split) {
    // enough CTAs to fill the GPU: tiles(N,K) * splits ~ 4 waves, each split >= 256 rows
    const long long tiles = (long long)((N + 63) / 64) * ((K + 63) / 64);
    long long s = (4LL * kNumSMs + tiles - 1) / tiles;
    const long long smax = (M + 255) / 256;
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    if (s > 512) s = 512;
    int mps = (int)((M + s - 1) / s);
    mps = (mps + 15) / 16 * 16;
    *m_per_split = mps;
    *splits = (M + mps - 1) / mps;
}

extern "C" size_t vsw_linear_wgrad_workspace(int M, int N, int K) {
    int s, mps;
    wgrad_split(M, N, K, &s, &mps);
    size_t a = (size_t)s * N * K * sizeof(float);
    size_t b = colsum_ws_bytes(N);
    size_t t = tc_wgrad_workspace(M, N, K);
    size_t m = a > b ? a : b;
    return m > t ? m : t;
}

extern "C" int vsw_linear_wgrad(const void* dy, const void* x, void* dw, void* db, int M, int N, int K, int dtype,
                                int grad_dtype, void* ws, size_t ws_bytes, void* stream) {
    VSW_REQUIRE(dy && x && dw && ws && M > 0 && N > 0 && K > 0, VSW_ERR_ARG, "vsw_linear_wgrad: bad args");
    VSW_REQUIRE(ws_bytes >= vsw_linear_wgrad_workspace(M, N, K), VSW_ERR_WORKSPACE,
                "vsw_linear_wgrad: workspace %zu < %zu", ws_bytes, vsw_linear_wgrad_workspace(M, N, K));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = VSW_ERR_UNSUPPORTED;
    if (want_tc(dtype)) {
        rc = tc_wgrad(dy, x, dw, db, M, N, K, dtype, grad_dtype, ws, ws_bytes, st);   // db: fused column sums of dy
        if (rc != VSW_OK && (rc != VSW_ERR_UNSUPPORTED || backend() == VSW_GEMM_TCGEN05)) return rc;
        if (rc == VSW_OK) return VSW_OK;
    }
    if (rc == VSW_ERR_UNSUPPORTED) {
        int splits, mps;
        wgrad_split(M, N, K, &splits, &mps);
        SimtGemmParams p{};
        // dw[n,k] = sum_m dy[m,n] x[m,k] -> C (N x K), reduction over M; A(n,m) = dy[m*N+n], B(k,m) = x[m*K+k]
        p.A = dy; p.B = x; p.C = nullptr; p.M = N; p.N = K; p.K = M;
        p.sam = 1; p.sak = N; p.sbn = 1; p.sbk = K; p.ldc = K;
        p.epi = SE_PARTIAL; p.ksplit = splits; p.k_per_split = mps; p.partial = (float*)ws;
        rc = launch_simt_gemm(p, false, false, dtype, st);
        if (rc) return rc;
        rc = launch_partial_reduce((const float*)ws, splits, (long long)N * K, dw, grad_dtype, st);
        if (rc) return rc;
    }
    if (db) return launch_colsum(dy, M, N, db, dtype, grad_dtype, ws, st);
    return VSW_OK;
}
