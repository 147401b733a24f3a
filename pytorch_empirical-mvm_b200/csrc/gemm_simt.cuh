// CUDA-core (fp32 FMA) GEMM used (a) for VSW_F32, where tensor-core rounding would break the
// fp32 rel-1e-4 parity bar (SURVEY section 7 hard part 3), and (b) as the cross-check for the
// tcgen05 kernels.  One generic strided kernel covers forward, dgrad and wgrad.
#pragma once
#include "common.cuh"

namespace vsw {

enum SimtEpi { SE_BIAS = 0, SE_GELU = 1, SE_RESIDUAL = 2, SE_DGRAD = 3, SE_PARTIAL = 4 };

struct SimtGemmParams {
    const void* A; const void* B; void* C;
    int M, N, K;                 // C (M x N) = A (M x K) * B (N x K)^T, reduction over K
    long long sam, sak, sbn, sbk;  // element strides of A(m,k), B(n,k)
    long long ldc;
    // optional gather / scale of A rows (dgrad of the scatter+residual epilogue)
    const int32_t* a_rowmap; const float* a_rowscale; int rows_per_batch; int src_rows_per_batch; void* a_out;
    // epilogue
    int epi; const void* bias; void* aux_out; const void* res; const int32_t* rowmap; const float* rowscale;
    int dst_rows_per_batch; const void* gelu_pre;
    int aux_is_grad;   // SE_GELU: aux_out receives gelu'(pre); SE_DGRAD: gelu_pre already holds gelu'(pre)
    // split-K (wgrad)
    int ksplit; int k_per_split; float* partial;
};

int launch_simt_gemm(const SimtGemmParams& p, bool a_kcontig, bool b_kcontig, int dtype, cudaStream_t st);

// sum of `splits` fp32 partial matrices -> out (grad dtype)
int launch_partial_reduce(const float* partial, int splits, long long elems, void* out, int out_dtype, cudaStream_t st);

// db[n] = sum_m dy[m,n]; uses ws (>= vsw_colsum_ws_bytes(N)) for fixed-order partials
size_t colsum_ws_bytes(int N);
int launch_colsum(const void* dy, int M, int N, void* db, int dtype, int out_dtype, void* ws, cudaStream_t st);

}  // namespace vsw
