// Interface of the tcgen05/TMEM/TMA bf16 / fp16 GEMM family (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace vsw {

struct TcLinearArgs {
    const void* x; const void* w; const void* bias; void* y;
    int M, N, K; int epi;
    void* aux_out; const void* res; const int32_t* rowmap; const float* rowscale;
    int rows_per_batch, dst_rows_per_batch;
    const void* gelu_pre;   // dgrad-style epilogue multiplier (used by tc_dgrad)
    int dtype;              // VSW_BF16 or VSW_F16
};

struct TcDgradArgs {
    const void* dy; const void* w; void* dx;
    int M, N, K;
    const int32_t* a_rowmap; const float* a_rowscale; int rows_per_batch, src_rows_per_batch;
    void* a_out; const void* gelu_pre;
    bool pre_is_grad;   // gelu_pre already holds gelu'(pre-activation): the epilogue is a plain multiply
    int dtype;          // VSW_BF16 or VSW_F16
};

// All return VSW_ERR_UNSUPPORTED (without launching anything) for shapes outside the tiling.
int tc_linear(const TcLinearArgs& a, cudaStream_t st);
int tc_dgrad(const TcDgradArgs& a, cudaStream_t st);
size_t tc_wgrad_workspace(int M, int N, int K);
int tc_wgrad(const void* dy, const void* x, void* dw, void* db, int M, int N, int K, int dtype, int grad_dtype, void* ws,
             size_t ws_bytes, cudaStream_t st);

}  // namespace vsw
