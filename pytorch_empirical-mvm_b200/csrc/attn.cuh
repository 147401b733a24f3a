// Internal interface of the window-attention kernel families.
#pragma once
#include "common.cuh"

namespace vsw {

// CUDA-core family (attn_simt.cu)
int simt_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                  const uint8_t* region, const void* dmask, void* out, float* lse, int B_, int nW, int N, int nH,
                  int hd, int L, float scale, int dtype, cudaStream_t st);
size_t simt_attn_bwd_workspace(int B_, int N, int nH, int hd, int L);
int simt_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                  const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, const void* dmask, void* dqkv,
                  float* dbias, int B_, int nW, int N, int nH, int hd, int L, float scale, int dtype, void* ws,
                  size_t ws_bytes, cudaStream_t st);

// tcgen05 family (attn_tc.cu): bf16, head_dim 32, region-id masks.  UNSUPPORTED otherwise.
int tc_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                const uint8_t* region, void* out, float* lse, int B_, int nW, int N, int nH, int hd, int L,
                float scale, int window_dims, cudaStream_t st);
size_t tc_attn_bwd_workspace(int B_, int N, int nH, int hd, int L);
int tc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, void* dqkv, float* dbias,
                int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims, void* ws, size_t ws_bytes,
                cudaStream_t st);

// tcgen05 family, second generation (attn_tc2.cu): bf16 / fp16, head_dim 32, window rows of <= 8 tokens, N <= 448 a whole
// number of rows, the configured window given as layout hint.  UNSUPPORTED otherwise.
int tc2_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                 const uint8_t* region, void* out, float* lse, int B_, int nW, int N, int nH, int hd, int L,
                 float scale, int window_dims, int dtype, cudaStream_t st);
size_t tc2_attn_bwd_workspace(int B_, int N, int nH, int hd, int L);
int tc2_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                 const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, void* dqkv, float* dbias,
                 int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims, int dtype, void* ws, size_t ws_bytes,
                 cudaStream_t st);

}  // namespace vsw
