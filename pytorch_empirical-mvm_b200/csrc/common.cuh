// Shared device/host helpers for the vsw (Video-Swin B200) kernel library.
// sm_100a only.  No torch headers: the library is a plain C-ABI .so (see include/vsw.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vsw.h"

namespace vsw {

// ---------------------------------------------------------------------------------------------
// error plumbing: every entry point returns 0 or a negative vsw_status; message kept per thread
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> VSW_ERR_CUDA

#define VSW_REQUIRE(cond, code, ...)                                    \
    do {                                                                \
        if (!(cond)) { ::vsw::set_error(__VA_ARGS__); return (code); }  \
    } while (0)

constexpr int kNumSMs = 148;  // B200
constexpr int kMaxDevices = 64;   // per-device caches (function attributes, occupancy) are indexed by the current device
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }

// ---------------------------------------------------------------------------------------------
// storage types: float, __nv_bfloat16, __half; all arithmetic is fp32
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 16-byte vector of T
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T, int N>
struct alignas(sizeof(T) * N) Pack { T v[N]; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[Vec16<T>::N]) {
    constexpr int N = Vec16<T>::N;
    Pack<T, N> pk = *reinterpret_cast<const Pack<T, N>*>(p);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = to_f<T>(pk.v[i]);
}
template <typename T>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&in)[Vec16<T>::N]) {
    constexpr int N = Vec16<T>::N;
    Pack<T, N> pk;
#pragma unroll
    for (int i = 0; i < N; ++i) pk.v[i] = from_f<T>(in[i]);
    *reinterpret_cast<Pack<T, N>*>(p) = pk;
}

// N consecutive elements of T (N * sizeof(T) bytes, aligned to that): the narrow-type side of a mixed-width kernel whose vector
// length is set by the WIDER type (fp32 rows with a 16-bit output or gradient: 4 elements = 8 bytes)
template <typename T, int N>
__device__ __forceinline__ void load_n(const T* __restrict__ p, float (&out)[N]) {
    Pack<T, N> pk = *reinterpret_cast<const Pack<T, N>*>(p);
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = to_f<T>(pk.v[i]);
}
template <typename T, int N>
__device__ __forceinline__ void store_n(T* __restrict__ p, const float (&in)[N]) {
    Pack<T, N> pk;
#pragma unroll
    for (int i = 0; i < N; ++i) pk.v[i] = from_f<T>(in[i]);
    *reinterpret_cast<Pack<T, N>*>(p) = pk;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU and its derivative -- nn.GELU() default, reference video_swin.py:71
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// dtype dispatch for host launchers
#define VSW_DISPATCH_DTYPE(dtype, T, ...)                                         \
    switch (dtype) {                                                              \
        case VSW_F32:  { using T = float;          __VA_ARGS__; break; }          \
        case VSW_BF16: { using T = __nv_bfloat16;  __VA_ARGS__; break; }          \
        case VSW_F16:  { using T = __half;         __VA_ARGS__; break; }          \
        default: ::vsw::set_error("unsupported dtype %d", (int)(dtype)); return VSW_ERR_DTYPE; \
    }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace vsw
