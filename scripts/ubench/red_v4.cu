// Feasibility of the next attention-backward design: accumulate dS of every window into a per-CTA
// N x N fp32 matrix that lives in L2, with fire-and-forget vector reductions, instead of shared-memory histograms.
// Measures, for 148 CTAs x 256 threads each owning a private (N x ldg) fp32 matrix (N = 392, ldg = 400 -> 627 KB per CTA, 93 MB
// in total -- inside the 126 MB L2), the rate at which 128 x 128 fp32 tiles can be added into it:
//   MODE 0  red.global.add.v4.f32, thread = row   (the TMEM load layout: a warp instruction touches 32 rows x 16 B)
//   MODE 1  red.global.add.v4.f32, coalesced      (a warp instruction covers 512 contiguous bytes of one row)
//   MODE 2  st.global.v4.f32,      thread = row   (no read-modify-write: the store-path bound)
//   MODE 3  red.global.add.f32 x 4, thread = row  (scalar reductions)
// The real kernel produces one 128 x 128 dS tile per ~3.8 k cycles per SM (64 KB -> 17 B/clk/SM); the design needs the
// reduction path to sustain that on all 148 SMs at once, ideally with cycles to spare.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_v4 red_v4.cu && ./red_v4
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_s(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void st_v4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int N = 392, LDG = 400, TILES = 3;   // 3 x 3 full 128-tiles (the 8-row tails are ignored here)

template <int MODE>
__global__ void __launch_bounds__(256) k(float* G, int windows, long long* cycles) {
    float* g = G + (size_t)blockIdx.x * N * LDG;
    const int t = threadIdx.x;
    const long long t0 = clock64();
    for (int w = 0; w < windows; ++w)
        for (int kb = 0; kb < TILES; ++kb)
            for (int qt = 0; qt < TILES; ++qt) {
                const float v = 1e-3f * (float)(w + kb + qt + 1);
                if (MODE == 1) {
                    // coalesced: warp = 4 rows per instruction group? no: lane owns one float4 of a 128-float row segment
                    const int lane = t & 31, warp = t >> 5;               // 8 warps x 16 rows each
                    for (int r = 0; r < 16; ++r) {
                        float* p = g + (size_t)(qt * 128 + warp * 16 + r) * LDG + kb * 128 + lane * 4;
                        red_v4(p, v, v, v, v);
                    }
                } else {
                    const int row = t & 127, half = t >> 7;               // thread = row, 64 columns each
                    float* p = g + (size_t)(qt * 128 + row) * LDG + kb * 128 + half * 64;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        if (MODE == 0) red_v4(p + c * 4, v, v, v, v);
                        if (MODE == 2) st_v4(p + c * 4, v, v, v, v);
                        if (MODE == 3) { red_s(p + c * 4, v); red_s(p + c * 4 + 1, v); red_s(p + c * 4 + 2, v); red_s(p + c * 4 + 3, v); }
                    }
                }
            }
    __threadfence();
    const long long t1 = clock64();
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    const int ctas = 148, windows = 64;
    const size_t bytes = (size_t)ctas * N * LDG * sizeof(float);
    float* G; long long* cyc;
    cudaMalloc(&G, bytes); cudaMalloc(&cyc, ctas * sizeof(long long));
    cudaMemset(G, 0, bytes);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("L2 %d MB, persisting max %d MB, window max %d MB, matrix %.1f MB\n", prop.l2CacheSize >> 20,
           prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20, bytes / 1e6);
    cudaStream_t st; cudaStreamCreate(&st);
    const char* names[] = {"red.v4 thread=row", "red.v4 coalesced", "st.v4 thread=row", "red.f32 x4 thread=row"};
    for (int persist = 0; persist < 2; ++persist) {
        if (persist) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize);
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = G;
            av.accessPolicyWindow.num_bytes = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        }
        for (int mode = 0; mode < 4; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {      // rep 0 warms the L2
                cudaEventRecord(e0, st);
                switch (mode) {
                    case 0: k<0><<<ctas, 256, 0, st>>>(G, windows, cyc); break;
                    case 1: k<1><<<ctas, 256, 0, st>>>(G, windows, cyc); break;
                    case 2: k<2><<<ctas, 256, 0, st>>>(G, windows, cyc); break;
                    case 3: k<3><<<ctas, 256, 0, st>>>(G, windows, cyc); break;
                }
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
            const double tile_bytes = 128.0 * 128 * 4, tiles = (double)windows * TILES * TILES;
            printf("%-24s persist=%d: %.3f ms, %.0f cycles per 128x128 tile per SM (%.1f B/clk/SM), %.2f TB/s aggregate  [%s]\n",
                   names[mode], persist, ms, mx / tiles, tile_bytes * tiles / mx, ctas * tiles * tile_bytes / ms / 1e9,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
