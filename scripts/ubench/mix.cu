// does MUFU.EX2 overlap with FFMA issue on the same SMSP?  (8 MUFU + NF FFMA per iteration, independent chains)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int NM, int NF>
__global__ void k_mix(float* out, int iters) {
    float a[8], b[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < NF / 32 + (NF % 32 != 0); ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i) if (r * 32 + i < NF) b[i] = fmaf(b[i], 1.0001f, 0.5f);
#pragma unroll
        for (int i = 0; i < NM; ++i) a[i] = ex2(a[i]);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    for (int i = 0; i < 32; ++i) s += b[i];
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0);
    if (s == 123.456f) out[0] = s;
}
template <int NM, int NF> void run(float* d, int nt) {
    int iters = 2048;
    k_mix<NM, NF><<<148, nt>>>(d, iters);
    cudaDeviceSynchronize();
    float h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    printf("MUFU %d + FFMA %3d per iter, %4d threads: %.1f cycles per iteration per SMSP-resident warp set (%d warps/SMSP) -> per warp-iteration %.1f\n",
           NM, NF, nt, h / iters, nt / 128, h / iters / (nt / 128.0));
}
int main() {
    float* d; cudaMalloc(&d, 4096);
    for (int nt : {128, 256, 512}) { run<8, 0>(d, nt); run<0, 64>(d, nt); run<8, 64>(d, nt); run<8, 128>(d, nt); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
