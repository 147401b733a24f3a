// micro-benchmark: shared-memory fp32 atomic add throughput with warp-distinct addresses
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_atoms(float* out, int iters, int mode) {
    extern __shared__ float hist[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v = threadIdx.x * 1e-3f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int idx = (lane + it * 7 + warp * 131) & 4095;   // distinct within a warp
        if (mode == 0) atomicAdd(&hist[idx], v);
        else if (mode == 1) { asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&hist[idx])), "f"(v) : "memory"); }
        else { hist[idx] += v; }   // racy plain RMW (reference cost)
        v += 1e-6f;
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = (float)(t1 - t0); out[blockIdx.x * 2 + 1] = hist[5]; }
}
int main() {
    float* d; cudaMalloc(&d, 1024 * 8);
    for (int mode = 0; mode < 3; ++mode)
        for (int nt : {128, 256, 512}) {
            int iters = 2048;
            k_atoms<<<148, nt, 16384>>>(d, iters, mode);
            cudaDeviceSynchronize();
            float h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("mode %d threads %d: %.1f cycles per warp-instruction-per-SM (%.0f cyc for %d iters x %d warps)\n", mode, nt,
                   h[0] / (iters * (nt / 32.0)), h[0], iters, nt / 32);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
