// throughput of packed half-precision exponentials vs fp32 (B200, sm_100a): is ex2.approx.f16x2 one MUFU op for two elements?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    uint32_t a[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0x3c003800u + threadIdx.x + i; f[i] = threadIdx.x * 1e-3f + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            else if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
            else if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
            else if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(a[i]));
            else if (MODE == 4) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(f[i]), "f"(f[(i + 1) & 7])); f[i] += 1.0f; }
        }
    }
    long long t1 = clock64();
    float s = 0; uint32_t x = 0;
    for (int i = 0; i < 8; ++i) { s += f[i]; x ^= a[i]; }
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0);
    if (s == 123.456f || x == 77) out[0] = s;
}

int main() {
    float* d; cudaMalloc(&d, 4096);
    const char* names[] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2", "tanh.approx.f16x2", "cvt.rn.f16x2.f32 (+FADD)"};
    for (int mode = 0; mode < 5; ++mode)
        for (int nt : {256, 512, 1024}) {
            int iters = 4096;
            if (mode == 0) k<0><<<148, nt>>>(d, iters);
            if (mode == 1) k<1><<<148, nt>>>(d, iters);
            if (mode == 2) k<2><<<148, nt>>>(d, iters);
            if (mode == 3) k<3><<<148, nt>>>(d, iters);
            if (mode == 4) k<4><<<148, nt>>>(d, iters);
            cudaDeviceSynchronize();
            float h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
            double ops = (double)iters * 8 * nt;
            printf("%-26s threads %4d: %.2f instr-lanes/clk/SM  (%.1f cycles per warp-instr per SMSP)\n", names[mode], nt, ops / h,
                   h / (iters * 8.0 * (nt / 32) / 4.0));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
