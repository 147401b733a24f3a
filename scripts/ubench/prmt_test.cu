// one-off check of the PRMT sign-replicate mask arithmetic used by the attention kernels
#include <cstdio>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ uint32_t ne_msb4(uint32_t a, uint32_t b) {
    const uint32_t d = a ^ b;
    return d | ((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu);
}
__global__ void k(const uint32_t* in, float* out, uint32_t* outw, float xin) {
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float MASKC = 100.0f * LOG2E / 1.7014118346046923e38f;
    const uint32_t rg = in[0], ri = in[1];
    const uint32_t nq = ne_msb4(rg, ri);
    outw[0] = nq;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint32_t m;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(nq), "r"(0u), "r"(0x8444u | ((uint32_t)(e & 3) << 12)));
        outw[1 + e] = m;
        out[e] = fmaf(__uint_as_float(m), MASKC, xin);
    }
    out[4] = MASKC;
}
int main() {
    uint32_t h[2] = {0x03020100u, 0x02020202u}, *d; float* o; uint32_t* w;
    cudaMalloc(&d, 8); cudaMalloc(&o, 64); cudaMalloc(&w, 64);
    cudaMemcpy(d, h, 8, cudaMemcpyHostToDevice);
    for (float x : {1.5f, -INFINITY}) {
        k<<<1, 1>>>(d, o, w, x);
        float ho[5]; uint32_t hw[5];
        cudaMemcpy(ho, o, 20, cudaMemcpyDeviceToHost); cudaMemcpy(hw, w, 20, cudaMemcpyDeviceToHost);
        printf("x=%g nq=%08x m=%08x %08x %08x %08x out=%g %g %g %g MASKC=%g\n", x, hw[0], hw[1], hw[2], hw[3], hw[4], ho[0], ho[1], ho[2], ho[3], ho[4]);
    }
    return 0;
}
