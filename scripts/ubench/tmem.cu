// micro-benchmarks of the TMEM -> register path and of softmax-shaped loop bodies built on it (B200, sm_100a)
//   mode 0: tcgen05.ld 32x32b.x16 only          mode 1: 32x32b.x32 only          mode 2: 32x32b.x64 only
//   mode 3: x32 ld + FFMA + ex2 + pack + tcgen05.st x16 (forward softmax body without bias)
//   mode 4: mode 3 + bias from shared memory as bf16, one LDS.128 per 8 elements (dense bias tile)
//   mode 5: two x32 lds (S and dP) + bias + ex2 + dS math + 2 packs + 2 STS.128 per 8 elements (backward body, no histogram)
//   mode 6: mode 5 + shared-memory histogram read-modify-write per element
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// the step of the second-generation backward: FLAGS bit 0 = histogram RMW, 1 = {lse, delta} loads, 2 = dS tile stores, 3 = TMEM stores
template <int FLAGS>
__device__ __forceinline__ void bwd_step(const uint32_t (&rs)[16], const uint32_t (&rd)[16], uint32_t ld_a, uint32_t tab0, uint32_t tab1,
                                         uint32_t hist0, uint32_t hist1, uint32_t tile_a, uint32_t trow, float scale_log2) {
    float4 st[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) st[v] = (FLAGS & 2) ? lds_f4(ld_a + v * 16) : make_float4(1.f, 0.1f, 2.f, 0.2f);
    uint4 bias[2];
    bias[0] = lds_u4(tab0); bias[1] = lds_u4(tab1);
    float4 hh[4];
    if (FLAGS & 1) { hh[0] = lds_f4(hist0); hh[1] = lds_f4(hist0 + 16); hh[2] = lds_f4(hist1); hh[3] = lds_f4(hist1 + 16); }
    float ds[16], pp[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int sl = e & 7, r = e >> 3;
        if (sl >= 7) { pp[e] = 0.f; ds[e] = 0.f; continue; }
        const float4 t = st[e >> 1];
        const float l2 = (e & 1) ? t.z : t.x, dl = (e & 1) ? t.w : t.y;
        const uint4 bb = bias[r];
        const uint32_t w = (sl >> 1) == 0 ? bb.x : (sl >> 1) == 1 ? bb.y : (sl >> 1) == 2 ? bb.z : bb.w;
        const float x = fmaf(__uint_as_float(rs[e]), scale_log2, (sl & 1) ? bf_hi(w) : bf_lo(w)) - l2;
        const float pe = ex2(x);
        pp[e] = pe;
        ds[e] = pe * (__uint_as_float(rd[e]) - dl);
    }
    if (FLAGS & 1) {
        hh[0].x += ds[0]; hh[0].y += ds[1]; hh[0].z += ds[2]; hh[0].w += ds[3];
        hh[1].x += ds[4]; hh[1].y += ds[5]; hh[1].z += ds[6]; hh[1].w += ds[7];
        hh[2].x += ds[8]; hh[2].y += ds[9]; hh[2].z += ds[10]; hh[2].w += ds[11];
        hh[3].x += ds[12]; hh[3].y += ds[13]; hh[3].z += ds[14]; hh[3].w += ds[15];
        sts_f4(hist0, hh[0]); sts_f4(hist0 + 16, hh[1]); sts_f4(hist1, hh[2]); sts_f4(hist1 + 16, hh[3]);
    }
    uint32_t pw[8], dw[8];
#pragma unroll
    for (int e = 0; e < 16; e += 2) { pw[e >> 1] = pack_bf16(pp[e], pp[e + 1]); dw[e >> 1] = pack_bf16(ds[e], ds[e + 1]); }
    if (FLAGS & 8) { tmem_st8(trow, pw); tmem_st8(trow + 64, dw); }
    if (FLAGS & 4) { sts_u4(tile_a, make_uint4(dw[0], dw[1], dw[2], dw[3])); sts_u4(tile_a + 16, make_uint4(dw[4], dw[5], dw[6], dw[7])); }
    if (!(FLAGS & 12)) { if (pw[0] == 0x12345678u && dw[3] == 0x9abcdef0u) sts_u4(tile_a, make_uint4(pw[0], dw[1], 0, 0)); }
}

template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 1) k_tmem(float* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + i;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int grp = warp >> 2, ngrp = blockDim.x >> 7;          // warps with the same quadrant split the 512 columns
    const int cols = 512 / ngrp / 32 * 32;                       // my column range (multiple of 32)
    const uint32_t trow = tmem + lane_base + grp * cols;
    // initialise my TMEM range with finite values
    {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.001f * (lane + i));
        for (int c = 0; c < cols; c += 16) tmem_st16(trow + c, z);
        st_wait();
    }
    const uint32_t bias_a = smem_u32(smem) + grp * 8192 + (threadIdx.x & 127) * 16;   // [chunk of 8 keys][128 rows][16 B]
    const uint32_t tile_a = smem_u32(smem) + 65536 + (warp * 32 + lane) * 16;
    const uint32_t hist_a = smem_u32(smem) + 131072 + (warp & 7) * 3072 + lane * 4;
    float acc = 0.f;
    const float sc = 0.25f, nl = -3.0f, delta = 0.125f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < cols; c += 32) {
            if (MODE == 0) {
                uint32_t r[16], r2[16];
                tmem_ld16(trow + c, r); tmem_ld16(trow + c + 16, r2);
                ld_wait();
                acc += __uint_as_float(r[0]) + __uint_as_float(r2[15]);
            } else if (MODE == 1 || MODE == 2) {
                uint32_t r[32];
                tmem_ld32(trow + c, r);
                ld_wait();
                acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
            } else if (MODE == 3 || MODE == 4) {
                uint32_t r[32], pw[16];
                tmem_ld32(trow + c, r);
                uint4 b[4];
                if (MODE == 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = lds_u4(bias_a + ((c / 8 + u) & 3) * 2048);
                }
                ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float b0 = nl, b1 = nl;
                    if (MODE == 4) {
                        const uint4 bb = b[e >> 3];
                        const uint32_t w = ((e >> 1) & 3) == 0 ? bb.x : ((e >> 1) & 3) == 1 ? bb.y : ((e >> 1) & 3) == 2 ? bb.z : bb.w;
                        b0 = bf_lo(w); b1 = bf_hi(w);
                    }
                    const float p0 = ex2(fmaf(__uint_as_float(r[e]), sc, b0));
                    const float p1 = ex2(fmaf(__uint_as_float(r[e + 1]), sc, b1));
                    pw[e / 2] = pack_bf16(p0, p1);
                }
                tmem_st16(trow + c / 2, pw);   // (aliases the S columns like the real kernel; values stay finite)
            } else if (MODE >= 16) {
                // second-generation backward step: S in [c, c+16), dP in [c+16, c+32)
                uint32_t rs[16], rd[16];
                tmem_ld16(trow + c, rs); tmem_ld16(trow + c + 16, rd);
                ld_wait();
                const int rho = (c >> 5) & 31;
                const uint32_t rowi = (uint32_t)((threadIdx.x * 7 + rho * 13) % 1360);
                bwd_step<MODE - 16>(rs, rd, smem_u32(smem) + rho * 128, smem_u32(smem) + 8192 + rowi * 16, smem_u32(smem) + 8192 + ((rowi + 1) % 1360) * 16,
                                    smem_u32(smem) + 32768 + (warp >> 2 & 1) * 45056 + rowi * 32, smem_u32(smem) + 32768 + (warp >> 2 & 1) * 45056 + ((rowi + 1) % 1360) * 32,
                                    smem_u32(smem) + 131072 + (threadIdx.x & 127) * 128 + ((c >> 5) & 3) * 32, trow + (c >> 1), sc);
            } else {
                // backward body: columns [c, c+16) are "S", [c+16, c+32) are "dP"
                uint32_t rs[16], rd[16];
                tmem_ld16(trow + c, rs); tmem_ld16(trow + c + 16, rd);
                uint4 b[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) b[u] = lds_u4(bias_a + ((c / 8 + u) & 3) * 2048);
                ld_wait();
                uint32_t pw[8], dw[8];
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const uint4 bb = b[e >> 3];
                    const uint32_t w = ((e >> 1) & 3) == 0 ? bb.x : ((e >> 1) & 3) == 1 ? bb.y : ((e >> 1) & 3) == 2 ? bb.z : bb.w;
                    const float p0 = ex2(fmaf(__uint_as_float(rs[e]), sc, bf_lo(w)) + nl);
                    const float p1 = ex2(fmaf(__uint_as_float(rs[e + 1]), sc, bf_hi(w)) + nl);
                    const float d0 = p0 * (__uint_as_float(rd[e]) - delta), d1 = p1 * (__uint_as_float(rd[e + 1]) - delta);
                    pw[e / 2] = pack_bf16(p0, p1);
                    dw[e / 2] = pack_bf16(d0, d1);
                    if (MODE == 6) {
                        const uint32_t h0 = hist_a + ((e * 37 + c) & 63) * 128, h1 = hist_a + (((e + 1) * 37 + c) & 63) * 128;
                        sts_f32(h0, lds_f32(h0) + d0);
                        sts_f32(h1, lds_f32(h1) + d1);
                    }
                }
                sts_u4(tile_a + ((c >> 4) & 1) * 16384, make_uint4(pw[0], pw[1], pw[2], pw[3]));
                sts_u4(tile_a + ((c >> 4) & 1) * 16384 + 512 * 16, make_uint4(pw[4], pw[5], pw[6], pw[7]));
                sts_u4(tile_a + 32768 + ((c >> 4) & 1) * 16384, make_uint4(dw[0], dw[1], dw[2], dw[3]));
                sts_u4(tile_a + 32768 + ((c >> 4) & 1) * 16384 + 512 * 16, make_uint4(dw[4], dw[5], dw[6], dw[7]));
            }
        }
        if (MODE >= 3 && MODE < 16) st_wait();
        if (MODE >= 16 && ((MODE - 16) & 8)) st_wait();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0);
    if (acc == 123.456f) out[1] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, int NT>
void run1(float* d, const char* name) {
    cudaFuncSetAttribute(k_tmem<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int nt = NT, iters = 512;
    const int ngrp = nt / 128, cols = 512 / ngrp / 32 * 32;
    k_tmem<MODE, NT><<<148, nt, 200 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s threads %d: %s\n", name, nt, cudaGetErrorString(e)); return; }
    float h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    // elements: MODE 0-4 every column is one element; MODE 5-6 two TMEM columns per element
    const double tcols = (double)iters * cols * nt;        // 32-bit TMEM cells read
    const double elems = MODE >= 5 ? tcols / 2 : tcols;
    printf("%-44s threads %4d: %7.2f TMEM B/clk/SM  %6.2f elements/clk/SM\n", name, nt, tcols * 4 / h, elems / h);
}
template <int MODE>
void run(float* d, const char* name) {
    run1<MODE, 128>(d, name); run1<MODE, 256>(d, name); run1<MODE, 384>(d, name); run1<MODE, 512>(d, name);
    if (MODE >= 16) return;
    run1<MODE, 640>(d, name); run1<MODE, 768>(d, name); run1<MODE, 1024>(d, name);
}

int main() {
    float* d; cudaMalloc(&d, 4096);
    run<0>(d, "tcgen05.ld x16 (2 per wait)");
    run<1>(d, "tcgen05.ld x32");
    run<3>(d, "fwd body: ld + ffma + ex2 + pack + st");
    run<4>(d, "fwd body + bf16 bias tile (LDS.128)");
    run<5>(d, "bwd body: 2 ld + bias + ex2 + dS + 4 STS.128");
    run<6>(d, "bwd body + smem histogram RMW");
    run<16 + 15>(d, "bwd2 step: full (hist + stats + dS tile + STTM)");
    run<16 + 14>(d, "bwd2 step: no histogram");
    run<16 + 13>(d, "bwd2 step: no {lse, delta} loads");
    run<16 + 11>(d, "bwd2 step: no dS tile stores");
    run<16 + 7>(d, "bwd2 step: no TMEM stores");
    run<16 + 0>(d, "bwd2 step: math only");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
