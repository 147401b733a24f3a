// micro-benchmark of tcgen05.mma issue / execution rates for the small shapes of the window-attention kernels (B200, sm_100a):
// one thread issues R MMAs back to back, commits to an mbarrier and waits; cycles per MMA.  Optionally 12 other warps hammer
// TMEM with tcgen05.ld (what the exp warps do) or run a MUFU loop, to see what contention costs.
//   shapes: SS  M=128 N=112 K=16  A,B K-major 64-byte swizzle          (S = Q K^T block)
//           TS  M=128 N=48  K=16  A from TMEM, B MN-major 64B swizzle  ([O | l] += P [V | 1])
//           TS  M=128 N=32  K=16
//           SS  M=128 N=128 K=16  K-major 128-byte swizzle              (a GEMM-like reference point)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
constexpr uint32_t idesc_bf16(int m, int n, int amn, int bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// MODE: 0 SS N=112 SW64 | 1 TS N=48 SW64 MN-major B | 2 TS N=32 | 3 SS N=128 SW128 | 4 mix per block: 2 x mode 0 + 7 x mode 1
// DEP : 1 = every MMA accumulates into the SAME columns (a dependent chain), 0 = rotates over 4 column ranges
// LOAD: 0 none | 1 twelve warps loop over tcgen05.ld of other TMEM columns | 2 twelve warps loop over MUFU.EX2
template <int MODE, int DEP, int LOAD>
__global__ void __launch_bounds__(512, 1) k_mma(float* out, int R) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t sa = smem_u32(smem);
    if (warp == 15) {
        uint32_t el = 0;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
        if (el) {   // elect.sync: the MMA operands stay warp-uniform (no R2UR.BROADCAST loop)
            int j7 = 0; uint32_t cbuf = 0, cbuf2 = 224;
            long long t0 = clock64();
            for (int r = 0; r < R; ++r) {
                const uint32_t dcol = DEP ? 0u : (uint32_t)(r & 3) * 112u;
                if (MODE == 0) umma_ss(tmem + dcol, sdesc(sa + (r & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (r & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r > 3);
                if (MODE == 1) umma_ts(tmem + 448, tmem + dcol + 8 * (r % 7), sdesc(sa + 32768 + (r % 7) * 1024, 65536 - (r % 7) * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 0);
                if (MODE == 2) umma_ts(tmem + 448, tmem + dcol + 8 * (r % 7), sdesc(sa + 32768 + (r % 7) * 1024, 0, 512, 4), idesc_bf16(128, 32, 0, 1), r > 0);
                if (MODE == 3) umma_ss(tmem + (DEP ? 0u : (uint32_t)(r & 1) * 128u), sdesc(sa + (r & 3) * 32, 0, 1024, 2), sdesc(sa + 16384 + (r & 3) * 32, 0, 1024, 2), idesc_bf16(128, 128, 0, 0), r > 1);
                if (MODE == 5) {
                    const int j = r % 18;
                    if (j < 4) umma_ss(tmem + (uint32_t)(((r / 18) * 2 + (j >> 1)) & 3) * 112u, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j & 1);
                    else umma_ts(tmem + 448, tmem + (uint32_t)(((r / 18) * 2 + 2 + (j - 4) / 7) & 3) * 112u + 8 * ((j - 4) % 7), sdesc(sa + 32768 + ((j - 4) % 7) * 1024, 65536 - ((j - 4) % 7) * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 4);
                }
                if (MODE == 6) {   // all-SS block: 2 x (N=112, B K-major) + 7 x (N=48, A = P tile K-major 128B-swizzled, B MN-major)
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 3) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ss(tmem + 448, sdesc(sa + 65536 + ((j - 2) >> 2) * 16384 + ((j - 2) & 3) * 32, 0, 1024, 2), sdesc(sa + 32768 + (j - 2) * 1024, 24576 - (j - 2) * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 2);
                }
                if (MODE == 7) {   // only the P.V part as SS
                    const int j = r % 7;
                    umma_ss(tmem + 448, sdesc(sa + 65536 + (j >> 2) * 16384 + (j & 3) * 32, 0, 1024, 2), sdesc(sa + 32768 + j * 1024, 24576 - j * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 0);
                }
                if (MODE == 8) {   // alternate N only (both operands K-major)
                    if (r & 1) umma_ss(tmem + 448, sdesc(sa, 0, 512, 4), sdesc(sa + 16384, 0, 512, 4), idesc_bf16(128, 48, 0, 0), r > 1);
                    else umma_ss(tmem + (uint32_t)((r >> 1) & 3) * 112u, sdesc(sa + 32, 0, 512, 4), sdesc(sa + 16384 + 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), 0);
                }
                if (MODE == 9) {   // alternate the B major-ness only (N = 48)
                    if (r & 1) umma_ss(tmem + 448, sdesc(sa, 0, 512, 4), sdesc(sa + 16384, 0, 512, 4), idesc_bf16(128, 48, 0, 0), r > 1);
                    else umma_ss(tmem + 448, sdesc(sa + 32, 0, 512, 4), sdesc(sa + 32768, 24576, 512, 4), idesc_bf16(128, 48, 0, 1), 1);
                }
                if (MODE == 10) {  // N = 48 K-major B only
                    umma_ss(tmem + 448, sdesc(sa + (r & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (r & 1) * 32, 0, 512, 4), idesc_bf16(128, 48, 0, 0), r > 0);
                }
                if (MODE == 11) {  // 2 : 7 mix, all N = 112 K-major (same instruction descriptor), only D differs
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 1) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ss(tmem + 336, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r > 2);
                }
                if (MODE == 12) {  // 2 : 7 mix, N = 112 K-major / N = 48 K-major
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 3) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ss(tmem + 448, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 48, 0, 0), r > 2);
                }
                if (MODE == 13) {  // N = 32 MN-major B from a 128B-swizzled V tile (K = 16 rows x 64 B... as 128-B rows of two heads)
                    umma_ss(tmem + 448, sdesc(sa + (r & 1) * 32, 0, 512, 4), sdesc(sa + 32768 + (r % 7) * 2048, 0, 1024, 2), idesc_bf16(128, 64, 0, 1), r > 0);
                }
                if (MODE == 14) {  // S-type pairs only: (acc=0, acc=1) on rotating D
                    umma_ss(tmem + (uint32_t)((r >> 1) & 3) * 112u, sdesc(sa + (r & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (r & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r & 1);
                }
                if (MODE == 15) {  // mode 11 with accumulate always on
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 1) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), 1);
                    else umma_ss(tmem + 336, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), 1);
                }
                if (MODE == 16) {  // mode 11 with the chain's D rotating over two ranges per block
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 1) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ss(tmem + 224 + (uint32_t)(j & 1) * 112u, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r > 3);
                }
                if (MODE == 17) {  // 2 : 7 mix where the 7 use DIFFERENT smem operand addresses each (like V slices), same D chain
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 1) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ss(tmem + 336, sdesc(sa + 32768 + (j - 2) * 4096, 0, 512, 4), sdesc(sa + 65536 + (j - 2) * 4096, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r > 2);
                }
                if (MODE == 18) {  // 1 : 8 (single S-type MMA per block)
                    const int j = r % 9;
                    if (j < 1) umma_ss(tmem + (uint32_t)((r / 9) & 1) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), 0);
                    else umma_ss(tmem + 336, sdesc(sa + (j & 1) * 32, 0, 512, 4), sdesc(sa + 16384 + (j & 1) * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), r > 2);
                }
                if (MODE == 20) {  // TS N=32, 7-step chains, counters instead of % (clean issue loop)
                    umma_ts(tmem + 448, tmem + cbuf + 8 * j7, sdesc(sa + 32768 + j7 * 1024, 0, 512, 4), idesc_bf16(128, 32, 0, 1), r > 0);
                    if (++j7 == 7) { j7 = 0; cbuf = (cbuf + 112) & 511; if (cbuf >= 448) cbuf = 0; }
                }
                if (MODE == 21) {  // TS N=48 (LBO jump), clean
                    umma_ts(tmem + 448, tmem + cbuf + 8 * j7, sdesc(sa + 32768 + j7 * 1024, 65536 - j7 * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 0);
                    if (++j7 == 7) { j7 = 0; cbuf = (cbuf + 112) & 511; if (cbuf >= 448) cbuf = 0; }
                }
                if (MODE == 22) {  // block mix, clean: 2 x SS N112 then 7 x TS N32
                    if (j7 < 2) umma_ss(tmem + cbuf, sdesc(sa + j7 * 32, 0, 512, 4), sdesc(sa + 16384 + j7 * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j7);
                    else umma_ts(tmem + 448, tmem + cbuf2 + 8 * (j7 - 2), sdesc(sa + 32768 + (j7 - 2) * 1024, 0, 512, 4), idesc_bf16(128, 32, 0, 1), r > 2);
                    if (++j7 == 9) { j7 = 0; cbuf2 = cbuf; cbuf += 112; if (cbuf >= 448) cbuf = 0; }
                }
                if (MODE == 4) {
                    const int j = r % 9;
                    if (j < 2) umma_ss(tmem + (uint32_t)((r / 9) & 3) * 112u, sdesc(sa + j * 32, 0, 512, 4), sdesc(sa + 16384 + j * 32, 0, 512, 4), idesc_bf16(128, 112, 0, 0), j);
                    else umma_ts(tmem + 448, tmem + (uint32_t)(((r / 9) + 1) & 3) * 112u + 8 * (j - 2), sdesc(sa + 32768 + (j - 2) * 1024, 65536 - (j - 2) * 1024, 512, 4), idesc_bf16(128, 48, 0, 1), r > 2);
                }
            }
            long long t1 = clock64();
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, 0)) {}
            long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = (float)(t1 - t0) / R; out[1] = (float)(t2 - t0) / R; }
            stop = 1;
        }
    } else if (LOAD == 1 && warp < 12) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float acc = 0.f;
        while (!stop) {
#pragma unroll 1
            for (int c = 0; c < 448; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem + lane_base + c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[3]);
            }
        }
        if (acc == 1.2345f) out[2] = acc;
    } else if (LOAD == 3 && warp < 12) {
        // the forward softmax body: tcgen05.ld x16 + 2 LDS.128 (bias) + ffma + ex2 + pack + tcgen05.st x8, on the S buffers
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t bias_a = sa + 49152 + (threadIdx.x & 127) * 16;
        while (!stop) {
#pragma unroll 1
            for (int c = 0; c < 448; c += 16) {
                uint32_t r[16], pw[8];
                tmem_ld16(tmem + lane_base + c, r);
                uint4 b0, b1;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b0.x), "=r"(b0.y), "=r"(b0.z), "=r"(b0.w) : "r"(bias_a + ((c >> 4) & 7) * 2048));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b1.x), "=r"(b1.y), "=r"(b1.z), "=r"(b1.w) : "r"(bias_a + ((c >> 4) & 7) * 2048 + 16384));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    float p0, p1;
                    float v0 = fmaf(__uint_as_float(r[e]), 0.25f, __uint_as_float(bw[e >> 1] << 16));
                    float v1 = fmaf(__uint_as_float(r[e + 1]), 0.25f, __uint_as_float(bw[e >> 1] & 0xFFFF0000u));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(v0));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(v1));
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pw[e >> 1]) : "f"(p1), "f"(p0));
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                             ::"r"(tmem + lane_base + 496), "r"(pw[0]), "r"(pw[1]), "r"(pw[2]), "r"(pw[3]), "r"(pw[4]), "r"(pw[5]), "r"(pw[6]), "r"(pw[7]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    } else if (LOAD == 2 && warp < 12) {
        float a = threadIdx.x * 1e-3f;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        }
        if (a == 1.2345f) out[2] = a;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE, int DEP, int LOAD>
void run(float* d, const char* name) {
    cudaFuncSetAttribute(k_mma<MODE, DEP, LOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k_mma<MODE, DEP, LOAD><<<148, 512, 100 * 1024>>>(d, 1800);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-62s dep %d load %d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", name, DEP, LOAD, h[0], h[1]);
}

int main() {
    float* d; cudaMalloc(&d, 4096);
    run<0, 1, 0>(d, "SS M128 N112 K16 K-major SW64"); run<0, 0, 0>(d, "SS M128 N112 K16 K-major SW64"); run<0, 0, 1>(d, "SS M128 N112 K16 K-major SW64"); run<0, 0, 2>(d, "SS M128 N112 K16 K-major SW64");
    run<1, 1, 0>(d, "TS M128 N48 K16 B MN-major SW64 (LBO jump)"); run<1, 1, 1>(d, "TS M128 N48 K16 B MN-major SW64 (LBO jump)"); run<1, 1, 2>(d, "TS M128 N48 K16 B MN-major SW64 (LBO jump)");
    run<2, 1, 0>(d, "TS M128 N32 K16 B MN-major SW64"); run<2, 1, 1>(d, "TS M128 N32 K16 B MN-major SW64");
    run<3, 1, 0>(d, "SS M128 N128 K16 K-major SW128"); run<3, 0, 0>(d, "SS M128 N128 K16 K-major SW128"); run<3, 0, 1>(d, "SS M128 N128 K16 K-major SW128");
    run<4, 0, 0>(d, "block mix: 2 x SS N112 + 7 x TS N48"); run<4, 0, 1>(d, "block mix: 2 x SS N112 + 7 x TS N48"); run<4, 0, 2>(d, "block mix: 2 x SS N112 + 7 x TS N48");
    run<4, 0, 3>(d, "block mix: 2 x SS N112 + 7 x TS N48"); run<0, 0, 3>(d, "SS M128 N112 K16 K-major SW64"); run<1, 1, 3>(d, "TS M128 N48 K16 B MN-major SW64 (LBO jump)");
    run<6, 0, 0>(d, "all-SS block: 2 x SS N112 + 7 x SS N48 (P from smem)"); run<6, 0, 3>(d, "all-SS block: 2 x SS N112 + 7 x SS N48 (P from smem)");
    run<7, 1, 0>(d, "SS M128 N48 K16 A K-major SW128, B MN-major SW64"); run<7, 1, 3>(d, "SS M128 N48 K16 A K-major SW128, B MN-major SW64");
    run<8, 0, 0>(d, "alternate SS N112 / SS N48 (all K-major)"); run<9, 0, 0>(d, "alternate SS N48 B K-major / B MN-major");
    run<10, 1, 0>(d, "SS M128 N48 K16 all K-major"); run<11, 0, 0>(d, "2:7 mix, all N112 K-major, D differs"); run<12, 0, 0>(d, "2:7 mix N112 / N48 all K-major");
    run<13, 1, 0>(d, "SS M128 N64 K16 B MN-major SW128");
    run<14, 0, 0>(d, "S-type pairs (acc 0, acc 1) rotating D"); run<15, 0, 0>(d, "2:7 mix same idesc, accumulate always 1");
    run<16, 0, 0>(d, "2:7 mix same idesc, chain D alternates"); run<17, 0, 0>(d, "2:7 mix same idesc, chain operands at 7 addresses"); run<18, 0, 0>(d, "1:8 mix same idesc");
    run<20, 1, 0>(d, "clean TS M128 N32 K16"); run<20, 1, 1>(d, "clean TS M128 N32 K16"); run<20, 1, 3>(d, "clean TS M128 N32 K16");
    run<21, 1, 0>(d, "clean TS M128 N48 K16 (LBO jump)"); run<21, 1, 3>(d, "clean TS M128 N48 K16 (LBO jump)");
    run<22, 0, 0>(d, "clean block mix 2 x SS N112 + 7 x TS N32"); run<22, 0, 1>(d, "clean block mix 2 x SS N112 + 7 x TS N32"); run<22, 0, 3>(d, "clean block mix 2 x SS N112 + 7 x TS N32");
    run<5, 0, 0>(d, "pair mix: 4 x SS N112 + 14 x TS N48"); run<5, 0, 3>(d, "pair mix: 4 x SS N112 + 14 x TS N48");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
