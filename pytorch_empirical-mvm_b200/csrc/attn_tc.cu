// tcgen05 / TMEM fused window attention for sm_100a (bf16, head_dim 32, N <= 448 tokens per window).
// Reference op: WindowAttention3D.forward, visbackbone/video_swin.py:149-169.
//
// One persistent CTA per SM walks (window, head) work items.  Per item the whole Q, K, V of the head
// (N x 32 bf16 each) is TMA-loaded into 64-byte-swizzled shared memory (double-buffered across items,
// 3-D tensor map so rows >= N read as zero), then for each 128-query tile:
//   S = Q K^T        tcgen05.mma (SS), fp32 accumulator 128 x Npad in TMEM (whole key range: no online rescale)
//   softmax          8 warps, one thread per (row, column half): pass 1 row max over TMEM, pass 2
//                    exp2(S*scale*log2e + bias*log2e [+ mask] - max) with the relative-position bias looked up
//                    from a shared-memory copy of the table column (index = rowcode[i] + colcode[j]) and the
//                    shift mask derived from uint8 region ids -- nothing N x N ever exists in HBM;
//                    P is written back to TMEM as packed bf16, aliasing S
//   O = P V          tcgen05.mma (TS: A = P from TMEM, B = V MN-major from smem), 128 x 32 fp32 in TMEM
//   epilogue         O / rowsum -> bf16 -> global, log-sum-exp -> global (for the recompute backward)
#include "attn.cuh"
#include "tc_common.cuh"

namespace vsw {
namespace {

constexpr int HD = 32;
constexpr int QT = 128;                       // query rows per tile
constexpr int MAXROWS = 512;                  // smem rows per operand
constexpr int OPER_BYTES = MAXROWS * HD * 2;  // 32 KB
constexpr int STAGE_BYTES = 3 * OPER_BYTES;   // Q, K, V
constexpr int BOX_BYTES = QT * HD * 2;        // 8 KB per TMA box
constexpr int NTHREADS = 384;                 // warp 0 TMA, 1 MMA, 2 aux, 3 idle, 4..11 softmax
constexpr int S_COL = 0, O_COL = 480, TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct FwdParams {
    const __nv_bfloat16* table; const int32_t* rowcode; const int32_t* colcode; const uint8_t* region;
    __nv_bfloat16* out; float* lse;
    int B_, nW, N, nH, L, Lpad;
    float scale_log2;
    int Npad, nq, split;
};

struct Smem {
    uint8_t* stage[2];
    float* tab[2];
    uint8_t* reg[2];
    int* rc; int* cc;
    float* xmax; float* xsum;   // [2][128]
    float* maxbias; int* masked;  // [2]
    uint64_t* qkv_full; uint64_t* qkv_empty; uint64_t* aux_full; uint64_t* aux_empty;  // [2] each
    uint64_t* s_full; uint64_t* p_full; uint64_t* o_full;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* base, int Lpad) {
    Smem s;
    s.stage[0] = base; s.stage[1] = base + STAGE_BYTES;
    uint8_t* p = base + 2 * STAGE_BYTES;
    s.tab[0] = (float*)p; p += (size_t)Lpad * 4;
    s.tab[1] = (float*)p; p += (size_t)Lpad * 4;
    s.rc = (int*)p; p += MAXROWS * 4;
    s.cc = (int*)p; p += MAXROWS * 4;
    s.xmax = (float*)p; p += 2 * QT * 4;
    s.xsum = (float*)p; p += 2 * QT * 4;
    s.maxbias = (float*)p; p += 8;
    s.masked = (int*)p; p += 8;
    s.qkv_full = (uint64_t*)p; p += 16;
    s.qkv_empty = (uint64_t*)p; p += 16;
    s.aux_full = (uint64_t*)p; p += 16;
    s.aux_empty = (uint64_t*)p; p += 16;
    s.s_full = (uint64_t*)p; p += 8;
    s.p_full = (uint64_t*)p; p += 8;
    s.o_full = (uint64_t*)p; p += 8;
    s.tmem_slot = (uint32_t*)p; p += 8;
    s.reg[0] = p; p += MAXROWS;
    s.reg[1] = p; p += MAXROWS;
    return s;
}
size_t fwd_smem_bytes(int Lpad) {
    return 1024 + 2 * (size_t)STAGE_BYTES + 2 * (size_t)Lpad * 4 + 2 * MAXROWS * 4 + 4 * QT * 4 + 16 + 4 * 16 + 3 * 8 + 8 +
           2 * MAXROWS + 64;
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const Smem s = carve(base, p.Lpad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.nH * HD;
    const int items = p.B_ * p.nH;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.qkv_full[i], 1); tc::mbar_init(&s.qkv_empty[i], 1);
            tc::mbar_init(&s.aux_full[i], 1); tc::mbar_init(&s.aux_empty[i], 8);
        }
        tc::mbar_init(s.s_full, 1); tc::mbar_init(s.p_full, 8); tc::mbar_init(s.o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(s.tmem_slot, TMEM_COLS);
    for (int n = threadIdx.x; n < p.N; n += NTHREADS) { s.rc[n] = p.rowcode[n]; s.cc[n] = p.colcode[n]; }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s.tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                const int b_ = w / p.nH, h = w - b_ * p.nH;
                tc::mbar_wait(&s.qkv_empty[st], ph ^ 1);
                tc::mbar_expect_tx(&s.qkv_full[st], 3 * p.nq * BOX_BYTES);
                for (int which = 0; which < 3; ++which)
                    for (int t = 0; t < p.nq; ++t)
                        tc::tma_load_3d(&tmQKV, &s.qkv_full[st], s.stage[st] + which * OPER_BYTES + t * BOX_BYTES,
                                        which * C + h * HD, t * QT, b_);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int it = 0; uint32_t pph = 0;
            const uint32_t idesc_pv = tc::idesc_bf16(QT, HD, 0, 1);
            const int n0len = p.Npad < 256 ? p.Npad : 256, n1len = p.Npad - n0len;
            for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                tc::mbar_wait(&s.qkv_full[st], ph);
                tc::tc_fence_after();
                const uint32_t qa = tc::smem_u32(s.stage[st]), ka = qa + OPER_BYTES, va = ka + OPER_BYTES;
                auto issue_qk = [&](int t) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc::umma_bf16(tmem + S_COL, tc::smem_desc_sw64(qa + t * BOX_BYTES + k * 32, 0, 512),
                                      tc::smem_desc_sw64(ka + k * 32, 0, 512), tc::idesc_bf16(QT, n0len, 0, 0), k);
                    if (n1len > 0) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::umma_bf16(tmem + S_COL + 256, tc::smem_desc_sw64(qa + t * BOX_BYTES + k * 32, 0, 512),
                                          tc::smem_desc_sw64(ka + 256 * 64 + k * 32, 0, 512),
                                          tc::idesc_bf16(QT, n1len, 0, 0), k);
                    }
                    tc::umma_commit(s.s_full);
                };
                issue_qk(0);
                for (int t = 0; t < p.nq; ++t) {
                    tc::mbar_wait(s.p_full, pph); pph ^= 1;
                    tc::tc_fence_after();
                    for (int kc = 0; kc < p.Npad / 16; ++kc) {
                        const int key0 = kc * 16;
                        const int pcol = key0 < p.split ? key0 / 2 : p.split + (key0 - p.split) / 2;
                        tc::umma_bf16_ts(tmem + O_COL, tmem + S_COL + pcol, tc::smem_desc_sw64(va + key0 * 64, 0, 512),
                                         idesc_pv, kc);
                    }
                    tc::umma_commit(s.o_full);
                    if (t + 1 < p.nq) issue_qk(t + 1);
                }
                tc::umma_commit(&s.qkv_empty[st]);
            }
        }
    } else if (warp == 2) {
        // ===================== aux loader: bias-table column (x log2e), region ids =====================
        int it = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int b_ = w / p.nH, h = w - b_ * p.nH, win = b_ % p.nW;
            tc::mbar_wait(&s.aux_empty[st], ph ^ 1);
            float mb = -INFINITY;
            for (int l = lane; l < p.L; l += 32) {
                const float v = __bfloat162float(p.table[(long long)l * p.nH + h]) * LOG2E;
                s.tab[st][l] = v;
                mb = fmaxf(mb, v);
            }
            mb = warp_max(mb);
            int diff = 0;
            if (p.region) {
                const uint8_t* rg = p.region + (long long)win * p.N;
                const uint8_t r0 = rg[0];
                for (int n = lane; n < p.N; n += 32) {
                    const uint8_t r = rg[n];
                    s.reg[st][n] = r;
                    diff |= (r != r0);
                }
            }
            diff = __any_sync(0xffffffffu, diff);
            if (lane == 0) { s.maxbias[st] = mb; s.masked[st] = diff; }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_full[st]);
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue warps =====================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int cbeg = half == 0 ? 0 : p.split, cend = half == 0 ? p.split : p.Npad;
        const int pbase = half == 0 ? 0 : p.split;   // P (packed bf16) column base, aliases S
        int it = 0; uint32_t sph = 0, oph = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int b_ = w / p.nH, h = w - b_ * p.nH;
            tc::mbar_wait(&s.aux_full[st], ph);
            const bool masked = s.masked[st] != 0;
            const float mb = s.maxbias[st];
            const float* tab = s.tab[st];
            const uint8_t* reg = s.reg[st];
            for (int t = 0; t < p.nq; ++t) {
                const int i = t * QT + row;
                const bool valid = i < p.N;
                const int ic = valid ? i : p.N - 1;
                const int rci = s.rc[ic];
                const uint8_t regi = masked ? reg[ic] : 0;
                tc::mbar_wait(s.s_full, sph); sph ^= 1;
                tc::tc_fence_after();
                // ---- pass 1: row max of scale*S (+mask); the bias is bounded by its per-head maximum
                float mx = -INFINITY;
                for (int c = cbeg; c < cend; c += 16) {
                    uint32_t r[16];
                    tc::tmem_ld_32x16(tmem + lane_base + S_COL + c, r);
                    tc::tmem_ld_wait();
                    const bool tail = c + 16 > p.N;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float v = __uint_as_float(r[e]) * p.scale_log2;
                        if (masked) v += (reg[min(c + e, p.N - 1)] != regi) ? -100.0f * LOG2E : 0.0f;
                        if (!tail || c + e < p.N) mx = fmaxf(mx, v);
                    }
                }
                s.xmax[half * QT + row] = mx;
                tc::named_bar_sync(1 + q, 64);
                mx = fmaxf(mx, s.xmax[(half ^ 1) * QT + row]) + mb;
                // ---- pass 2: p = exp2(s - max) -> packed bf16 into TMEM (aliasing S), row sum in fp32
                float sum = 0.f;
                for (int c = cbeg; c < cend; c += 16) {
                    uint32_t r[16];
                    tc::tmem_ld_32x16(tmem + lane_base + S_COL + c, r);
                    tc::tmem_ld_wait();
                    const bool tail = c + 16 > p.N;
                    uint32_t pw[8];
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        float pv[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int j = c + e + u;
                            const int jc = tail ? min(j, p.N - 1) : j;
                            float v = fmaf(__uint_as_float(r[e + u]), p.scale_log2, tab[rci + s.cc[jc]]) - mx;
                            if (masked) v += (reg[jc] != regi) ? -100.0f * LOG2E : 0.0f;
                            float pe = tc::ex2_approx(v);
                            if (tail && j >= p.N) pe = 0.f;
                            pv[u] = pe;
                            sum += pe;
                        }
                        pw[e / 2] = tc::pack_bf16(pv[0], pv[1]);
                    }
                    tc::tmem_st_32x8(tmem + lane_base + S_COL + pbase + (c - cbeg) / 2, pw);
                }
                tc::tmem_st_wait();
                s.xsum[half * QT + row] = sum;
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(s.p_full);
                // ---- epilogue: O / l -> bf16 -> global; lse
                tc::mbar_wait(s.o_full, oph); oph ^= 1;
                tc::tc_fence_after();
                const float l = s.xsum[row] + s.xsum[QT + row];
                const float inv = __fdividef(1.0f, l);
                uint32_t o[16];
                tc::tmem_ld_32x16(tmem + lane_base + O_COL + half * 16, o);
                tc::tmem_ld_wait();
                if (valid) {
                    __nv_bfloat16* dst = p.out + ((long long)b_ * p.N + i) * C + h * HD + half * 16;
                    uint4 u0, u1;
                    u0.x = tc::pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                    u0.y = tc::pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                    u0.z = tc::pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                    u0.w = tc::pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                    u1.x = tc::pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                    u1.y = tc::pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                    u1.z = tc::pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                    u1.w = tc::pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                    reinterpret_cast<uint4*>(dst)[0] = u0;
                    reinterpret_cast<uint4*>(dst)[1] = u1;
                    if (half == 0) p.lse[((long long)b_ * p.nH + h) * p.N + i] = (mx + __log2f(l)) * LN2;
                }
                tc::tc_fence_before();  // O reads complete before the next tile's p_full arrive lets PV overwrite O
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_empty[st]);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int tc_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                const uint8_t* region, void* out, float* lse, int B_, int nW, int N, int nH, int hd, int L,
                float scale, cudaStream_t st) {
    const int Lpad = (L + 3) / 4 * 4;
    const size_t smem = fwd_smem_bytes(Lpad);
    if (hd != HD || N > 448 || N < 1 || smem > 227 * 1024 || !aligned16(qkv) || !aligned16(out)) {
        set_error("tcgen05 window attention: needs head_dim 32, N <= 448 and a bias table that fits shared memory "
                  "(hd=%d N=%d L=%d)", hd, N, L);
        return VSW_ERR_UNSUPPORTED;
    }
    const int C = nH * HD;
    CUtensorMap tm;
    if (!make_tmap_3d_bf16(&tm, qkv, B_, N, 3 * C, 3 * C, (uint64_t)N * 3 * C, QT, HD, 64)) return VSW_ERR_CUDA;
    FwdParams p{};
    p.table = (const __nv_bfloat16*)table; p.rowcode = rowcode; p.colcode = colcode; p.region = region;
    p.out = (__nv_bfloat16*)out; p.lse = lse;
    p.B_ = B_; p.nW = nW; p.N = N; p.nH = nH; p.L = L; p.Lpad = Lpad;
    p.scale_log2 = scale * LOG2E;
    p.Npad = (N + 15) / 16 * 16;
    p.nq = (N + QT - 1) / QT;
    p.split = (p.Npad / 64) * 32;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attn fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
        configured = true;
    }
    const int items = B_ * nH;
    const int grid = items < kNumSMs ? items : kNumSMs;
    attn_fwd_tc_kernel<<<grid, NTHREADS, smem, st>>>(tm, p);
    return check_launch("attn_fwd_tc");
}

}  // namespace vsw

#include "attn_tc_bwd.inl"
