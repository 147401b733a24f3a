// tcgen05 / TMEM fused window attention for sm_100a, second generation (bf16 / fp16, head_dim 32, window rows of <= 8 tokens,
// N <= 448 tokens per window).  Reference op: WindowAttention3D.forward, visbackbone/video_swin.py:149-169.
//
// What bounds this op on B200 (scripts/ubench/tmem.cu, profiles/r02_ubench_tmem.txt): at head_dim 32 one score costs 128 tensor
// FLOP (1/64 clk of the tensor pipe) but one MUFU.EX2 (1/16 clk), so the softmax loop -- not the tensor core, not TMEM
// (tcgen05.ld sustains > 500 B/clk/SM) -- is the floor: 14.5 scores/clk/SM for a bare ld+ffma+ex2+pack+st body.  The kernel is
// therefore organised so that the exp warps never wait and execute nothing per score except that body:
//   * keys are laid out in shared memory PADDED to 8 slots per window row (a 4-D tensor map whose box is wider than the
//     tensor: the dummy slot arrives as zeros), so that the relative-position bias of the 8 keys (d_j, h_j, 0..7) seen from
//     query (d_i, h_i, w_i) is ONE aligned 16-byte vector  T[w_i][d_i-d_j][h_i-h_j][0..7]  of a per-w_i pre-shifted copy of the
//     head's table column (21.8 KB for an 8x7x7 window, built once per call by a tiny kernel, x log2e, bf16): one LDS.128 per
//     8 scores, no gather, no index arrays, nothing N x N anywhere;
//   * S = Q K^T is produced per (128-query tile, 14-row key block = 112 columns) into a RING of four TMEM buffers; three
//     groups of four warps (one warp per TMEM lane quadrant) take the blocks round-robin, write P back as packed 16-bit
//     aliasing S, and the MMA thread chains O += P V (A from TMEM), l += P 1 and the S of the block four ahead: the tensor
//     core always runs ahead of the exp warps;
//   * no max is subtracted when |score| and |bias| are provably <= 50 log2 units for the whole window (Cauchy-Schwarz bound);
//     otherwise (and always for fp16, whose P must stay <= 1) the tile's blocks are issued twice: a max pass, then exp(s - max);
//   * a short last query tile (392 = 3 x 128 + 8) is only processed by the lane quadrants that hold real rows.
#include <stdlib.h>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>
#include "attn.cuh"
#include "tc_common.cuh"

namespace vsw {
namespace {

constexpr int HD = 32;
constexpr int QT = 128;                 // query rows per tile
constexpr int SLOT = 8;                 // key slots per window row (ww <= 8 real + dummies)
constexpr int KRB = 14;                 // key rows per block
constexpr int BW = KRB * SLOT;          // 112 score columns per block
constexpr int NSB = 4;                  // S buffers in TMEM
constexpr int NGRP = 3;                 // softmax groups (4 warps each)
constexpr int MAXKR = 56;               // key rows per window (N <= 448 and KR * 8 <= 448)
constexpr int MAXCOLS = MAXKR * SLOT;   // 448
constexpr int MAXNQ = 4, MAXNKB = 4;
constexpr int Q_BYTES = MAXNQ * QT * HD * 2;        // 32 KB
constexpr int KV_BYTES = MAXCOLS * HD * 2;          // 28 KB
constexpr int KV_BOX_BYTES = KRB * SLOT * HD * 2;   // 7 KB per TMA box
constexpr int STAGE_BYTES = Q_BYTES + 2 * KV_BYTES; // 88 KB
constexpr int TAB_MAX_BYTES = 24 * 1024;
constexpr int FWD_THREADS = 128 + NGRP * 128;
constexpr int W_AUX0 = 4 * NGRP, W_AUX1 = W_AUX0 + 1, W_TMA = W_AUX0 + 2, W_MMA = W_AUX0 + 3;   // the arbiter favours high warp ids: the MMA issuer is the last warp
constexpr int O_COL = NSB * BW, TMEM_COLS = 512;   // O: two 32-column buffers behind the four S buffers (448 + 64 = all 512 columns)
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr float MASKV = -100.0f * LOG2E;   // the reference's additive -100 (video_swin.py:304-306), in log2 units
constexpr float MASKC = 100.0f * LOG2E / 1.7014118346046923e38f;   // times -2^127 (0xFF000000) = MASKV
constexpr float SAFE = 50.0f;              // |exponent| bound (log2 units) under which no max is subtracted
constexpr float F16_TOP = 8.0f;            // fp16 single pass: largest exponent after the uniform shift (P <= 256)

// ---- shared-memory carve-up (offsets from a 1024-byte aligned base) ------------------------------------------------------
constexpr int OFF_STAGE = 0;
constexpr int OFF_TAB = 2 * STAGE_BYTES;                 // 176 KB
constexpr int OFF_LPART = OFF_TAB + TAB_MAX_BYTES;     // [4][MAXNKB][128] fp32: row sums of P per (tile slot, key block)
constexpr int OFF_KOFF = OFF_LPART + 4 * MAXNKB * QT * 4;                // int[64]: byte offset of the table row of key row rho
constexpr int OFF_REGK = OFF_KOFF + 256;                 // [2][448] region id per padded key column
constexpr int OFF_REGQ = OFF_REGK + 2 * MAXCOLS;         // [2][512] region id per query
constexpr int OFF_XMAX = OFF_REGQ + 2 * 512;             // [4][MAXNKB][128] fp32: exact-path partial row maxima
constexpr int OFF_AROW = OFF_XMAX + 4 * MAXNKB * QT * 4;   // [MAXNQ][128] byte offset of each query's table row (key row 0)
constexpr int OFF_FLAGS = OFF_AROW + MAXNQ * QT * 4;   // per stage: masked, q2max, k2max, head's bias max / min
constexpr int OFF_BARS = OFF_FLAGS + 64;
constexpr int FWD_SMEM = OFF_BARS + 256 + 1024;          // + alignment slack

struct FwdParams {
    const uint16_t* tabg; const float* tabstat; const int* poison; const uint8_t* region;
    void* out; float* lse;
    int B_, nW, N, nH, wh, ww, KR, nq, nkb, tab_bytes, NHt, blk;   // NHt = 2 wh - 1 (table rows per depth offset); blk = rows per w_i block
    int wdc;                                                  // configured window depth
    float scale_log2;
    int force_exact;
    int pair;         // blocks consumed per turn of the MMA loop (1 or 2)
    long long* dbg;   // optional phase counters (VSW_ATTN_DEBUG=1; never set otherwise)
};

struct Bars {
    uint64_t *qkv_full, *qkv_empty, *aux_full, *aux_empty, *s_full, *p_full, *o_full, *o_empty, *max_done;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ Bars bars_of(uint8_t* base) {
    uint64_t* b = reinterpret_cast<uint64_t*>(base + OFF_BARS);
    Bars s;
    s.qkv_full = b; s.qkv_empty = b + 2; s.aux_full = b + 4; s.aux_empty = b + 6; s.s_full = b + 8; s.p_full = b + 12;
    s.o_full = b + 16; s.o_empty = b + 20; s.max_done = b + 22; s.tmem_slot = reinterpret_cast<uint32_t*>(b + 26);
    return s;
}
struct StageFlags { int masked; float q2max, k2max, bmax, bmin; int pad[3]; };   // 32 bytes

// squared L2 norm of one head row (64 B) staged in shared memory (the 64-byte swizzle only permutes its 16-byte units)
template <bool F16>
__device__ __forceinline__ float row_norm2(uint32_t row_addr) {
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const uint4 a = tc::lds_u4(row_addr + v * 16);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 x;
            if (F16) x = __half22float2(*reinterpret_cast<const __half2*>(&w[e])); else x = tc::unpack_bf16(w[e]);
            acc = fmaf(x.x, x.x, acc);
            acc = fmaf(x.y, x.y, acc);
        }
    }
    return acc;
}
template <bool F16> __device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    if (F16) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
    return tc::pack_bf16(lo, hi);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t word_of(const uint4& b, int k) { return k == 0 ? b.x : (k == 1 ? b.y : (k == 2 ? b.z : b.w)); }

// ---- one 16-column chunk (two key rows) of a block: scores -> exponents -----------------------------------------------
// r: 16 fp32 scores of this thread's query row; b0 / b1: bias vectors (8 x bf16, x log2e) of the two key rows; nq4: per-byte
// "region differs" flags of the 16 columns (MASKED only); rows: 1 if only the first key row exists.
// per byte: MSB set <=> the bytes of a and b differ (any byte values; four integer instructions, no per-byte compare)
__device__ __forceinline__ uint32_t ne_msb4(uint32_t a, uint32_t b) {
    const uint32_t d = a ^ b;
    return d | ((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu);
}
// 0xFF000000 (= -2^127 as a float) if the MSB of byte k of w is set, else 0: one PRMT in sign-replicate mode (selector nibble
// bit 3; __byte_perm masks that bit off, hence the inline PTX)
__device__ __forceinline__ uint32_t msb_to_top(uint32_t w, int k) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(0x8444u | ((uint32_t)k << 12)));
    return r;
}
template <int WW, bool MASKED, bool EXACT>
__device__ __forceinline__ void chunk_exponents(const uint32_t (&r)[16], const uint4& b0, const uint4& b1, const uint32_t (&nq4)[4],
                                                float scale_log2, float nm, float (&v)[16]) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int slot = e & 7;
        if (slot >= WW) { v[e] = -INFINITY; continue; }
        const uint32_t w = word_of(e < 8 ? b0 : b1, slot >> 1);
        float x = fmaf(__uint_as_float(r[e]), scale_log2, (slot & 1) ? bf_hi(w) : bf_lo(w));
        if (EXACT) x += nm;
        // byte MSB = region ids differ; replicated into the top byte it is -2^127 or +0.0 (one PRMT), times MASKC = the additive -100
        if (MASKED) x = fmaf(__uint_as_float(msb_to_top(nq4[e >> 2], e & 3)), MASKC, x);
        v[e] = x;
    }
}

struct BlockArgs {
    uint32_t sbuf;        // TMEM address (lane quadrant + buffer column) of the S block
    uint32_t arow;        // shared address of this thread's table row for key row 0
    uint32_t koff_a;      // shared address of koff[first key row of the block]
    uint32_t regk_a;      // shared address of regk[first padded column of the block]
    uint32_t regi4;       // this row's region id replicated into four bytes
    int nrows;            // key rows of the block that exist (1..14)
    float scale_log2, nm;
};

// exp pass over one block: P = exp2(S*scale*log2e + bias [+ mask] [- max]) -> packed 16-bit back into TMEM, aliasing S
template <int WW, bool MASKED, bool EXACT, bool F16>
__device__ __forceinline__ float block_exp(const BlockArgs& a) {
    const int nfull = a.nrows >> 1;   // chunks whose two key rows both exist (all of them unless the window has an odd row count)
    float ls0 = 0.f, ls1 = 0.f;       // row sum of the (unrounded) exponentials of this block, two add chains
    auto body = [&](const uint32_t (&r)[16], int k, bool two) {
        const uint2 ko = lds_u2(a.koff_a + k * 8);
        const uint4 b0 = tc::lds_u4(a.arow - ko.x);
        const uint4 b1 = two ? tc::lds_u4(a.arow - ko.y) : make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
        uint32_t nq4[4] = {0, 0, 0, 0};
        if (MASKED) {
            const uint4 rg = tc::lds_u4(a.regk_a + k * 16);
            nq4[0] = ne_msb4(rg.x, a.regi4); nq4[1] = ne_msb4(rg.y, a.regi4);
            nq4[2] = ne_msb4(rg.z, a.regi4); nq4[3] = ne_msb4(rg.w, a.regi4);
        }
        float v[16];
        chunk_exponents<WW, MASKED, EXACT>(r, b0, b1, nq4, a.scale_log2, a.nm, v);
        uint32_t pw[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            const float p0 = ((e & 7) >= WW) ? 0.f : tc::ex2_approx(v[e]);
            const float p1 = (((e + 1) & 7) >= WW) ? 0.f : tc::ex2_approx(v[e + 1]);
            ls0 += p0; ls1 += p1;
            pw[e >> 1] = pack16<F16>(p0, p1);
        }
        tc::tmem_st_32x8(a.sbuf + 8 * k, pw);
    };
    uint32_t ra[16], rb[16];
    if (nfull > 0) { tc::tmem_ld_32x16(a.sbuf, ra); tc::tmem_ld_wait(); }
    for (int k = 0; k < nfull; k += 2) {
        const bool hb = k + 1 < nfull, ha = k + 2 < nfull;
        if (hb) tc::tmem_ld_32x16(a.sbuf + 16 * (k + 1), rb);
        body(ra, k, true);
        if (hb) {
            tc::tmem_ld_wait();
            if (ha) tc::tmem_ld_32x16(a.sbuf + 16 * (k + 2), ra);
            body(rb, k + 1, true);
            if (ha) tc::tmem_ld_wait();
        }
    }
    if (a.nrows & 1) {   // odd row count: the last chunk holds one key row; its second half is padding (P = 0)
        tc::tmem_ld_32x16(a.sbuf + 16 * nfull, ra);
        tc::tmem_ld_wait();
        body(ra, nfull, false);
    }
    return ls0 + ls1;
}

// max pass over one block (exact path): row maximum of the exponents of this block
template <int WW, bool MASKED>
__device__ __forceinline__ float block_max(const BlockArgs& a) {
    const int nch = (a.nrows + 1) >> 1;
    float mx = -INFINITY;
#pragma unroll 1
    for (int k = 0; k < nch; ++k) {
        uint32_t r[16];
        tc::tmem_ld_32x16(a.sbuf + 16 * k, r);
        const uint2 ko = lds_u2(a.koff_a + k * 8);
        const uint4 b0 = tc::lds_u4(a.arow - ko.x);
        const bool two = 2 * k + 1 < a.nrows;
        const uint4 b1 = two ? tc::lds_u4(a.arow - ko.y) : make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
        uint32_t nq4[4] = {0, 0, 0, 0};
        if (MASKED) {
            const uint4 rg = tc::lds_u4(a.regk_a + k * 16);
            nq4[0] = ne_msb4(rg.x, a.regi4); nq4[1] = ne_msb4(rg.y, a.regi4);
            nq4[2] = ne_msb4(rg.z, a.regi4); nq4[3] = ne_msb4(rg.w, a.regi4);
        }
        tc::tmem_ld_wait();
        float v[16];
        chunk_exponents<WW, MASKED, false>(r, b0, b1, nq4, a.scale_log2, 0.f, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) mx = fmaxf(mx, v[e]);
    }
    return mx;
}

// bounded mbarrier wait with a site tag in the time-out message.  (Keep this free of out-of-line calls: one real function
// call on the MMA thread's path makes ptxas give up uniform registers for the whole issue loop -- every UTCHMMA then pays an
// ELECT / R2UR.BROADCAST sequence of ~100 cycles, which is what bounded the first version of this kernel.)
__device__ __forceinline__ void wait_dbg(uint64_t* bar, uint32_t parity, int tag, long long*) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (tc::mbar_try_wait(bar, parity)) return;
    printf("vsw attn2: mbarrier wait timed out (block %d thread %d, wait site %d, parity %u)\n", blockIdx.x, threadIdx.x, tag, parity);
    __trap();
}

// phase counters are compiled in only with -DVSW_ATTN2_PROF=1 (VSW_NVCC_EXTRA of build.py): they cost ~25 registers
#ifndef VSW_ATTN2_SETMAXNREG
#define VSW_ATTN2_SETMAXNREG 0
#endif
#ifndef VSW_ATTN2_PROF
#define VSW_ATTN2_PROF 0
#endif
#define DBGP (VSW_ATTN2_PROF ? p.dbg : (long long*)nullptr)

template <int WW, bool F16>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const Bars s = bars_of(base);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.nH * HD;
    // work items = (head, window) pairs in HEAD-MAJOR order, one contiguous range per CTA (the table stays resident)
    const int items = p.B_ * p.nH;
    const int w_begin = (int)blockIdx.x * (items / (int)gridDim.x) + min((int)blockIdx.x, items % (int)gridDim.x);
    const int w_end = w_begin + items / (int)gridDim.x + ((int)blockIdx.x < items % (int)gridDim.x ? 1 : 0);
    StageFlags* flags = reinterpret_cast<StageFlags*>(base + OFF_FLAGS);
    int* koff = reinterpret_cast<int*>(base + OFF_KOFF);

    if (*p.poison) {
        // the codes are not the dense (d, h, w) codes of the window the caller named: refuse loudly (NaN everywhere)
        const uint32_t nan2 = F16 ? 0x7E007E00u : 0x7FC07FC0u;
        for (int w = w_begin; w < w_end; ++w) {
            const int h = w / p.B_, b_ = w - h * p.B_;
            for (int n = threadIdx.x; n < p.N * 16; n += FWD_THREADS)
                reinterpret_cast<uint32_t*>((uint16_t*)p.out + ((long long)b_ * p.N + (n >> 4)) * C + h * HD)[n & 15] = nan2;
            for (int n = threadIdx.x; n < p.N; n += FWD_THREADS) p.lse[((long long)b_ * p.nH + h) * p.N + n] = __int_as_float(0x7FC00000);
        }
        return;
    }

    if (warp == W_TMA && lane == 0) {
        tc::prefetch_tmap(&tmQ);
        tc::prefetch_tmap(&tmKV);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.qkv_full[i], 1); tc::mbar_init(&s.qkv_empty[i], 1);
            tc::mbar_init(&s.aux_full[i], 2); tc::mbar_init(&s.aux_empty[i], 4 * NGRP);
        }
        for (int i = 0; i < NSB; ++i) { tc::mbar_init(&s.s_full[i], 1); tc::mbar_init(&s.p_full[i], 4); }
        for (int i = 0; i < 4; ++i) tc::mbar_init(&s.o_full[i], 1);
        tc::mbar_init(&s.o_empty[0], 4); tc::mbar_init(&s.o_empty[1], 4);
        for (int i = 0; i < 4; ++i) tc::mbar_init(&s.max_done[i], 4 * p.nkb);
        tc::fence_barrier_init();
    }
    if (warp == W_MMA) tc::tmem_alloc(s.tmem_slot, TMEM_COLS);
    for (int n = threadIdx.x; n < 64; n += FWD_THREADS) koff[n] = ((n / p.wh) * p.NHt + n % p.wh) * 16;
    for (int n = threadIdx.x; n < MAXNQ * QT; n += FWD_THREADS) {   // the divisions are done once, not once per tile
        const int ic = min(n, p.N - 1);
        const int wi = ic % p.ww, hi = (ic / p.ww) % p.wh, di = ic / (p.ww * p.wh);
        reinterpret_cast<uint32_t*>(base + OFF_AROW)[n] = (wi * p.blk + (di + p.wdc - 1) * p.NHt + hi + p.wh - 1) * 16;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s.tmem_slot;
    const uint32_t stage_a = tc::smem_u32(base + OFF_STAGE);
    auto rows_of = [&](int kb) { return min(KRB, p.KR - kb * KRB); };   // key rows of block kb
    // bf16: the exponents may be used as they are while |exponent| <= 2 SAFE (2^100 is no problem for bf16 / fp32).
    // fp16: P must stay inside half's range.  With bound >= every |exponent| of the item (Cauchy-Schwarz + bias range) and the
    // uniform shift c = bound - F16_TOP, P = 2^(x - c) <= 2^F16_TOP = 256, and the largest P of any row is >= 2^(F16_TOP - 2 bound)
    // >= 2^-8 while bound <= F16_TOP (normal halves, full precision); beyond that the two-pass path with the true row maximum.
    auto f16_bound = [&](int st) {
        const StageFlags f = flags[st];
        return sqrtf(f.q2max * f.k2max) * p.scale_log2 + fmaxf(fabsf(f.bmax), fabsf(f.bmin));
    };
    auto is_exact = [&](int st) {
        if (p.force_exact) return true;
        if (F16) return !(f16_bound(st) <= F16_TOP);
        const StageFlags f = flags[st];
        const bool bias_ok = f.bmax <= SAFE && f.bmin >= -SAFE;
        return !bias_ok || !(f.q2max * f.k2max * p.scale_log2 * p.scale_log2 <= SAFE * SAFE);
    };

    // Register re-partitioning (65536 = 3 x 128 x 152 + 128 x 56): the exp warps need ~150 registers for a software-pipelined
    // 16-column chunk; the four service warps (aux, TMA, MMA: one warpgroup) get by with 56.
    if (warp == W_TMA) {
        // ===================== TMA producer =====================
        if (VSW_ATTN2_SETMAXNREG) tc::setmaxnreg_dec<56>();
        if (tc::elect_one()) {
            int it = 0;
            for (int w = w_begin; w < w_end; ++w, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                const int h = w / p.B_, b_ = w - h * p.B_;
                wait_dbg(&s.qkv_empty[st], ph ^ 1, 1, DBGP);
                tc::mbar_expect_tx(&s.qkv_full[st], p.nq * QT * HD * 2 + 2 * p.nkb * KV_BOX_BYTES);
                uint8_t* dst = base + OFF_STAGE + st * STAGE_BYTES;
                for (int t = 0; t < p.nq; ++t) tc::tma_load_3d(&tmQ, &s.qkv_full[st], dst + t * QT * HD * 2, h * HD, t * QT, b_);
                for (int which = 1; which < 3; ++which)
                    for (int kb = 0; kb < p.nkb; ++kb)
                        tc::tma_load_4d(&tmKV, &s.qkv_full[st], dst + Q_BYTES + (which - 1) * KV_BYTES + kb * KV_BOX_BYTES,
                                        which * C + h * HD, 0, kb * KRB, b_);
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (VSW_ATTN2_SETMAXNREG) tc::setmaxnreg_dec<56>();
        // The S blocks of the whole item sequence are issued in order, four ahead of the block whose P the exp warps have
        // just finished; per finished exp block:  O (+)= P V,  l (+)= P 1,  then the S of the block four ahead.
        // One ELECTED thread issues every tcgen05.mma.  UTCHMMA takes its operands from uniform registers: with elect.sync
        // ptxas knows that exactly one thread is here and moves them over with a plain R2UR; behind `lane == 0` it cannot, and
        // wraps every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop of ~100 cycles -- that loop, not the tensor pipe, is what
        // bounded the first version of this kernel (and the round-1 kernels) at ~2 k cycles per 128 x 112 block.
        if (tc::elect_one()) {
            if (tmem != 0) __trap();   // all 512 columns are allocated, so the base is column 0 / lane 0; a literal keeps it uniform
            const bool lead = true;
            const uint32_t fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc_base = (1u << 4) | fmt | ((uint32_t)(QT >> 4) << 24);
            const uint32_t idesc_pv = idesc_base | (1u << 16) | ((uint32_t)(HD >> 3) << 17);   // B = V slice, MN-major
            const uint32_t desc_hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO 512 B, descriptor version 1, 64-byte swizzle
            // Two cursors walk the same block sequence (items -> tiles -> [max pass] exp pass -> key blocks): the ISSUE cursor
            // (S = Q K^T of a block into S buffer ci % NSB) runs up to NSB blocks ahead of the CONSUME cursor (P.V of a block).
            // (Plain straight-line code, no by-reference lambdas: cursor state that ends up in local memory is no longer
            // warp-uniform for ptxas.)
            int iw = w_begin, iit = 0, it_t = 0, it_kb = 0, it_pass = 1; bool it_open = false, it_exact = false;
            int ci = 0;   // blocks issued so far
            int cit = 0, c_t = 0, c_kb = 0, c_pass = 1; bool c_open = false, c_exact = false;
            int cc = 0;   // blocks consumed so far
            int otile = 0;   // exp tiles whose first P.V has been issued
            long long pa8 = 0, pa10 = 0, pa12 = 0, pa13 = 0, pa14 = 0, pa15 = 0, pa16 = 0;   // phase counters (debug runs), flushed once at the end
            const long long k_t0 = DBGP ? clock64() : 0;
            unsigned long long k_g0 = 0;
            if (DBGP) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_g0));
            for (;;) {
                // ---- issue: S blocks up to NSB ahead.  With blocks in flight, never WAIT for the next item's tiles / side
                // data (the exp warps of the item in flight may still need this warp -- their epilogue waits for its P.V --
                // and on a head change the aux warp in turn waits for them): probe, and come back after the next pair.
                const bool prof0 = DBGP && blockIdx.x == 0;
                long long i_a = 0;
                if (prof0) i_a = clock64();
                while (ci - cc < NSB && iw < w_end) {
                    const int st = iit & 1; const uint32_t ph = (iit >> 1) & 1;
                    if (!it_open) {
                        if (ci == cc) {
                            long long b_a = 0;
                            if (prof0) b_a = clock64();
                            wait_dbg(&s.qkv_full[st], ph, 2, DBGP);
                            if (prof0) { pa14 += clock64() - b_a; b_a = clock64(); }
                            wait_dbg(&s.aux_full[st], ph, 3, DBGP);
                            if (prof0) { pa15 += clock64() - b_a; pa16 += 1; }
                        }
                        else if (!(tc::mbar_test(&s.qkv_full[st], ph) && tc::mbar_test(&s.aux_full[st], ph))) break;
                        tc::tc_fence_after();
                        it_exact = is_exact(st);
                        it_open = true; it_t = 0; it_kb = 0; it_pass = it_exact ? 0 : 1;
                    }
                    const int nrows = min(KRB, p.KR - it_kb * KRB);
                    const uint32_t idesc_s = idesc_base | ((uint32_t)(((nrows + 1) >> 1) * 2) << 17);   // N = 16 * ceil(nrows / 2)
                    const uint32_t qa = stage_a + st * STAGE_BYTES + it_t * (QT * HD * 2), ka = stage_a + st * STAGE_BYTES + Q_BYTES + it_kb * KV_BOX_BYTES;
                    const int buf = ci & (NSB - 1);
                    if (lead) {
                        tc::umma_bf16(buf * BW, tc::smem_desc_sw64(qa, 0, 512), tc::smem_desc_sw64(ka, 0, 512), idesc_s, 0u);
                        tc::umma_bf16(buf * BW, tc::smem_desc_sw64(qa + 32, 0, 512), tc::smem_desc_sw64(ka + 32, 0, 512), idesc_s, 1u);
                        tc::umma_commit(&s.s_full[buf]);
                    }
                    if (++it_kb == p.nkb) {
                        it_kb = 0;
                        if (it_exact && it_pass == 0) it_pass = 1;
                        else {
                            it_pass = it_exact ? 0 : 1;
                            if (++it_t == p.nq) { ++iw; ++iit; it_open = false; }
                        }
                    }
                    ++ci;
                }
                if (prof0) pa13 += clock64() - i_a;
                if (cc == ci) break;   // nothing in flight and nothing left to issue
                // ---- consume a PAIR of blocks: the P.V MMAs of two blocks (A operand from TMEM) back to back, then (above)
                // the S MMAs of the two blocks four ahead (both operands from shared memory)
                const int npair = min(p.pair, ci - cc);
                const bool prof = DBGP && blockIdx.x == 0 && lead;
                long long m_a = 0, m_w = 0;
                if (prof) m_a = clock64();
                for (int j = 0; j < npair; ++j) {
                    const int c_st = cit & 1;
                    if (!c_open) {   // the issue cursor has been here: the item's side data is ready
                        c_exact = is_exact(c_st);
                        c_open = true; c_t = 0; c_kb = 0; c_pass = c_exact ? 0 : 1;
                    }
                    const int buf = cc & (NSB - 1);
                    long long m_b = 0;
                    if (prof) m_b = clock64();
                    wait_dbg(&s.p_full[buf], (cc / NSB) & 1, 4, DBGP);
                    tc::tc_fence_after();
                    if (c_pass == 1) {
                        if (c_kb == 0) {
                            // O is double-buffered: the tile two back must have been read out of this buffer
                            if (otile > 1) { wait_dbg(&s.o_empty[otile & 1], ((otile >> 1) - 1) & 1, 5, DBGP); tc::tc_fence_after(); }
                            ++otile;
                        }
                        if (prof) m_w += clock64() - m_b;
                        const uint32_t va = stage_a + c_st * STAGE_BYTES + Q_BYTES + KV_BYTES + c_kb * KV_BOX_BYTES;
                        const int nch = (min(KRB, p.KR - c_kb * KRB) + 1) >> 1;
                        const uint32_t pcol = buf * BW, ocol = O_COL + ((otile - 1) & 1) * HD;
                        const uint32_t acc = c_kb > 0 ? 1u : 0u;
#pragma unroll
                        for (int k = 0; k < KRB / 2; ++k)   // O (+)= P V, 16 keys per MMA
                            if (k < nch && lead)
                                tc::umma_bf16_ts(ocol, pcol + 8 * k, tc::smem_desc_sw64(va + k * 1024, 0, 512), idesc_pv, k ? 1u : acc);
                        if (c_kb == p.nkb - 1 && lead) tc::umma_commit(&s.o_full[(otile - 1) & 3]);
                    } else if (prof) m_w += clock64() - m_b;
                    ++cc;
                    if (++c_kb == p.nkb) {
                        c_kb = 0;
                        if (c_exact && c_pass == 0) c_pass = 1;
                        else {
                            c_pass = c_exact ? 0 : 1;
                            if (++c_t == p.nq) {   // last block of the item: its Q / K / V stage may be refilled once these MMAs retire
                                if (lead) tc::umma_commit(&s.qkv_empty[c_st]);
                                ++cit; c_open = false;
                            }
                        }
                    }
                }
                if (prof) { pa8 += m_w; pa10 += clock64() - m_a - m_w; pa12 += npair; }
            }
            if (DBGP && blockIdx.x == 0) { unsigned long long k_g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_g1)); DBGP[20] += clock64() - k_t0; DBGP[21] += (long long)(k_g1 - k_g0); }
            if (DBGP && blockIdx.x == 0) { DBGP[8] += pa8; DBGP[10] += pa10; DBGP[12] += pa12; DBGP[13] += pa13; DBGP[14] += pa14; DBGP[15] += pa15; DBGP[16] += pa16; }
        }
    } else if (warp == W_AUX0 || warp == W_AUX1) {
        // ===================== aux warps =====================
        if (VSW_ATTN2_SETMAXNREG) tc::setmaxnreg_dec<56>();
        // first aux warp: the head's shifted bias table (only when the head changes) and max_j |k_j|^2;
        // second aux warp: region ids of the window (padded key order and query order) and max_i |q_i|^2.
        int it = 0, tab_head = -1;
        for (int w = w_begin; w < w_end; ++w, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int h = w / p.B_, b_ = w - h * p.B_, win = b_ % p.nW;
            wait_dbg(&s.aux_empty[st], ph ^ 1, 6, DBGP);
            if (warp == W_AUX0) {
                if (tab_head != h) {
                    // every exp warp must have left the previous item (it still reads the old head's table)
                    if (it > 0) wait_dbg(&s.aux_empty[st ^ 1], ((it - 1) >> 1) & 1, 13, DBGP);
                    tab_head = h;
                    const uint4* src = reinterpret_cast<const uint4*>(p.tabg + (size_t)h * (p.tab_bytes / 2));
                    uint4* dst = reinterpret_cast<uint4*>(base + OFF_TAB);
                    for (int n = lane; n < p.tab_bytes / 16; n += 32) dst[n] = __ldg(src + n);
                }
                if (lane == 0) { flags[st].bmax = p.tabstat[2 * h]; flags[st].bmin = p.tabstat[2 * h + 1]; }
                wait_dbg(&s.qkv_full[st], ph, 7, DBGP);
                const uint32_t ka = stage_a + st * STAGE_BYTES + Q_BYTES;
                float k2 = 0.f;   // max_j |k_j|^2: with max_i |q_i| it bounds every score of the window (Cauchy-Schwarz)
                for (int n = lane; n < p.KR * SLOT; n += 32) k2 = fmaxf(k2, row_norm2<F16>(ka + n * 64));
                k2 = warp_max(k2);
                if (lane == 0) flags[st].k2max = k2;
            } else {
                uint8_t* regk = base + OFF_REGK + st * MAXCOLS;
                uint8_t* regq = base + OFF_REGQ + st * 512;
                int diff = 0;
                if (p.region) {
                    const uint8_t* rg = p.region + (long long)win * p.N;
                    const uint8_t r0 = rg[0];
                    for (int n = lane; n < MAXCOLS; n += 32) {
                        const int row = n >> 3, sl = n & 7;
                        const bool real = sl < p.ww && row < p.KR;
                        const uint8_t r = real ? rg[row * p.ww + sl] : r0;
                        regk[n] = r;
                        diff |= (r != r0);
                    }
                    for (int n = lane; n < 512; n += 32) regq[n] = rg[min(n, p.N - 1)];
                }
                diff = __any_sync(0xffffffffu, diff);
                if (lane == 0) flags[st].masked = diff;
                wait_dbg(&s.qkv_full[st], ph, 8, DBGP);
                const uint32_t qa = stage_a + st * STAGE_BYTES;
                float q2 = 0.f;
                for (int n = lane; n < p.N; n += 32) q2 = fmaxf(q2, row_norm2<F16>(qa + n * 64));
                q2 = warp_max(q2);
                if (lane == 0) flags[st].q2max = q2;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_full[st]);
        }
    } else {
        // ===================== softmax + epilogue warps =====================
        if (VSW_ATTN2_SETMAXNREG) tc::setmaxnreg_inc<152>();
        const int q = warp & 3, g = warp >> 2;      // TMEM lane quadrant, group
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t tab_a = tc::smem_u32(base + OFF_TAB), koff_a = tc::smem_u32(koff);
        float* xmax = reinterpret_cast<float*>(base + OFF_XMAX);
        float* lpart = reinterpret_cast<float*>(base + OFF_LPART);
        int c = 0;            // global block counter (all warps walk the whole sequence, a group acts on c % NGRP == g)
        int cown = g;         // blocks until this group's next one
        int otile = 0;        // exp tiles seen
        int etile = 0;        // exact tiles seen
        int it = 0;
        long long pe0 = 0, pe1 = 0, pe2 = 0, pe3 = 0, pe4 = 0, pe5 = 0;
        int pend_tile = -1, pend_b = 0, pend_h = 0, pend_i = 0; float pend_mx = 0.f; bool pend_active = false;
        const long long e_t0 = DBGP ? clock64() : 0;
        long long t_prev = e_t0, pe7 = 0, pe8 = 0, pe9 = 0;
        // epilogue of exp tile number `tile` (item (eb, eh), this thread's query row ei): O / l -> 16-bit -> global, log-sum-exp
        auto epilogue = [&](int tile, int eb, int eh, int ei, float emx, bool eactive) {
            const bool prof = DBGP && blockIdx.x == 0 && warp == 0 && lane == 0;
            long long t_d = 0;
            if (prof) t_d = clock64();
            // o_full is a RING of four barriers: only the tile's own group waits here, so a single barrier's parity would be
            // ambiguous for a group that skipped a tile
            wait_dbg(&s.o_full[tile & 3], (tile >> 2) & 1, 12, DBGP);
            tc::tc_fence_after();
            if (eactive) {
                uint32_t o[32];
                tc::tmem_ld_32x32(tmem + lane_base + O_COL + (tile & 1) * HD, o);
                tc::tmem_ld_wait();
                float l = 0.f;   // row sum of the unrounded exponentials: one partial per key block (summed in a fixed order)
                for (int k2 = 0; k2 < p.nkb; ++k2) l += lpart[((tile & 3) * MAXNKB + k2) * QT + row];
                const float inv = __fdividef(1.0f, l);
                if (ei < p.N) {
                    uint4* dst = reinterpret_cast<uint4*>((uint16_t*)p.out + ((long long)eb * p.N + ei) * C + eh * HD);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        uint4 u;
                        u.x = pack16<F16>(__uint_as_float(o[8 * v4 + 0]) * inv, __uint_as_float(o[8 * v4 + 1]) * inv);
                        u.y = pack16<F16>(__uint_as_float(o[8 * v4 + 2]) * inv, __uint_as_float(o[8 * v4 + 3]) * inv);
                        u.z = pack16<F16>(__uint_as_float(o[8 * v4 + 4]) * inv, __uint_as_float(o[8 * v4 + 5]) * inv);
                        u.w = pack16<F16>(__uint_as_float(o[8 * v4 + 6]) * inv, __uint_as_float(o[8 * v4 + 7]) * inv);
                        dst[v4] = u;
                    }
                    p.lse[((long long)eb * p.nH + eh) * p.N + ei] = (emx + __log2f(l)) * LN2;
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.o_empty[tile & 1]);
            if (prof) { pe3 += clock64() - t_d; pe4 += 1; }
        };
        for (int w = w_begin; w < w_end; ++w, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int h = w / p.B_, b_ = w - h * p.B_;
            const long long t_aux = (DBGP && blockIdx.x == 0) ? clock64() : 0;
            wait_dbg(&s.aux_full[st], ph, 9, DBGP);
            if (DBGP && blockIdx.x == 0) pe5 += clock64() - t_aux;
            const bool masked = flags[st].masked != 0;
            const bool exact = is_exact(st);
            const uint32_t regk_a = tc::smem_u32(base + OFF_REGK + st * MAXCOLS);
            const uint8_t* regq = base + OFF_REGQ + st * 512;
            for (int t = 0; t < p.nq; ++t) {
                const int i = t * QT + row;
                const bool active = t * QT + q * 32 < p.N;        // warp-uniform: any real row in this warp?
                const int ic = min(i, p.N - 1);
                const uint32_t arow = tab_a + reinterpret_cast<const uint32_t*>(base + OFF_AROW)[t * QT + row];
                const uint32_t regi4 = masked ? (uint32_t)regq[ic] * 0x01010101u : 0u;
                float mx = (F16 && !exact) ? f16_bound(st) - F16_TOP : 0.f;   // fp16 single pass: uniform shift instead of the row maximum
                const int xslot = etile & 3;
                for (int pass = exact ? 0 : 1; pass < 2; ++pass) {
                    if (exact && pass == 1) {
                        wait_dbg(&s.max_done[xslot], (etile >> 2) & 1, 10, DBGP);
                        mx = -INFINITY;
                        for (int kb = 0; kb < p.nkb; ++kb) mx = fmaxf(mx, xmax[(xslot * MAXNKB + kb) * QT + row]);
                    }
                    for (int kb = 0; kb < p.nkb; ++kb, ++c) {
                        if (cown-- != 0) continue;
                        cown = NGRP - 1;
                        const int buf = c & (NSB - 1);
                        const bool prof = DBGP && blockIdx.x == 0 && warp == 0 && lane == 0;
                        long long t_a = 0, t_b = 0;
                        if (prof) { t_a = clock64(); pe7 += t_a - t_prev; }
                        wait_dbg(&s.s_full[buf], (c / NSB) & 1, 11, DBGP);
                        tc::tc_fence_after();
                        if (prof) t_b = clock64();
                        BlockArgs a;
                        a.sbuf = tmem + lane_base + buf * BW;
                        a.arow = arow; a.koff_a = koff_a + kb * KRB * 4; a.regk_a = regk_a + kb * BW; a.regi4 = regi4;
                        a.nrows = rows_of(kb); a.scale_log2 = p.scale_log2; a.nm = -mx;
                        if (pass == 0) {
                            float m = -INFINITY;
                            if (active) m = masked ? block_max<WW, true>(a) : block_max<WW, false>(a);
                            xmax[(xslot * MAXNKB + kb) * QT + row] = m;
                            tc::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) { tc::mbar_arrive(&s.max_done[xslot]); tc::mbar_arrive(&s.p_full[buf]); }
                            continue;
                        }
                        if (active) {
                            float ls;
                            // one instantiation per mask flavour: the single-pass case runs the same code with nm = 0 (bf16) or the
                            // uniform shift (fp16).  A separate no-offset variant saved one FADD per score but pushed the kernel over
                            // its 128-register budget (900 bytes of spills through a ~1 KB L1).
                            ls = masked ? block_exp<WW, true, true, F16>(a) : block_exp<WW, false, true, F16>(a);
                            lpart[((otile & 3) * MAXNKB + kb) * QT + row] = ls;
                        }
                        tc::tmem_st_wait();
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&s.p_full[buf]);
                        if (prof) { const long long t_c = clock64(); pe0 += t_b - t_a; pe1 += t_c - t_b; pe2 += 1; t_prev = t_c; }
                        // ---- deferred epilogue: the tile whose last block this group finished ONE OWN BLOCK AGO.  By now its P.V
                        // MMAs have long retired, so the wait below costs nothing (done right after the tile's last block, the group
                        // sat ~4 k cycles waiting for the tensor core to get to it)
                        if (pend_tile >= 0) { epilogue(pend_tile, pend_b, pend_h, pend_i, pend_mx, pend_active); pend_tile = -1; }
                        if (kb == p.nkb - 1) { pend_tile = otile; pend_b = b_; pend_h = h; pend_i = i; pend_mx = mx; pend_active = active; }
                    }
                }
                ++otile;
                if (exact) ++etile;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_empty[st]);
        }
        if (pend_tile >= 0) epilogue(pend_tile, pend_b, pend_h, pend_i, pend_mx, pend_active);   // the group's very last tile
        if (DBGP && blockIdx.x == 0 && warp == 0 && lane == 0) { DBGP[0] += pe0; DBGP[1] += pe1; DBGP[2] += pe2; DBGP[3] += pe3; DBGP[4] += pe4; DBGP[5] += pe5; DBGP[6] += clock64() - e_t0; DBGP[7] += pe7; }
    }

    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

#undef DBGP

// ---- the per-call table kernel ---------------------------------------------------------------------------------------------
// tabg[h][w_i][dd][dh][slot]  =  table[((dd) NH + dh) NW + (w_i - slot + ww - 1)][h] * log2e   (bf16; -inf in dummy slots)
// TRANSPOSED: slot is the QUERY's w and the leading index the KEY's:  table[... + (slot - w_j + ww - 1)]  (backward kernel)
// tabstat[h] = {max, min} of the column (x log2e); poison = 1 if rowcode / colcode are not the dense codes of (wdc, wh, ww).
template <typename T>
__global__ void attn2_table_kernel(const T* __restrict__ table, const int32_t* __restrict__ rowcode,
                                   const int32_t* __restrict__ colcode, int N, int nH, int L, int wdc, int wh, int ww,
                                   int transposed, int blk, uint16_t* __restrict__ tabg, float* __restrict__ tabstat, int* __restrict__ poison) {
    // blk = rows per w block (>= ND * NH; the backward pads it so that neighbouring w blocks start 4 rows apart mod 8 --
    // the two key columns of a quarter-warp then read distinct shared-memory banks)
    const int h = blockIdx.x;
    const int ND = 2 * wdc - 1, NH = 2 * wh - 1, NW = 2 * ww - 1;
    const int total = ww * blk * SLOT;
    float mx = -INFINITY, mn = INFINITY;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int sl = idx & 7;
        int r = idx >> 3;
        const int wl = r / blk; r -= wl * blk;
        const int dh = r % NH;
        const int dd = r / NH;
        uint16_t o = 0xFF80;   // bf16 -inf
        if (sl < ww && dd < ND) {
            const int dw = transposed ? (sl - wl + ww - 1) : (wl - sl + ww - 1);
            const float v = to_f<T>(table[(long long)((dd * NH + dh) * NW + dw) * nH + h]) * LOG2E;
            mx = fmaxf(mx, v); mn = fminf(mn, v);
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            o = *reinterpret_cast<const uint16_t*>(&b);
        }
        tabg[(size_t)h * total + idx] = o;
    }
    __shared__ float smx[32], smn[32];
    mx = warp_max(mx); mn = -warp_max(-mn);
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); }
        tabstat[2 * h] = mx; tabstat[2 * h + 1] = mn;
    }
    if (h == 0) {
        bool ok = (rowcode[0] + colcode[0] == (L - 1) / 2);
        for (int n = threadIdx.x; n < N && ok; n += blockDim.x) {
            const int e = (n / (wh * ww)) * NH * NW + ((n / ww) % wh) * NW + n % ww;
            ok = rowcode[n] - rowcode[0] == e && colcode[0] - colcode[n] == e;
        }
        if (!ok) *poison = 1;
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// VSW_ATTN_DEBUG=1 (read once per process): phase counters of CTA 0, printed to stderr and cleared at every launch
long long* attn2_debug_buffer(const char* which) {
    static long long* dbg = nullptr;
    static std::once_flag once;
    std::call_once(once, [] { if (VSW_ATTN2_PROF && getenv("VSW_ATTN_DEBUG")) { cudaMalloc(&dbg, 256); cudaMemset(dbg, 0, 256); } });
    if (!dbg) return nullptr;
    long long h[32];
    cudaMemcpy(h, dbg, 256, cudaMemcpyDeviceToHost);
    if (h[2]) {
        fprintf(stderr, "[vsw attn2, the launch before this %s launch] exp warp 0: blocks=%lld wait_s=%lld wait_ds=%lld work=%lld (cycles per block) lifetime=%lld"
                        " epilogues=%lld x %lld side-data wait=%lld gaps/dq=%lld aux=%lld gap(in tile)=%lld gap(tile start)=%lld gap(item start)=%lld gap(in tile, loop only)=%lld\n", which, h[2], h[0] / h[2], h[3] / h[2], h[1] / h[2], h[6], h[4], h[4] ? h[5] / h[4] : 0, h[5], h[7], h[14], h[15], h[17], h[18], h[19]);
        fprintf(stderr, "      mma thread: blocks=%lld wait_p=%lld wait_loads=%lld other=%lld issue_phase=%lld (cycles per block) lifetime=%lld drained=%lld\n",
                h[12], h[8] / (h[12] + 1), h[9] / (h[12] + 1), h[10] / (h[12] + 1), h[13] / (h[12] + 1), h[20], h[16]);
    }
    cudaMemset(dbg, 0, 256);
    return dbg;
}

// grow-only scratch buffer per (device, stream)
uint8_t* scratch_for(cudaStream_t st, size_t bytes) {
    struct Buf { uint8_t* p; size_t n; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> bufs;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    Buf& b = bufs[{dev, st}];
    if (b.n < bytes) {
        if (b.p) { cudaStreamSynchronize(st); cudaFree(b.p); }   // rare: only when a larger table than ever before is needed
        b.p = nullptr; b.n = 0;
        const size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        cudaError_t e = cudaMalloc((void**)&b.p, want);
        if (e != cudaSuccess) { set_error("attn: scratch cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); b.p = nullptr; return nullptr; }
        b.n = want;
    }
    return b.p;
}

struct Geometry { int wdc, wh, ww, KR, nq, nkb, tab_bytes, blk; };
// the shapes this kernel family takes: head_dim 32, window rows of ww <= 8 tokens, N a whole number of rows, N <= 448
bool geometry_of(int N, int hd, int L, int window_dims, Geometry* g) {
    const int wh = (window_dims >> 8) & 0xFF, ww = (window_dims >> 16) & 0xFF;
    if (hd != HD || wh < 1 || ww < 1 || ww > SLOT || N < 1 || N > 448 || N % ww != 0) return false;
    const int NH = 2 * wh - 1, NW = 2 * ww - 1;
    if (L % (NH * NW) != 0) return false;
    const int ND = L / (NH * NW);
    if (ND % 2 == 0) return false;
    const int wdc = (ND + 1) / 2;
    const int KR = N / ww;
    if (KR > MAXKR || N > wdc * wh * ww) return false;
    g->wdc = wdc; g->wh = wh; g->ww = ww; g->KR = KR;
    g->nq = (N + QT - 1) / QT; g->nkb = (KR + KRB - 1) / KRB;
    // rows per w_i block, padded to 7 mod 8: the 16-byte bias vector of query (h, w) then starts in bank group (h - w) mod 8,
    // and 8 consecutive queries of a 7-wide window row order hit 8 different groups (no shared-memory bank conflicts)
    g->blk = ND * NH;
    while (g->blk % 8 != 7) ++g->blk;
    g->tab_bytes = ww * g->blk * SLOT * 2;
    return g->tab_bytes <= TAB_MAX_BYTES;
}

bool make_qkv_maps(CUtensorMap* tmQ, CUtensorMap* tmKV, const void* qkv, int B_, int N, int C3, const Geometry& g) {
    // 64-byte L2 promotion: a head's slice of a qkv row is 64 bytes; with the GEMMs' 256 the forward read 2.3x its algorithmic
    // bytes from DRAM at 4 heads (three other heads' slices per miss, evicted before their CTAs came by) -- measured, same speed
    constexpr int promo = 64;
    if (!make_tmap_3d_bf16(tmQ, qkv, B_, N, C3, C3, (uint64_t)N * C3, QT, HD, 64, promo)) return false;
    const uint64_t dims[4] = {(uint64_t)C3, (uint64_t)g.ww, (uint64_t)g.KR, (uint64_t)B_};
    const uint64_t strides[4] = {1, (uint64_t)C3, (uint64_t)g.ww * C3, (uint64_t)N * C3};
    const uint32_t box[4] = {HD, SLOT, KRB, 1};
    return make_tmap_nd_bf16(tmKV, qkv, 4, dims, strides, box, 64, promo);
}

}  // namespace

int tc2_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                 const uint8_t* region, void* out, float* lse, int B_, int nW, int N, int nH, int hd, int L,
                 float scale, int window_dims, int dtype, cudaStream_t st) {
    Geometry g;
    if ((dtype != VSW_BF16 && dtype != VSW_F16) || !geometry_of(N, hd, L, window_dims, &g) || !aligned16(qkv) || !aligned16(out)) {
        set_error("tcgen05 window attention: needs bf16/fp16, head_dim 32, the configured window (rows of <= 8 tokens) as "
                  "layout hint and N <= 448 a whole number of window rows (hd=%d N=%d L=%d window_dims=0x%x)", hd, N, L, window_dims);
        return VSW_ERR_UNSUPPORTED;
    }
    const int C = nH * HD;
    CUtensorMap tmQ, tmKV;
    if (!make_qkv_maps(&tmQ, &tmKV, qkv, B_, N, 3 * C, g)) return VSW_ERR_CUDA;
    // scratch (shifted tables + column statistics + poison flag): one grow-only buffer per (device, stream), reused by every
    // call on that stream -- stream order makes that safe, and it keeps allocator calls out of the launch path
    const size_t tab_total = (size_t)nH * g.tab_bytes;
    uint8_t* scratch = scratch_for(st, tab_total + (size_t)nH * 8 + 16);
    if (!scratch) return VSW_ERR_CUDA;
    cudaError_t e = cudaSuccess;
    uint16_t* tabg = (uint16_t*)scratch;
    float* tabstat = (float*)(scratch + tab_total);
    int* poison = (int*)(scratch + tab_total + (size_t)nH * 8);
    cudaMemsetAsync(poison, 0, 4, st);
    if (dtype == VSW_BF16)
        attn2_table_kernel<__nv_bfloat16><<<nH, 1024, 0, st>>>((const __nv_bfloat16*)table, rowcode, colcode, N, nH, L, g.wdc, g.wh, g.ww, 0, g.blk, tabg, tabstat, poison);
    else
        attn2_table_kernel<__half><<<nH, 1024, 0, st>>>((const __half*)table, rowcode, colcode, N, nH, L, g.wdc, g.wh, g.ww, 0, g.blk, tabg, tabstat, poison);
    int rc = check_launch("attn2_table");
    if (rc) return rc;
    FwdParams p{};
    p.tabg = tabg; p.tabstat = tabstat; p.poison = poison; p.region = region; p.out = out; p.lse = lse;
    p.B_ = B_; p.nW = nW; p.N = N; p.nH = nH; p.wh = g.wh; p.ww = g.ww; p.KR = g.KR; p.nq = g.nq; p.nkb = g.nkb;
    p.tab_bytes = g.tab_bytes; p.NHt = 2 * g.wh - 1; p.wdc = g.wdc; p.blk = g.blk;
    p.scale_log2 = scale * LOG2E;
    { static const int fx = getenv("VSW_ATTN2_EXACT") ? atoi(getenv("VSW_ATTN2_EXACT")) : 0; p.force_exact = fx; }   // 1: always the two-pass softmax (tests)
    p.pair = 1;
    p.dbg = attn2_debug_buffer("fwd");
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = B_ * nH;
    const int grid = items < sms ? items : sms;
#define VSW_LAUNCH_FWD2(WWV, F16V)                                                                                        \
    do {                                                                                                                  \
        auto kern = attn2_fwd_kernel<WWV, F16V>;                                                                          \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);                            \
        if (e == cudaSuccess) kern<<<grid, FWD_THREADS, FWD_SMEM, st>>>(tmQ, tmKV, p);                                    \
    } while (0)
    if (dtype == VSW_BF16) { if (g.ww == 7) VSW_LAUNCH_FWD2(7, false); else VSW_LAUNCH_FWD2(8, false); }
    else { if (g.ww == 7) VSW_LAUNCH_FWD2(7, true); else VSW_LAUNCH_FWD2(8, true); }
#undef VSW_LAUNCH_FWD2
    if (e != cudaSuccess) { set_error("attn fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
    return check_launch("attn2_fwd");
}

}  // namespace vsw

#include "attn_tc2_bwd.inl"
