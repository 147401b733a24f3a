// tcgen05 / TMEM fused window attention for sm_100a (bf16, head_dim 32, N <= 448 tokens per window).
// Reference op: WindowAttention3D.forward, visbackbone/video_swin.py:149-169.
//
// One persistent CTA per SM walks a contiguous range of (head, window) work items (head-major, so the bias column stays
// in shared memory).  Per item the whole Q, K, V of the head (N x 32 bf16 each) is TMA-loaded into 64-byte-swizzled shared
// memory (double-buffered across items, 3-D tensor map so rows >= N read as zero), then for each 128-query tile:
//   S = Q K^T        tcgen05.mma (SS), fp32 accumulators in TMEM over the whole key range (no online rescale), issued as two
//                    key halves so that S of one half / the next tile is computed while the softmax warps work on the other
//   softmax          12 warps (3 column parts x 4 TMEM lane quadrants), one thread per row and part:
//                    p = exp2(S*scale*log2e + bias*log2e [+ mask]) with the relative-position bias gathered from a padded
//                    shared-memory copy of the table column (index = rowcode[i] + colcode[j]) and the shift mask derived from
//                    uint8 region ids -- nothing N x N ever exists in HBM.  No max is subtracted when |score| and |bias| are
//                    provably <= 50 log2 units (Cauchy-Schwarz bound from |q_i|, max|k_j|); otherwise an exact two-pass path.
//                    P is written back to TMEM as packed bf16, aliasing S
//   O = P V, l = P 1 tcgen05.mma (TS: A = P from TMEM, B = V MN-major from smem / a tile of ones): 128 x 32 (+16) fp32 in TMEM
//   epilogue         O / l -> bf16 -> global, log-sum-exp -> global (for the recompute backward); deferred by one tile
// Two aux warps stage the per-item side data (bias column when the head changes, region ids, |q|^2, max|k|^2).
#include <stdlib.h>
#include <vector>
#include "attn.cuh"
#include "tc_common.cuh"

// phase counters (VSW_ATTN_DEBUG=1 VSW_ATTN_DEBUG_DUMP=1) exist only in builds with -DVSW_ATTN_PROF=1 (VSW_NVCC_EXTRA of build.py)
#ifndef VSW_ATTN_PROF
#define VSW_ATTN_PROF 0
#endif
#if VSW_ATTN_PROF
#define ACLK() clock64()
#else
#define ACLK() 0LL
#endif

namespace vsw {
namespace {

constexpr int HD = 32;
constexpr int QT = 128;                       // query rows per tile
constexpr int MAXROWS = 512;                  // smem rows per operand
constexpr int AUXROWS = 448;                  // per-token side arrays (N <= 448)
constexpr int OPER_BYTES = MAXROWS * HD * 2;  // 32 KB
constexpr int STAGE_BYTES = 3 * OPER_BYTES;   // Q, K, V
constexpr int BOX_BYTES = QT * HD * 2;        // 8 KB per TMA box
constexpr int NTHREADS = 384;                 // backward: warp 0 TMA, 1 MMA, 2 aux, 3 idle, 4..11 softmax
constexpr int NPARTS = 3;                     // forward: column parts per row (3 softmax warps per TMEM lane quadrant)
constexpr int NHALF = 2;                      // forward: the key range is processed as two independently pipelined halves
constexpr int FWD_THREADS = 128 + NPARTS * 128;  // warp 0 TMA, 1 MMA, 2 aux, 3 idle, 4.. softmax
constexpr int S_COL = 0, O_COL = 480, L_COL = 464, TMEM_COLS = 512;   // L: 16 (identical) columns of row sums, needs Npad <= 464
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// On-chip layout of the bias column (and of the backward's histograms).  The reference's table index is
// (dd)(2wh-1)(2ww-1) + (dh)(2ww-1) + dw  (video_swin.py:127-141); with the window's h-stride 2ww-1 = 13, rows h and h+2..3 of a
// 7x7 plane alias the same shared-memory banks inside every 32-token group, so each bias gather / histogram update is a 2-way
// bank conflict.  Padding the h-stride to R2 + pad with (R2 + pad) % 32 == ww makes a plane's 49 tokens hit consecutive banks;
// pad is accepted only if the padded index is still injective (planes interleave into the gaps: 169 = 4*39 + 13) -- the table
// grows by 12 %, the conflict replays drop from 1.92 to 1.54 wavefronts per access.
struct BiasLayout { int wh, ww, R1, R2, pad, Lphys, M1, M2; };   // pad == 0: dense layout; M1/M2: 2^20-scaled reciprocals of R1/R2
__host__ __device__ inline int bias_dh(int l, int R1, int R2, int M1, int M2) {   // (l % R1) / R2 without a division
    const int l1 = l - (int)(((unsigned)l * (unsigned)M1) >> 20) * R1;
    return (int)(((unsigned)l1 * (unsigned)M2) >> 20);
}
inline BiasLayout bias_layout(int window_dims, int L) {
    BiasLayout b{0, 0, 0, 0, 0, L, 0, 0};
    const int wh = (window_dims >> 8) & 0xFF, ww = (window_dims >> 16) & 0xFF;
    if (wh < 1 || ww < 1) return b;
    const int R2 = 2 * ww - 1, R1 = (2 * wh - 1) * R2;
    if (L % R1 != 0) return b;                       // not this window's table
    const int nd = L / R1;                            // 2wd-1
    const int pad = ((ww - R2) % 32 + 32) % 32;
    if (pad == 0) return b;
    const int Lphys = L + pad * (2 * wh - 2);
    if (Lphys > 3 * L / 2) return b;                  // keep the growth modest
    const int M1 = ((1 << 20) + R1 - 1) / R1, M2 = ((1 << 20) + R2 - 1) / R2;
    std::vector<char> seen((size_t)Lphys, 0);         // injectivity of l -> l + pad * ((l % R1) / R2), and the reciprocals
    for (int l = 0; l < L; ++l) {
        if (L >= 4096 || bias_dh(l, R1, R2, M1, M2) != (l % R1) / R2) return b;
        const int lp = l + pad * ((l % R1) / R2);
        if (seen[lp]) return b;
        seen[lp] = 1;
    }
    (void)nd;
    b.wh = wh; b.ww = ww; b.R1 = R1; b.R2 = R2; b.pad = pad; b.Lphys = Lphys; b.M1 = M1; b.M2 = M2;
    return b;
}

struct FwdParams {
    const __nv_bfloat16* qkv; const __nv_bfloat16* table; const int32_t* rowcode; const int32_t* colcode; const uint8_t* region;
    __nv_bfloat16* out; float* lse;
    int B_, nW, N, nH, L, Lpad;
    float scale_log2;
    int Npad, nq;
    int h0;           // keys [0, h0) form half 0, [h0, Npad) half 1 (both multiples of 16)
    int wh, ww, R1, R2, pad, M1, M2;   // padded on-chip bias layout (pad == 0: dense); see bias_layout()
    long long* dbg;   // optional per-phase cycle counters (profiling builds only)
};

struct Smem {
    uint8_t* stage[2];
    float* tab[2];
    uint8_t* reg[2];
    int* rc; int* cc;
    float* xmax;                // [4][128] exact-path row-max exchange
    uint8_t* ones;              // 1 KB of bf16 1.0: B operand of the row-sum MMA (l = P . 1 on the tensor core)
    float* maxbias; int* masked;  // [2]
    float* k2max;                 // [2] max_j |k_j|^2 of the staged item; [2..3] = min of the bias column (x log2e)
    float* q2[2];                 // [2][512] |q_i|^2 of the staged item
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
    s.rc = (int*)p; p += AUXROWS * 4;
    s.cc = (int*)p; p += AUXROWS * 4;
    s.xmax = (float*)p; p += 4 * QT * 4;
    s.ones = p; p += 1024;
    s.maxbias = (float*)p; p += 8;
    s.masked = (int*)p; p += 8;
    s.k2max = (float*)p; p += 16;
    s.qkv_full = (uint64_t*)p; p += 16;
    s.qkv_empty = (uint64_t*)p; p += 16;
    s.aux_full = (uint64_t*)p; p += 16;
    s.aux_empty = (uint64_t*)p; p += 16;
    s.s_full = (uint64_t*)p; p += 16;   // [2]: one per key half
    s.p_full = (uint64_t*)p; p += 16;   // [2]
    s.o_full = (uint64_t*)p; p += 8;
    s.tmem_slot = (uint32_t*)p; p += 8;
    s.reg[0] = p; p += AUXROWS;
    s.reg[1] = p; p += AUXROWS;
    s.q2[0] = (float*)p; p += AUXROWS * 4;
    s.q2[1] = (float*)p; p += AUXROWS * 4;
    return s;
}
size_t fwd_smem_bytes(int Lpad) {
    return 1024 + 2 * (size_t)STAGE_BYTES + 2 * (size_t)Lpad * 4 + 2 * AUXROWS * 4 + 4 * QT * 4 + 1024 + 32 + 4 * 16 + 16 + 16 + 8 + 8 +
           2 * AUXROWS + 2 * AUXROWS * 4 + 16;
}


// ---- softmax helpers -------------------------------------------------------------------------
constexpr float MASKV = -100.0f * LOG2E;   // the reference's additive -100 (video_swin.py:304-306), in log2 units

// 16 consecutive ints / bytes from 16-byte aligned shared memory
__device__ __forceinline__ void lds16i(uint32_t addr, uint32_t (&v)[16]) {
    const uint4 a = tc::lds_u4(addr), b = tc::lds_u4(addr + 16), c = tc::lds_u4(addr + 32), d = tc::lds_u4(addr + 48);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
}
// per-byte "region differs" flags of 16 keys: 0xFF where reg[j] != regi
__device__ __forceinline__ void neq16(uint32_t reg16_addr, uint32_t regi4, uint32_t (&w)[4]) {
    const uint4 r = tc::lds_u4(reg16_addr);
    w[0] = __vcmpne4(r.x, regi4); w[1] = __vcmpne4(r.y, regi4); w[2] = __vcmpne4(r.z, regi4); w[3] = __vcmpne4(r.w, regi4);
}
__device__ __forceinline__ float mask_add(float v, const uint32_t (&w)[4], int e) {
    return (w[e >> 2] & (0xFFu << (8 * (e & 3)))) ? v + MASKV : v;
}

// squared L2 norm of one head row (64 B) already staged in shared memory (the 64-byte swizzle only permutes its 16-byte units)
__device__ __forceinline__ float row_norm2_smem(uint32_t row_addr) {
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const uint4 a = tc::lds_u4(row_addr + v * 16);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = tc::unpack_bf16(w[e]);
            acc = fmaf(x.x, x.x, acc);
            acc = fmaf(x.y, x.y, acc);
        }
    }
    return acc;
}

// The Npad key columns of a row are cut into two halves (S of one half is recomputed by the tensor core for the
// next query tile while the softmax warps work on the other), each half into NPARTS parts (one softmax warp per
// TMEM lane quadrant each); all boundaries are multiples of 16.  part_off = offset of part `part` inside a half.
__device__ __forceinline__ int part_off(int len, int part) { return ((((len >> 4) * part) / NPARTS) << 4); }

// ---- forward softmax over the full chunks [cbeg, cfull) of one row ---------------------------------------------
// exact-path pass 1 (rare): true row max of  acc*scale_log2 + bias (+mask)  -- a compact un-pipelined loop
__device__ __forceinline__ float fwd_rowmax_exact(bool masked, uint32_t srow, int cbeg, int cfull, uint32_t tabrow, uint32_t cc_a,
                                                  uint32_t reg_a, uint32_t regi4, float scale_log2) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = cbeg; c < cfull; c += 16) {
        uint32_t r[16], cj[16], nq4[4] = {0, 0, 0, 0};
        tc::tmem_ld_32x16(srow + c, r);
        lds16i(cc_a + c * 4, cj);
        if (masked) neq16(reg_a + c, regi4, nq4);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
            mx = fmaxf(mx, mask_add(fmaf(__uint_as_float(r[e]), scale_log2, tc::lds_f32(tabrow + cj[e])), nq4, e));
    }
    return mx;
}

// pass 2: p = exp2(acc*scale_log2 + bias[rowcode+colcode] (+mask)) -> packed bf16 back into TMEM (the row sum comes from the
// tensor core: l = P . 1 is accumulated next to O = P . V)
template <bool MASKED>
__device__ __forceinline__ void fwd_exp(uint32_t srow, uint32_t prow, int cbase, int cbeg, int cfull, uint32_t tabrow,
                                        uint32_t cc_a, uint32_t reg_a, uint32_t regi4, float scale_log2) {
    if (cbeg >= cfull) return;
    uint32_t a[16], b[16];
    auto body = [&](const uint32_t (&r)[16], int c) {
        uint32_t cj[16], nq4[4];
        lds16i(cc_a + c * 4, cj);
        if (MASKED) neq16(reg_a + c, regi4, nq4);
        float tb[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) tb[e] = tc::lds_f32(tabrow + cj[e]);
        uint32_t pw[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            float v0 = fmaf(__uint_as_float(r[e]), scale_log2, tb[e]);
            float v1 = fmaf(__uint_as_float(r[e + 1]), scale_log2, tb[e + 1]);
            if (MASKED) { v0 = mask_add(v0, nq4, e); v1 = mask_add(v1, nq4, e + 1); }
            const float p0 = tc::ex2_approx(v0), p1 = tc::ex2_approx(v1);
            pw[e / 2] = tc::pack_bf16(p0, p1);
        }
        tc::tmem_st_32x8(prow + (c - cbase) / 2, pw);
    };
    tc::tmem_ld_32x16(srow + cbeg, a);
    tc::tmem_ld_wait();
    for (int c = cbeg; c < cfull; c += 32) {
        const bool hb = c + 16 < cfull, ha = c + 32 < cfull;
        if (hb) tc::tmem_ld_32x16(srow + c + 16, b);
        body(a, c);
        if (hb) {
            tc::tmem_ld_wait();
            if (ha) tc::tmem_ld_32x16(srow + c + 32, a);
            body(b, c + 16);
            if (ha) tc::tmem_ld_wait();
        }
    }
}

// exact-path variant (row max subtracted): rare, so a compact un-pipelined loop that does not weigh on the hot loop's registers
__device__ __forceinline__ void fwd_exp_sub(bool masked, uint32_t srow, uint32_t prow, int cbase, int cbeg, int cfull,
                                             uint32_t tabrow, uint32_t cc_a, uint32_t reg_a, uint32_t regi4, float scale_log2,
                                             float nm) {
#pragma unroll 1
    for (int c = cbeg; c < cfull; c += 16) {
        uint32_t r[16], cj[16], nq4[4] = {0, 0, 0, 0}, pw[8];
        tc::tmem_ld_32x16(srow + c, r);
        lds16i(cc_a + c * 4, cj);
        if (masked) neq16(reg_a + c, regi4, nq4);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            const float v0 = mask_add(fmaf(__uint_as_float(r[e]), scale_log2, tc::lds_f32(tabrow + cj[e])) + nm, nq4, e);
            const float v1 = mask_add(fmaf(__uint_as_float(r[e + 1]), scale_log2, tc::lds_f32(tabrow + cj[e + 1])) + nm, nq4, e + 1);
            const float p0 = tc::ex2_approx(v0), p1 = tc::ex2_approx(v1);
            pw[e / 2] = tc::pack_bf16(p0, p1);
        }
        tc::tmem_st_32x8(prow + (c - cbase) / 2, pw);
    }
}

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    const Smem s = carve(base, p.Lpad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.nH * HD;
    // Work items = (head, window) pairs in HEAD-MAJOR order, one contiguous range per CTA: a CTA then stays on one head
    // for (almost) all its items, so the bias column is loaded into shared memory once instead of once per item.
    const int items = p.B_ * p.nH;
    const int w_begin = (int)blockIdx.x * (items / (int)gridDim.x) + min((int)blockIdx.x, items % (int)gridDim.x);
    const int w_end = w_begin + items / (int)gridDim.x + ((int)blockIdx.x < items % (int)gridDim.x ? 1 : 0);

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.qkv_full[i], 1); tc::mbar_init(&s.qkv_empty[i], 1);
            tc::mbar_init(&s.aux_full[i], 2); tc::mbar_init(&s.aux_empty[i], 4 * NPARTS);
        }
        tc::mbar_init(s.o_full, 1);
        for (int i = 0; i < NHALF; ++i) { tc::mbar_init(&s.s_full[i], 1); tc::mbar_init(&s.p_full[i], 4 * NPARTS); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(s.tmem_slot, TMEM_COLS);
    for (int n = threadIdx.x; n < 256; n += FWD_THREADS) reinterpret_cast<uint32_t*>(s.ones)[n] = 0x3F803F80u;   // bf16 1.0 pairs
    // padded bias layout: only if the codes really are this window's dense (dd, dh, dw) codes (checked here, every CTA)
    int pad = p.pad;
    if (pad) {
        bool ok = p.rowcode[0] + p.colcode[0] == (p.L - 1) / 2;
        for (int n = threadIdx.x; n < p.N && ok; n += FWD_THREADS) {
            const int e = (n / (p.wh * p.ww)) * p.R1 + ((n / p.ww) % p.wh) * p.R2 + n % p.ww;
            ok = p.rowcode[n] - p.rowcode[0] == e && p.colcode[0] - p.colcode[n] == e;
        }
        if (!__syncthreads_and(ok)) pad = 0;
    }
    for (int n = threadIdx.x; n < AUXROWS; n += FWD_THREADS) {   // cc holds BYTE offsets into the fp32 table copy
        const int hn = pad ? (n / p.ww) % p.wh : 0;
        s.rc[n] = n < p.N ? p.rowcode[n] + pad * hn : 0;
        s.cc[n] = n < p.N ? (p.colcode[n] + pad * (p.wh - 1 - hn)) * 4 : 0;
    }
    tc::fence_proxy_async();   // the ones tile is read by the tensor core
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s.tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {   // elect.sync, not lane == 0: the TMA / MMA operands stay warp-uniform for ptxas
            int it = 0;
            for (int w = w_begin; w < w_end; ++w, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                const int h = w / p.B_, b_ = w - h * p.B_;
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
        // Per query tile the key range is handled as two halves with their own S accumulators, so the tensor
        // core recomputes S of one half for the next tile while the softmax warps are busy with the other:
        //   [P half 0 ready] O  = P0 V0 ; S0 = Q(t+1) K0^T      [P half 1 ready] O += P1 V1 ; S1 = Q(t+1) K1^T
        if (tc::elect_one()) {   // elect.sync, not lane == 0: the TMA / MMA operands stay warp-uniform for ptxas
            int it = 0; uint32_t pph = 0;
            const uint32_t idesc_pv = tc::idesc_bf16(QT, HD, 0, 1);
            const uint32_t idesc_l = tc::idesc_bf16(QT, 16, 0, 1);
            const uint64_t ones_desc = tc::smem_desc_sw64(tc::smem_u32(s.ones), 0, 512);   // all ones: only the 1 KB footprint matters
            const int hbeg[NHALF] = {0, p.h0}, hlen[NHALF] = {p.h0, p.Npad - p.h0};
            for (int w = w_begin; w < w_end; ++w, ++it) {
                const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
                tc::mbar_wait(&s.qkv_full[st], ph);
                tc::tc_fence_after();
                const uint32_t qa = tc::smem_u32(s.stage[st]), ka = qa + OPER_BYTES, va = ka + OPER_BYTES;
                const uint32_t vdesc0_hi = (uint32_t)(tc::smem_desc_sw64(va, 0, 512) >> 32);
                auto issue_qk = [&](int t, int hh) {
                    if (hlen[hh] > 0) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::umma_bf16(tmem + S_COL + hbeg[hh], tc::smem_desc_sw64(qa + t * BOX_BYTES + k * 32, 0, 512),
                                          tc::smem_desc_sw64(ka + hbeg[hh] * 64 + k * 32, 0, 512),
                                          tc::idesc_bf16(QT, hlen[hh], 0, 0), k);
                    }
                    tc::umma_commit(&s.s_full[hh]);
                };
                issue_qk(0, 0);
                issue_qk(0, 1);
                for (int t = 0; t < p.nq; ++t) {
                    uint32_t acc_pv = 0;
                    for (int hh = 0; hh < NHALF; ++hh) {
                        tc::mbar_wait(&s.p_full[hh], pph);
                        tc::tc_fence_after();
#pragma unroll
                        for (int part = 0; part < NPARTS; ++part) {
                            const int k00 = hbeg[hh] + part_off(hlen[hh], part), k01 = hbeg[hh] + part_off(hlen[hh], part + 1);
                            for (int key0 = k00; key0 < k01; key0 += 16) {
                                const uint64_t vd = ((uint64_t)vdesc0_hi << 32) | (((va + key0 * 64) & 0x3FFFFu) >> 4);
                                const uint32_t pcol = tmem + S_COL + k00 + (key0 - k00) / 2;
                                tc::umma_bf16_ts(tmem + O_COL, pcol, vd, idesc_pv, acc_pv);
                                tc::umma_bf16_ts(tmem + L_COL, pcol, ones_desc, idesc_l, acc_pv);   // l += P . 1 (row sums)
                                acc_pv = 1;
                            }
                        }
                        if (hh == NHALF - 1) tc::umma_commit(s.o_full);
                        if (t + 1 < p.nq) issue_qk(t + 1, hh);
                    }
                    pph ^= 1;
                }
                tc::umma_commit(&s.qkv_empty[st]);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===================== aux warps: bias-table column (x log2e), region ids, |q|^2 and max |k|^2 =====================
        // warp 2: bias column (only when the head differs from what the stage holds) and max_j |k_j|^2;
        // warp 3: region ids of the window and |q_i|^2.  The norms are taken from the TMA-staged Q / K tiles in shared memory.
        int it = 0;
        int tab_head[2] = {-1, -1};
        for (int w = w_begin; w < w_end; ++w, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int h = w / p.B_, b_ = w - h * p.B_, win = b_ % p.nW;
            long long ax_a = ACLK();
            tc::mbar_wait(&s.aux_empty[st], ph ^ 1);
            long long ax_b = ACLK();
            if (warp == 2) {
                if (tab_head[st] != h) {
                    tab_head[st] = h;
                    float mb = -INFINITY, nb = -INFINITY;   // max and -min of the bias column
                    for (int l = lane; l < p.L; l += 32) {
                        const float v = __bfloat162float(p.table[(long long)l * p.nH + h]) * LOG2E;
                        s.tab[st][pad ? l + pad * bias_dh(l, p.R1, p.R2, p.M1, p.M2) : l] = v;   // padded layout (bias_layout())
                        mb = fmaxf(mb, v);
                        nb = fmaxf(nb, -v);
                    }
                    mb = warp_max(mb);
                    nb = warp_max(nb);
                    if (lane == 0) { s.maxbias[st] = mb; s.k2max[2 + st] = -nb; }
                }
                tc::mbar_wait(&s.qkv_full[st], ph);
                const uint32_t ka = tc::smem_u32(s.stage[st]) + OPER_BYTES;
                float k2 = 0.f;   // max_j |k_j|^2: with |q_i| it bounds the scores of row i (Cauchy-Schwarz)
                for (int n = lane; n < p.N; n += 32) k2 = fmaxf(k2, row_norm2_smem(ka + n * 64));
                k2 = warp_max(k2);
                if (lane == 0) s.k2max[st] = k2;
            } else {
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
                if (lane == 0) s.masked[st] = diff;
                tc::mbar_wait(&s.qkv_full[st], ph);
                const uint32_t qa = tc::smem_u32(s.stage[st]);
                for (int n = lane; n < p.N; n += 32) s.q2[st][n] = row_norm2_smem(qa + n * 64);
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_full[st]);
            if (VSW_ATTN_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[6] += ax_b - ax_a; p.dbg[7] += ACLK() - ax_b; }
        }
    } else if (warp >= 4) {
        // ===================== softmax + epilogue warps =====================
        const int q = warp & 3, part = (warp - 4) >> 2;   // TMEM lane quadrant, column part of this warp (0..NPARTS-1)
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t srow = tmem + lane_base + S_COL;
        int cb[NHALF], ce[NHALF];                         // this warp's key columns inside each half
        {
            const int hbeg[NHALF] = {0, p.h0}, hlen[NHALF] = {p.h0, p.Npad - p.h0};
#pragma unroll
            for (int hh = 0; hh < NHALF; ++hh) {
                cb[hh] = hbeg[hh] + part_off(hlen[hh], part);
                ce[hh] = hbeg[hh] + part_off(hlen[hh], part + 1);
            }
        }
        const int nfull = p.N & ~15;                      // chunks below nfull have no padded columns
        int it = 0; uint32_t sph = 0, oph = 0;
        int pend_b = -1, pend_h = 0; float pend_mx = 0.f;   // item whose last tile still awaits its epilogue
        // ---- epilogue of tile te: O / l -> bf16 -> global; lse.  Runs one tile late (after the first half of
        //      the next tile's softmax) so the wait for the P.V MMAs is hidden; the next tile's P.V cannot
        //      start before every warp has passed this point (it needs their p_full[0] arrival).
        auto epilogue = [&](int eb, int eh, int te, float mxe) {   // (window, head) of the tile's item, tile, row max
            const int ie = te * QT + row;
            long long t_d = ACLK();
            tc::mbar_wait(s.o_full, oph); oph ^= 1;
            tc::tc_fence_after();
            long long t_e = ACLK();
            uint32_t o[16], lr[16];
            tc::tmem_ld_32x16(tmem + lane_base + O_COL + (part & 1) * 16, o);
            tc::tmem_ld_32x16(tmem + lane_base + L_COL, lr);   // 16 identical columns of the row sum
            tc::tmem_ld_wait();
            const float l = __uint_as_float(lr[0]);
            const float inv = __fdividef(1.0f, l);
            if (ie < p.N && part < 2) {
                __nv_bfloat16* dst = p.out + ((long long)eb * p.N + ie) * C + eh * HD + part * 16;
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
                if (part == 0) p.lse[((long long)eb * p.nH + eh) * p.N + ie] = (mxe + __log2f(l)) * LN2;
            }
            tc::tc_fence_before();  // O reads complete before the next p_full arrive lets P.V overwrite O
            if (VSW_ATTN_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 128) {
                long long t_f = ACLK();
                p.dbg[3] += t_e - t_d; p.dbg[4] += t_f - t_e;
            }
        };

        for (int w = w_begin; w < w_end; ++w, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            const int h = w / p.B_, b_ = w - h * p.B_;
            tc::mbar_wait(&s.aux_full[st], ph);
            const bool masked = s.masked[st] != 0;
            const float mb = s.maxbias[st];
            const float k2max = s.k2max[st];
            const bool bias_ok = mb <= 50.0f && s.k2max[2 + st] >= -50.0f;   // |bias| <= 50 in log2 units
            const float* tab = s.tab[st];
            const uint8_t* reg = s.reg[st];
            const uint32_t reg_a = tc::smem_u32(reg), cc_a = tc::smem_u32(s.cc);

            float mx_prev = 0.f;
            for (int t = 0; t < p.nq; ++t) {
                const int i = t * QT + row;
                const bool valid = i < p.N;
                const int ic = valid ? i : p.N - 1;
                const int rci = s.rc[ic];
                const uint8_t regi = masked ? reg[ic] : 0;
                const uint32_t regi4 = (uint32_t)regi * 0x01010101u;
                const bool warp_rows = t * QT + q * 32 < p.N;   // warp-uniform: any valid row in this warp?
                // Single-pass softmax when it is provably safe: |s_ij| <= |q_i| max_j|k_j| scale =: bound.  For bound <= 50 and
                // |bias| <= 50 (log2 units) every exponent lies in [-100, 100] (masked entries lower still, they flush to
                // 0 like the reference's e^-100), far inside the fp32 AND bf16 exponent range, so the exponentials need
                // no max subtraction at all (softmax = p / sum p is invariant to it; lse = log sum).  Otherwise (huge
                // logits or bias values) fall back to the exact two-pass row max.
                const float bound = sqrtf(s.q2[st][ic] * k2max) * p.scale_log2;
                const bool fast = __all_sync(0xffffffffu, bound <= 50.0f) && bias_ok;
                long long t_a = ACLK();
                tc::mbar_wait(&s.s_full[0], sph);
                if (!fast) tc::mbar_wait(&s.s_full[1], sph);
                tc::tc_fence_after();
                long long t_b = ACLK();
                float mx;
                if (fast) {
                    mx = 0.f;   // no subtraction at all: every exponent is within +-100, far inside the fp32 / bf16 range
                } else {
                    // ---- pass 1: exact row max of score*scale + bias (+ mask) over both halves
                    mx = -INFINITY;
                    if (warp_rows) {
                        const uint32_t tabrow1 = tc::smem_u32(tab + rci);
#pragma unroll 1
                        for (int hh = 0; hh < NHALF; ++hh) {
                            const int cfull = min(ce[hh], nfull);
                            mx = fmaxf(mx, fwd_rowmax_exact(masked, srow, cb[hh], cfull, tabrow1, cc_a, reg_a, regi4, p.scale_log2));
                            for (int c = max(cb[hh], cfull); c < ce[hh]; c += 16) {   // chunk with columns >= N
                                uint32_t r[16];
                                tc::tmem_ld_32x16(srow + c, r);
                                tc::tmem_ld_wait();
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    if (c + e < p.N) {
                                        float v = fmaf(__uint_as_float(r[e]), p.scale_log2, tc::lds_f32(tabrow1 + s.cc[c + e]));
                                        if (masked && reg[c + e] != regi) v += MASKV;
                                        mx = fmaxf(mx, v);
                                    }
                                }
                            }
                        }
                    }
                    s.xmax[part * QT + row] = mx;
                    tc::named_bar_sync(1 + q, 32 * NPARTS);
#pragma unroll
                    for (int k = 0; k < NPARTS; ++k) mx = fmaxf(mx, s.xmax[k * QT + row]);
                    tc::named_bar_sync(1 + q, 32 * NPARTS);   // xmax may be rewritten by the next tile
                    if (!(mx > -INFINITY)) mx = 0.f;          // rows beyond the window
                }
                long long t_c = ACLK();
                // ---- pass 2: p = exp2(s - max) -> packed bf16 into TMEM (aliasing S), row sum in fp32, half by half;
                //      after each half the MMA warp runs that half of P.V and then the half's S for the next tile
                const uint32_t tabrow = tc::smem_u32(tab + rci);
                const float nm = -mx;
#pragma unroll
                for (int hh = 0; hh < NHALF; ++hh) {
                    if (hh == 1 && fast) { tc::mbar_wait(&s.s_full[1], sph); tc::tc_fence_after(); }
                    const int qb = cb[hh], qe = ce[hh];
                    const uint32_t prow = srow + qb;   // P (packed bf16) aliases the part's own S columns
                    if (warp_rows) {
                        const int qf = min(qe, nfull);
                        if (fast)
                            if (masked) fwd_exp<true>(srow, prow, qb, qb, qf, tabrow, cc_a, reg_a, regi4, p.scale_log2);
                            else fwd_exp<false>(srow, prow, qb, qb, qf, tabrow, cc_a, reg_a, regi4, p.scale_log2);
                        else
                            fwd_exp_sub(masked, srow, prow, qb, qb, qf, tabrow, cc_a, reg_a, regi4, p.scale_log2, nm);
                        for (int c = max(qb, qf); c < qe; c += 16) {   // chunk with columns >= N
                            uint32_t r[16];
                            tc::tmem_ld_32x16(srow + c, r);
                            tc::tmem_ld_wait();
                            uint32_t pw[8];
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                float pv[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int j = c + e + u;
                                    float pe = 0.f;
                                    if (j < p.N) {
                                        float v = fmaf(__uint_as_float(r[e + u]), p.scale_log2, tc::lds_f32(tabrow + s.cc[j])) + nm;
                                        if (masked && reg[j] != regi) v += MASKV;
                                        pe = tc::ex2_approx(v);
                                    }
                                    pv[u] = pe;
                                }
                                pw[e / 2] = tc::pack_bf16(pv[0], pv[1]);
                            }
                            tc::tmem_st_32x8(prow + (c - qb) / 2, pw);
                        }
                    }
                    // deferred epilogue of the previous tile -- which, for tile 0, is the LAST tile of the previous item
                    if (hh == 0 && t > 0) epilogue(b_, h, t - 1, mx_prev);
                    if (hh == 0 && t == 0 && pend_b >= 0) epilogue(pend_b, pend_h, p.nq - 1, pend_mx);
                    tc::tmem_st_wait();
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&s.p_full[hh]);
                }
                sph ^= 1;
                mx_prev = mx;
                if (VSW_ATTN_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 128) {
                    long long t_d = ACLK();
                    p.dbg[0] += t_b - t_a; p.dbg[1] += t_c - t_b; p.dbg[2] += t_d - t_c; p.dbg[5] += 1;
                }
            }
            pend_b = b_; pend_h = h; pend_mx = mx_prev;   // the last tile's epilogue runs inside the next item's first tile
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_empty[st]);
        }
        if (pend_b >= 0) epilogue(pend_b, pend_h, p.nq - 1, pend_mx);   // the CTA's very last tile
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
                float scale, int window_dims, cudaStream_t st) {
    BiasLayout bl = bias_layout(window_dims, L);
    if (fwd_smem_bytes((bl.Lphys + 3) / 4 * 4) > 227 * 1024) bl = bias_layout(0, L);   // no room for the padded table
    const int Lpad = (bl.Lphys + 3) / 4 * 4;
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
    p.qkv = (const __nv_bfloat16*)qkv; p.table = (const __nv_bfloat16*)table; p.rowcode = rowcode; p.colcode = colcode; p.region = region;
    p.out = (__nv_bfloat16*)out; p.lse = lse;
    p.B_ = B_; p.nW = nW; p.N = N; p.nH = nH; p.L = L; p.Lpad = Lpad;
    p.wh = bl.wh; p.ww = bl.ww; p.R1 = bl.R1; p.R2 = bl.R2; p.pad = bl.pad; p.M1 = bl.M1; p.M2 = bl.M2;
    p.scale_log2 = scale * LOG2E;
    p.Npad = (N + 15) / 16 * 16;
    p.nq = (N + QT - 1) / QT;
    p.h0 = ((p.Npad / 16 + 1) / 2) * 16;
    {
        static long long* dbg = nullptr;
        static bool init = false;
        if (!init) {
            init = true;
            if (VSW_ATTN_PROF && getenv("VSW_ATTN_DEBUG")) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); }
        }
        p.dbg = dbg;
        if (VSW_ATTN_PROF && dbg && getenv("VSW_ATTN_DEBUG_DUMP")) {
            long long h[8];
            cudaMemcpy(h, dbg, 64, cudaMemcpyDeviceToHost);
            if (h[5]) fprintf(stderr, "[vsw attn fwd] tiles=%lld avg cycles: wait_s=%lld pass1=%lld pass2=%lld wait_o=%lld epi=%lld | aux warp per item: wait %lld work %lld\n", h[5], h[0]/h[5], h[1]/h[5], h[2]/h[5], h[3]/h[5], h[4]/h[5], h[6] * 4 / h[5], h[7] * 4 / h[5]);
            cudaMemset(dbg, 0, 64);
        }
    }
    static bool configured[kMaxDevices] = {};   // per device; benign race (the attribute is idempotent)
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attn fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
        configured[dev] = true;
    }
    const int items = B_ * nH;
    const int grid = items < kNumSMs ? items : kNumSMs;
    attn_fwd_tc_kernel<<<grid, FWD_THREADS, smem, st>>>(tm, p);
    return check_launch("attn_fwd_tc");
}

}  // namespace vsw

#include "attn_tc_bwd.inl"
