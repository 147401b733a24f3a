// Second-generation tcgen05 / TMEM recompute backward of fused window attention (included by attn_tc2.cu).
// Formulas: SURVEY A5.  Everything is held TRANSPOSED -- TMEM lane = KEY, TMEM column = QUERY -- so that the two products that
// reduce over queries take their A operand straight from tensor memory:
//   per (key tile T of 14 window rows = 128 lanes, query half-block hb of 8 window rows = 64 padded columns):
//     S^T  = K_T Q_hb^T,  dP^T = V_T dO_hb^T                        tcgen05.mma SS  -> TMEM (128 lanes x 64 columns each)
//     P^T  = exp2(S^T*scale*log2e + bias [+ mask] - lse*log2e),   dS^T = P^T (dP^T - delta)        8 exp warps, thread = key
//            P^T, dS^T -> packed 16-bit back into TMEM (aliasing S^T / dP^T);  dS^T also -> shared memory (MN-major A tile)
//     dV_T += P^T dO_hb,  dK_T += dS^T Q_hb                          tcgen05.mma TS  (A from TMEM, B = dO / Q MN-major)
//     dQ_qb += dS K_T                                                 tcgen05.mma SS  (A = the dS tile, B = K_T MN-major)
// Keys AND queries are laid out padded to 8 slots per window row (4-D tensor maps), so the bias of (key j, queries (d_i, h_i,
// 0..7)) is one 16-byte vector of the per-w_j shifted table copy, and the d(bias table) contribution of those 8 scores is one
// 32-byte read-modify-write of a per-w_j shifted fp32 histogram -- two LDS.128 + two STS.128 per 8 scores instead of eight
// scalar gathers, eight scalar loads and eight scalar stores.  Key lanes are dealt to the four TMEM lane quadrants BY w_j
// (quadrant q holds w_j = 2q, 2q+1), so the four warps of a group never touch the same histogram entry; the two groups (which
// work on alternate query half-blocks) own one histogram copy each.  Deterministic: fixed-order folds, no atomics.
namespace vsw {
namespace {

constexpr int B_THREADS = 512;          // warps 0..7: exp (2 groups x 4 lane quadrants); 8, 9: aux; 10: TMA; 11: MMA; 12..15: epilogue
constexpr int BW_EPI = 12;              // epilogue warp q = warp - 12 reads TMEM lane quadrant q (12 % 4 == 0)
constexpr int BW_AUX0 = 8, BW_AUX1 = 9, BW_TMA = 10, BW_MMA = 11;
constexpr int KTR = 14;                 // key rows per key tile (2 x 14 = 28 of the 32 lanes of a quadrant)
constexpr int QHR = 8;                  // query rows per half-block (64 padded query columns)
constexpr int NQD = 3;                  // Q / dO half-block stages
// tensor memory columns
constexpr int TB_ST = 0, TB_DP = 64, TB_STAGE = 128;        // stage s: S^T at 128 s, dP^T at 128 s + 64
constexpr int TB_DKV = 256;                                  // dK at 256 + 64 par, dV at + 32   (double-buffered by key-tile parity)
constexpr int TB_DQ = 384;                                   // dQ of query block qb at 384 + 32 qb
// shared memory
constexpr int BO_KV = 0;                                     // 2 stages x (K tile 8 KB + V tile 8 KB)
constexpr int BO_QD = 32768;                                 // NQD stages x (Q half-block 4 KB + dO half-block 4 KB)
constexpr int BO_DS = BO_QD + NQD * 8192;                    // dS chunk (even) 16 KB | zeros 16 KB | dS chunk (odd) 16 KB
constexpr int BO_TAB = BO_DS + 3 * 16384;                    // transposed shifted bias table (bf16), <= 24 KB
constexpr int BTAB_MAX_BYTES = 22016;                        // 8x7x7 window: 7 * 15 * 13 * 16 B = 21840
constexpr int BO_HIST = BO_TAB + BTAB_MAX_BYTES;             // 2 fp32 histogram copies (same layout as the table, 4 bytes per entry)
constexpr int HIST_MAX_BYTES = 2 * BTAB_MAX_BYTES;
constexpr int BO_LD = BO_HIST + 2 * HIST_MAX_BYTES;          // [2 stages][448] float2 {lse * log2e, delta}
constexpr int BO_REGQ = BO_LD + 2 * MAXCOLS * 8;             // [2][448] region id per padded query column
constexpr int BO_REGK = BO_REGQ + 2 * MAXCOLS;               // [2][512] region id per key lane (tile-major)
constexpr int BO_KOFF = BO_REGK + 2 * 512;                   // int[64]: table row index of query row rho
constexpr int BO_FLAG = BO_KOFF + 256;                     // int[2]: does the item's window have more than one region?
constexpr int BO_BARS = BO_FLAG + 16;
constexpr int BWD_SMEM = BO_BARS + 256 + 1024;
static_assert(BWD_SMEM <= 227 * 1024, "backward shared-memory layout exceeds 227 KB");

struct BwdParams2 {
    const uint16_t* tabg; const float* tabstat; const int* poison; const uint8_t* region;
    const void* out; const void* dout; const float* lse; void* dqkv; float* dbias_part;
    int B_, nW, N, nH, wh, ww, KR, nT, nhb, tab_bytes, NHt, wdc, L, groups, blk;   // blk = table rows per w_j block (padded)
    float scale, scale_log2;
    long long* dbg;
};
struct BBars {
    uint64_t *kv_full, *kv_empty, *qd_full, *qd_empty, *s_full, *p_full, *ds_free, *dkv_full, *dkv_empty, *dq_full, *dq_empty,
        *aux_full, *aux_empty;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ BBars bbars_of(uint8_t* base) {
    uint64_t* b = reinterpret_cast<uint64_t*>(base + BO_BARS);
    BBars s;
    s.kv_full = b; s.kv_empty = b + 2; s.qd_full = b + 4; s.qd_empty = b + 7; s.s_full = b + 10; s.p_full = b + 12;
    s.ds_free = b + 14; s.dkv_full = b + 16; s.dkv_empty = b + 18; s.dq_full = b + 20; s.dq_empty = b + 21;
    s.aux_full = b + 22; s.aux_empty = b + 24; s.tmem_slot = reinterpret_cast<uint32_t*>(b + 26);
    return s;
}
__device__ __forceinline__ void wait2(uint64_t* bar, uint32_t parity, int tag) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (tc::mbar_try_wait(bar, parity)) return;
    printf("vsw attn2 bwd: mbarrier wait timed out (block %d thread %d, wait site %d, parity %u)\n", blockIdx.x, threadIdx.x, tag, parity);
    __trap();
}
// byte offset of the 16-byte unit holding queries [8u, 8u+8) of key row r inside a 128-byte-swizzled dS chunk
__device__ __forceinline__ uint32_t sw128_unit(int row, int unit) { return row * 128 + ((unit ^ (row & 7)) << 4); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One step = two query rows (16 padded columns) of this thread's key.  All shared-memory loads of the step (statistics, bias
// vectors, histogram entries) are issued back to back BEFORE the arithmetic and every store comes after it: the loads and
// stores are volatile asm (the compiler keeps their order), so a row-by-row formulation serialises load -> math -> store chains
// of ~200 cycles each, which two warps per scheduler cannot hide.
template <int WW, bool MASKED, bool F16>
__device__ __forceinline__ void bwd_step(const uint32_t (&rs)[16], const uint32_t (&rd)[16], uint32_t ld_a, uint32_t tab0, uint32_t tab1,
                                         uint32_t hist0, uint32_t hist1, bool ok0, bool ok1, const uint32_t (&nq)[4], float scale_log2,
                                         uint32_t (&pw)[8], uint32_t (&dw)[8]) {
    float4 st[8];                      // {lse * log2e, delta} pairs of the 16 queries (broadcast loads)
#pragma unroll
    for (int v = 0; v < 8; ++v) st[v] = lds_f4(ld_a + v * 16);
    uint4 bias[2];
    bias[0] = tc::lds_u4(tab0); bias[1] = tc::lds_u4(tab1);
    // d(bias table): the 8 scores of a (key, query row) are 8 consecutive entries of the key's w_j-shifted histogram copy.
    // Query row rho of key row (d_j, h_j) and query row rho + 1 of key row (d_j, h_j + 1) are the SAME histogram row (in two
    // lanes of this warp), so the two rows' read-modify-writes must not overlap: row 1 is loaded after row 0 is stored.
    // (A row that does not exist adds zeros to a spare row behind the histogram: the stores stay unconditional.)
    float4 h0 = lds_f4(hist0), h1 = lds_f4(hist0 + 16);
    float ds[16], pp[16];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
            const int e = r * 8 + sl;
            if (sl >= WW) { pp[e] = 0.f; ds[e] = 0.f; continue; }
            const float4 t = st[e >> 1];
            const float l2 = (e & 1) ? t.z : t.x, dl = (e & 1) ? t.w : t.y;
            const uint32_t w = word_of(bias[r], sl >> 1);
            float x = fmaf(__uint_as_float(rs[e]), scale_log2, (sl & 1) ? bf_hi(w) : bf_lo(w)) - l2;
            // region mismatch: the byte's MSB replicated into the top byte (one PRMT) is -2^127 or +0.0; times 100 log2(e) / 2^127
        // that is the reference's additive -100 (video_swin.py:304-306) in log2 units
        if (MASKED) x = fmaf(__uint_as_float(msb_to_top(nq[e >> 2], e & 3)), MASKC, x);
            const float pe = (r ? ok1 : ok0) ? tc::ex2_approx(x) : 0.f;
            pp[e] = pe;
            ds[e] = pe * (__uint_as_float(rd[e]) - dl);
        }
        h0.x += ds[8 * r + 0]; h0.y += ds[8 * r + 1]; h0.z += ds[8 * r + 2]; h0.w += ds[8 * r + 3];
        h1.x += ds[8 * r + 4]; h1.y += ds[8 * r + 5]; h1.z += ds[8 * r + 6]; h1.w += ds[8 * r + 7];
        if (r == 0) {
            sts_f4(hist0, h0); sts_f4(hist0 + 16, h1);
            h0 = lds_f4(hist1); h1 = lds_f4(hist1 + 16);
        } else {
            sts_f4(hist1, h0); sts_f4(hist1 + 16, h1);
        }
    }
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        pw[e >> 1] = pack16<F16>(pp[e], pp[e + 1]);
        dw[e >> 1] = pack16<F16>(ds[e], ds[e + 1]);
    }
}
__global__ void tc2_dbias_reduce_kernel(const float* __restrict__ part, int groups, int nH, int L, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // out layout (L, nH); fixed summation order over the CTA groups
    if (idx >= L * nH) return;
    const int l = idx / nH, h = idx % nH;
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += part[((long long)g * nH + h) * L + l];
    out[idx] = t;
}

template <int WW, bool F16>
__global__ void __launch_bounds__(B_THREADS, 1)
attn2_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDKV, const BwdParams2 p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const BBars s = bbars_of(base);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.nH * HD;
    const int h = blockIdx.x / p.groups, gi = blockIdx.x % p.groups;   // CTA pinned to one head: table + histograms persist
    const int hist_bytes = p.tab_bytes * 2;                             // fp32 copy of the bf16 table layout
    int* koff = reinterpret_cast<int*>(base + BO_KOFF);
    const uint32_t base_a = tc::smem_u32(base);

    if (warp == BW_TMA && lane == 0) {
        tc::prefetch_tmap(&tmQ); tc::prefetch_tmap(&tmKV); tc::prefetch_tmap(&tmDO); tc::prefetch_tmap(&tmDKV);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.kv_full[i], 1); tc::mbar_init(&s.kv_empty[i], 1);
            tc::mbar_init(&s.s_full[i], 1); tc::mbar_init(&s.p_full[i], 4); tc::mbar_init(&s.ds_free[i], 1);
            tc::mbar_init(&s.dkv_full[i], 1); tc::mbar_init(&s.dkv_empty[i], 4);
            tc::mbar_init(&s.aux_full[i], 2); tc::mbar_init(&s.aux_empty[i], 8);
        }
        for (int i = 0; i < NQD; ++i) { tc::mbar_init(&s.qd_full[i], 1); tc::mbar_init(&s.qd_empty[i], 1); }
        tc::mbar_init(s.dq_full, 1); tc::mbar_init(s.dq_empty, 4);
        tc::fence_barrier_init();
    }
    if (warp == BW_MMA) tc::tmem_alloc(s.tmem_slot, TMEM_COLS);
    // one-time shared-memory state: zero the K / V tiles (the 4 lanes per quadrant no tensor-map box ever writes must stay 0),
    // the dS chunks and the zero chunk between them, the histograms; load the head's table; query-row offsets
    for (int n = threadIdx.x; n < BO_QD / 16; n += B_THREADS) reinterpret_cast<uint4*>(base + BO_KV)[n] = make_uint4(0, 0, 0, 0);
    for (int n = threadIdx.x; n < 3 * 16384 / 16; n += B_THREADS) reinterpret_cast<uint4*>(base + BO_DS)[n] = make_uint4(0, 0, 0, 0);
    for (int n = threadIdx.x; n < 2 * (hist_bytes + 64) / 16; n += B_THREADS) {
        const int c = n / ((hist_bytes + 64) / 16), o = n - c * ((hist_bytes + 64) / 16);
        reinterpret_cast<uint4*>(base + BO_HIST + c * HIST_MAX_BYTES)[o] = make_uint4(0, 0, 0, 0);
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.tabg + (size_t)h * (p.tab_bytes / 2));
        for (int n = threadIdx.x; n < p.tab_bytes / 16; n += B_THREADS) reinterpret_cast<uint4*>(base + BO_TAB)[n] = __ldg(src + n);
    }
    for (int n = threadIdx.x; n < 64; n += B_THREADS) koff[n] = (n / p.wh) * p.NHt + n % p.wh;
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s.tmem_slot;
    const int n_items = (p.B_ - gi + p.groups - 1) / p.groups;   // windows gi, gi + groups, ...
    const int hb_per_item = p.nT * p.nhb;

    if (*p.poison) {
        // codes are not the dense codes of the named window: poison the gradients (NaN) instead of returning wrong numbers
        for (int n = threadIdx.x; n < p.L; n += B_THREADS) p.dbias_part[((long long)gi * p.nH + h) * p.L + n] = __int_as_float(0x7FC00000);
    } else if (warp == BW_TMA) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int kt = 0, c = 0;
            for (int it = 0; it < n_items; ++it) {
                const int b_ = gi + it * p.groups;
                for (int T = 0; T < p.nT; ++T, ++kt) {
                    const int ks = kt & 1;
                    wait2(&s.kv_empty[ks], ((kt >> 1) & 1) ^ 1, 1);
                    tc::mbar_expect_tx(&s.kv_full[ks], 8 * (2 * KTR * HD * 2));
                    uint8_t* kv = base + BO_KV + ks * 16384;
                    for (int which = 0; which < 2; ++which)
                        for (int q = 0; q < 4; ++q)   // quadrant q: w_j = 2q, 2q+1 of the tile's 14 key rows
                            tc::tma_load_4d(&tmKV, &s.kv_full[ks], kv + which * 8192 + q * 2048, (1 + which) * C + h * HD, 2 * q, T * KTR, b_);
                    for (int hb = 0; hb < p.nhb; ++hb, ++c) {
                        const int qs = c % NQD;
                        wait2(&s.qd_empty[qs], ((c / NQD) & 1) ^ 1, 2);
                        tc::mbar_expect_tx(&s.qd_full[qs], 2 * (QHR * SLOT * HD * 2));
                        uint8_t* qd = base + BO_QD + qs * 8192;
                        tc::tma_load_4d(&tmQ, &s.qd_full[qs], qd, h * HD, 0, hb * QHR, b_);
                        tc::tma_load_4d(&tmDO, &s.qd_full[qs], qd + 4096, h * HD, 0, hb * QHR, b_);
                    }
                }
            }
        }
    } else if (warp == BW_MMA) {
        // ===================== MMA issuer (one elected thread) =====================
        if (tc::elect_one()) {
            if (tmem != 0) __trap();   // the whole tensor memory is allocated: base is column 0
            const uint32_t fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t id_base = (1u << 4) | fmt | ((uint32_t)(128 >> 4) << 24);
            const uint32_t id_s = id_base | ((uint32_t)(64 >> 3) << 17);                               // A, B K-major, N = 64
            const uint32_t id_t = id_base | (1u << 16) | ((uint32_t)(HD >> 3) << 17);                  // A from TMEM, B MN-major, N = 32
            const uint32_t id_q = id_base | (1u << 15) | (1u << 16) | ((uint32_t)(HD >> 3) << 17);     // A MN-major (dS tile), B MN-major
            // Software pipeline with a lag of one half-block (across tiles and items): S^T / dP^T of half-block c are issued, then
            // the dV / dK / dQ MMAs of half-block c - 1, whose exp pass ran meanwhile.  One consume site, straight-line code.
            const int total = n_items * hb_per_item;
            int i_it = 0, i_T = 0, i_hb = 0, i_kt = 0;          // issue cursor (half-block c)
            int q_it = 0, q_T = 0, q_hb = 0, q_kt = 0;          // half-block c - 1
            int p_it = 0, p_T = 0, p_hb = 0, p_kt = 0;          // half-block c - 2: the one consumed in this iteration
            const bool mprof = VSW_ATTN2_PROF && p.dbg && blockIdx.x == 0;
            long long m_wp = 0, m_wl = 0, m_wk = 0, m_we = 0, m_n = 0;
            const long long m_t0 = mprof ? clock64() : 0;
            // Order per iteration c:  dV / dK of half-block c - 2 (they read P^T / dS^T from stage c & 1)  ->  S^T / dP^T of half-block
            // c into that stage  ->  dQ of half-block c - 2 (shared memory operands only).  The exp group that finished c - 2 gets
            // its next scores after 12 MMAs instead of 20; the tensor pipe runs in order, so no barrier is needed between them.
            for (int c = 0; c <= total + 1; ++c) {
                const int cp = c - 2;
                const int pst = cp & 1, pqs = (cp + NQD) % NQD, pks = p_kt & 1, par = p_kt & 1;   // (unused while c < 2)
                if (c >= 2) {
                    const long long t_e = mprof ? clock64() : 0;
                    if (p_hb == 0 && p_kt >= 2) wait2(&s.dkv_empty[par], ((p_kt >> 1) - 1) & 1, 5);   // dK / dV buffer read out (tile kt - 2)
                    long long t_b = 0;
                    if (mprof) { t_b = clock64(); m_we += t_b - t_e; }
                    wait2(&s.p_full[pst], (cp >> 1) & 1, 3);
                    tc::tc_fence_after();
                    if (mprof) { m_wp += clock64() - t_b; m_n += 1; }
                    const uint32_t qda = base_a + BO_QD + pqs * 8192;
                    const uint32_t acc_kv = p_hb > 0 ? 1u : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // dV_T (+)= P^T dO_hb : 16 queries per MMA
                        tc::umma_bf16_ts(TB_DKV + 64 * par + 32, pst * TB_STAGE + TB_ST + 8 * k, tc::smem_desc_sw64(qda + 4096 + k * 1024, 0, 512), id_t, k ? 1u : acc_kv);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // dK_T (+)= dS^T Q_hb
                        tc::umma_bf16_ts(TB_DKV + 64 * par, pst * TB_STAGE + TB_DP + 8 * k, tc::smem_desc_sw64(qda + k * 1024, 0, 512), id_t, k ? 1u : acc_kv);
                    tc::umma_commit(&s.qd_empty[pqs]);
                }
                if (c < total) {
                    const int ks = i_kt & 1, st = c & 1, qs = c % NQD;
                    long long t_a = 0;
                    if (mprof) t_a = clock64();
                    if (i_hb == 0) wait2(&s.kv_full[ks], (i_kt >> 1) & 1, 4);
                    if (mprof) { const long long t_k = clock64(); m_wk += t_k - t_a; t_a = t_k; }
                    wait2(&s.qd_full[qs], (c / NQD) & 1, 7);
                    tc::tc_fence_after();
                    if (mprof) m_wl += clock64() - t_a;
                    const uint32_t kva = base_a + BO_KV + ks * 16384, qda = base_a + BO_QD + qs * 8192;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc::umma_bf16(st * TB_STAGE + TB_ST, tc::smem_desc_sw64(kva + k * 32, 0, 512), tc::smem_desc_sw64(qda + k * 32, 0, 512), id_s, k);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc::umma_bf16(st * TB_STAGE + TB_DP, tc::smem_desc_sw64(kva + 8192 + k * 32, 0, 512), tc::smem_desc_sw64(qda + 4096 + k * 32, 0, 512), id_s, k);
                    tc::umma_commit(&s.s_full[st]);
                }
                if (c >= 2) {
                    const long long t_e = mprof ? clock64() : 0;
                    if (p_T == 0 && p_hb == 0 && p_it > 0) wait2(s.dq_empty, (p_it - 1) & 1, 6);       // dQ of the previous item read out
                    if (mprof) m_we += clock64() - t_e;
                    // dQ_qb (+)= dS K_T with M = 128 query rows of which this half-block fills 64: the other 64-row chunk of the
                    // MN-major A operand is the zero chunk (even half-block: data | zeros, odd: zeros | data)
                    const uint32_t kva = base_a + BO_KV + pks * 16384;
                    const uint32_t dsa = base_a + BO_DS + ((p_hb & 1) ? 16384 : 0);
                    const uint32_t acc_q = (p_T == 0 && (p_hb & 1) == 0) ? 0u : 1u;
                    const int qb = p_hb >> 1;
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // 16 keys per MMA
                        tc::umma_bf16(TB_DQ + 32 * qb, tc::smem_desc_sw128(dsa + k * 2048, 16384, 1024), tc::smem_desc_sw64(kva + k * 1024, 0, 512), id_q, k ? 1u : acc_q);
                    tc::umma_commit(&s.ds_free[p_hb & 1]);
                    if (p_hb == p.nhb - 1) {
                        tc::umma_commit(&s.dkv_full[par]);
                        if (p_T == p.nT - 1) tc::umma_commit(s.dq_full);
                    }
                }
                // shift the cursors: c - 1 becomes c - 2, c becomes c - 1; advance the issue cursor
                p_it = q_it; p_T = q_T; p_hb = q_hb; p_kt = q_kt;
                q_it = i_it; q_T = i_T; q_hb = i_hb; q_kt = i_kt;
                if (++i_hb == p.nhb) { i_hb = 0; ++i_kt; if (++i_T == p.nT) { i_T = 0; ++i_it; } }
            }
            if (mprof) { p.dbg[8] += m_wp; p.dbg[9] += m_wl; p.dbg[10] += m_we; p.dbg[13] += m_wk; p.dbg[12] += m_n; p.dbg[20] += clock64() - m_t0; }
        }
    } else if (warp == BW_AUX0 || warp == BW_AUX1) {
        // ===================== aux warps: {lse * log2e, delta} per query; region ids =====================
        for (int it = 0; it < n_items; ++it) {
            const int st = it & 1;
            const int b_ = gi + it * p.groups, win = b_ % p.nW;
            wait2(&s.aux_empty[st], ((it >> 1) & 1) ^ 1, 8);
            if (warp == BW_AUX0) {
                float2* ld = reinterpret_cast<float2*>(base + BO_LD + st * MAXCOLS * 8);
                for (int n = lane; n < MAXCOLS; n += 32) {
                    const int row = n >> 3, sl = n & 7;
                    float2 v = make_float2(0.f, 0.f);
                    if (sl < p.ww && row < p.KR) {
                        const int i = row * p.ww + sl;
                        const uint4* op = reinterpret_cast<const uint4*>((const uint16_t*)p.out + ((long long)b_ * p.N + i) * C + h * HD);
                        const uint4* dp = reinterpret_cast<const uint4*>((const uint16_t*)p.dout + ((long long)b_ * p.N + i) * C + h * HD);
                        float d = 0.f;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint4 a = __ldg(op + u), b = __ldg(dp + u);
                            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 x, y;
                                if (F16) { x = __half22float2(*reinterpret_cast<const __half2*>(&aw[e])); y = __half22float2(*reinterpret_cast<const __half2*>(&bw[e])); }
                                else { x = tc::unpack_bf16(aw[e]); y = tc::unpack_bf16(bw[e]); }
                                d = fmaf(x.x, y.x, d); d = fmaf(x.y, y.y, d);
                            }
                        }
                        v = make_float2(p.lse[((long long)b_ * p.nH + h) * p.N + i] * LOG2E, d);
                    }
                    ld[n] = v;
                }
            } else {
                uint8_t* regq = base + BO_REGQ + st * MAXCOLS;
                uint8_t* regk = base + BO_REGK + st * 512;
                int diff = 0;
                if (p.region) {
                    const uint8_t* rg = p.region + (long long)win * p.N;
                    const uint8_t r0 = rg[0];
                    for (int n = lane; n < MAXCOLS; n += 32) {
                        const int row = n >> 3, sl = n & 7;
                        const bool real = sl < p.ww && row < p.KR;
                        const uint8_t r = real ? rg[row * p.ww + sl] : (uint8_t)0xFF;
                        regq[n] = r;
                        diff |= real && r != r0;
                    }
                    for (int n = lane; n < 512; n += 32) {   // key lane n of tile n / 128: quadrant (n % 128) / 32, lane l -> (row 14 T + l / 2, w 2 q + l % 2)
                        const int T = n >> 7, q = (n >> 5) & 3, l = n & 31;
                        const int row = T * KTR + (l >> 1), w = 2 * q + (l & 1);
                        regk[n] = (l < 2 * KTR && row < p.KR && w < p.ww) ? rg[row * p.ww + w] : (uint8_t)0xFE;
                    }
                }
                diff = __any_sync(0xffffffffu, diff);
                if (lane == 0) reinterpret_cast<int*>(base + BO_FLAG)[st] = diff;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_full[st]);
        }
    } else if (warp >= BW_EPI) {
        // ===================== epilogue warps: dK / dV per key tile, dQ per item =====================
        // (On the exp warps these read-outs -- waiting for the tile's last MMA, the TMEM loads, packing, the stores -- cost ~11 %
        // of a group's cycle and drained the pipeline at every item end; here they overlap the next tile's / item's scores.)
        const int q = warp - BW_EPI;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int krow = q * 32 + lane;                     // row of this thread's key in the K / V tiles
        int kt = 0;
        for (int it = 0; it < n_items; ++it) {
            const int b_ = gi + it * p.groups;
            for (int T = 0; T < p.nT; ++T, ++kt) {
                // dK / dV of the tile: TMEM -> registers -> (scaled, packed) into the tile's own K / V shared-memory buffer, whose
                // rows are in TMEM lane order -> tensor-map stores with the boxes the tile was loaded with (window-row padding and
                // rows beyond the window are clipped by the map; per-lane 16-byte global stores of 64-byte rows would cost 32 LSU
                // wavefronts per instruction on the pipe the exp warps are bound by)
                const int par = kt & 1;
                wait2(&s.dkv_full[par], (kt >> 1) & 1, 9);
                tc::tc_fence_after();
                uint32_t dk[32], dv[32];
                tc::tmem_ld_32x32(tmem + lane_base + TB_DKV + 64 * par, dk);
                tc::tmem_ld_32x32(tmem + lane_base + TB_DKV + 64 * par + 32, dv);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s.dkv_empty[par]);
                const uint32_t kva = base_a + BO_KV + par * 16384 + krow * 64;
                if (lane < 2 * KTR) {                      // rows 28..31 of a quadrant stay zero (no box ever covers them)
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        uint4 u, w;
                        u.x = pack16<F16>(__uint_as_float(dk[8 * v4 + 0]) * p.scale, __uint_as_float(dk[8 * v4 + 1]) * p.scale);
                        u.y = pack16<F16>(__uint_as_float(dk[8 * v4 + 2]) * p.scale, __uint_as_float(dk[8 * v4 + 3]) * p.scale);
                        u.z = pack16<F16>(__uint_as_float(dk[8 * v4 + 4]) * p.scale, __uint_as_float(dk[8 * v4 + 5]) * p.scale);
                        u.w = pack16<F16>(__uint_as_float(dk[8 * v4 + 6]) * p.scale, __uint_as_float(dk[8 * v4 + 7]) * p.scale);
                        w.x = pack16<F16>(__uint_as_float(dv[8 * v4 + 0]), __uint_as_float(dv[8 * v4 + 1]));
                        w.y = pack16<F16>(__uint_as_float(dv[8 * v4 + 2]), __uint_as_float(dv[8 * v4 + 3]));
                        w.z = pack16<F16>(__uint_as_float(dv[8 * v4 + 4]), __uint_as_float(dv[8 * v4 + 5]));
                        w.w = pack16<F16>(__uint_as_float(dv[8 * v4 + 6]), __uint_as_float(dv[8 * v4 + 7]));
                        const uint32_t off = (uint32_t)((v4 ^ ((krow >> 1) & 3)) << 4);     // 64-byte swizzle of the tile
                        tc::sts_u4(kva + off, u);
                        tc::sts_u4(kva + 8192 + off, w);
                    }
                }
                tc::fence_proxy_async();
                tc::named_bar_sync(1, 128);                // the four epilogue warps
                if (q == 0 && lane == 0) {
                    const uint8_t* kv = base + BO_KV + par * 16384;
#pragma unroll
                    for (int which = 0; which < 2; ++which)
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq)
                            tc::tma_store_4d(&tmDKV, kv + which * 8192 + qq * 2048, (1 + which) * C + h * HD, 2 * qq, T * KTR, b_);
                    tc::bulk_commit_group();
                    tc::bulk_wait_read_all();              // the stores have read the buffer: it goes back to the loader
                    tc::mbar_arrive(&s.kv_empty[par]);
                }
            }
            // ---- dQ of the item: all of its MMAs have retired
            wait2(s.dq_full, it & 1, 13);
            tc::tc_fence_after();
            for (int qb = 0; qb < (p.nhb + 1) / 2; ++qb) {
                uint32_t o[32];
                tc::tmem_ld_32x32(tmem + lane_base + TB_DQ + 32 * qb, o);
                tc::tmem_ld_wait();
                const int col = qb * 128 + q * 32 + lane, row = col >> 3, sl = col & 7;
                if (sl < p.ww && row < p.KR) {
                    uint4* dst = reinterpret_cast<uint4*>((uint16_t*)p.dqkv + (((long long)b_ * p.N + row * p.ww + sl) * 3) * C + h * HD);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4) {
                        uint4 u;
                        u.x = pack16<F16>(__uint_as_float(o[8 * v4 + 0]) * p.scale, __uint_as_float(o[8 * v4 + 1]) * p.scale);
                        u.y = pack16<F16>(__uint_as_float(o[8 * v4 + 2]) * p.scale, __uint_as_float(o[8 * v4 + 3]) * p.scale);
                        u.z = pack16<F16>(__uint_as_float(o[8 * v4 + 4]) * p.scale, __uint_as_float(o[8 * v4 + 5]) * p.scale);
                        u.w = pack16<F16>(__uint_as_float(o[8 * v4 + 6]) * p.scale, __uint_as_float(o[8 * v4 + 7]) * p.scale);
                        dst[v4] = u;
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(s.dq_empty);
        }
        if (q == 0 && lane == 0) tc::bulk_wait_all();      // this thread's dK / dV tile stores
    } else {
        // ===================== exp / dS warps: thread = key =====================
        const int q = warp & 3, g = warp >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int wj = 2 * q + (lane & 1);                 // this thread's key column w_j (7 = dummy for a 7-wide window)
        const int rl = lane >> 1;                           // key row inside the tile (0..13; 14, 15: unused lanes)
        const bool lane_real = lane < 2 * KTR && wj < p.ww;
        // histogram rows of odd w_j sit 16 bytes further: the two key columns of a quarter-warp then hit different banks
        const uint32_t tab_a = base_a + BO_TAB, hist_a = base_a + BO_HIST + g * HIST_MAX_BYTES + (lane & 1) * 16;
        const int krow = q * 32 + lane;                     // row of this key in the K / V / dS tiles
        const int dummy_row = p.tab_bytes / 16;             // spare row behind the table and behind each histogram copy
        int c = 0, kt = 0;
        long long x_ws = 0, x_wd = 0, x_work = 0, x_n = 0, x_epi = 0, x_nepi = 0, x_dq = 0, x_aux = 0, x_last = 0, x_g0 = 0, x_g1 = 0, x_g2 = 0, x_g3 = 0;
        const bool xprof = VSW_ATTN2_PROF && p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0;
        const long long x_t0 = xprof ? clock64() : 0;
        int use_ds0 = 0, use_ds1 = 0;                        // uses so far of the even / odd dS chunk (all half-blocks, owned or not; scalars: a
                                                             // dynamically indexed array would live in local memory, ~700 cycles per access here)
        // (the integer divisions are ~1000 cycles of dependent instructions: once per kernel for the first four key tiles)
        auto rowbase_of = [&](int T) {
            const int dj = (T * KTR + rl) / p.wh, hj = (T * KTR + rl) % p.wh;
            return wj * p.blk + (p.wdc - 1 - dj) * p.NHt - hj + p.wh - 1;
        };
        const int rb0 = rowbase_of(0), rb1 = rowbase_of(1), rb2 = rowbase_of(2), rb3 = rowbase_of(3);
        for (int it = 0; it < n_items; ++it) {
            const int ast = it & 1;
            const int b_ = gi + it * p.groups;
            const long long x_a0 = xprof ? clock64() : 0;
            wait2(&s.aux_full[ast], (it >> 1) & 1, 10);
            if (xprof) x_aux += clock64() - x_a0;
            const bool masked = reinterpret_cast<const int*>(base + BO_FLAG)[ast] != 0;
            const uint32_t ld_base = base_a + BO_LD + ast * MAXCOLS * 8;
            const uint32_t regq_a = base_a + BO_REGQ + ast * MAXCOLS;
            const uint8_t* regk = base + BO_REGK + ast * 512;
            for (int T = 0; T < p.nT; ++T, ++kt) {
                const bool key_real = lane_real && T * KTR + rl < p.KR;
                // table / histogram row of (this key, query row rho) = rowbase + koff[rho]
                const int rowbase = T == 0 ? rb0 : T == 1 ? rb1 : T == 2 ? rb2 : T == 3 ? rb3 : rowbase_of(T);
                const uint32_t regj4 = masked ? (uint32_t)regk[T * 128 + krow] * 0x01010101u : 0u;
                for (int hb = 0; hb < p.nhb; ++hb, ++c) {
                    const int ds_use = (hb & 1) ? use_ds1++ : use_ds0++;
                    if ((c & 1) != g) continue;
                    const int st = c & 1;
                    if (xprof && x_last && hb >= 2) x_g3 += clock64() - x_last;
                    long long x_a = 0, x_b = 0, x_c = 0;
                    if (xprof) {
                        x_a = clock64();
                        if (x_last) { if (hb >= 2) x_g0 += x_a - x_last; else if (T > 0) x_g1 += x_a - x_last; else x_g2 += x_a - x_last; }
                    }
                    wait2(&s.s_full[st], (c >> 1) & 1, 11);
                    if (xprof) x_b = clock64();
                    tc::tc_fence_after();
                    if (xprof) x_c = clock64();
                    const uint32_t trow = tmem + lane_base + st * TB_STAGE;
                    const uint32_t ds_a = base_a + BO_DS + ((hb & 1) ? 32768 : 0);
                    const int nrows = min(QHR, p.KR - hb * QHR);      // query rows of this half-block that exist
                    // 4 steps of 16 query columns (2 query rows), fully unrolled; the TMEM loads of step k + 1 are in flight while
                    // step k is computed (two warps per scheduler cannot hide a load-use latency per step)
                    auto steps = [&](auto masked_tag) {
                        constexpr bool MK = decltype(masked_tag)::value;
                        uint32_t rs[2][16], rd[2][16];
                        uint32_t dwk[8];                 // dS of step 0, stored with step 1's (below)
                        tc::tmem_ld_32x16(trow + TB_ST, rs[0]);
                        tc::tmem_ld_32x16(trow + TB_DP, rd[0]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int cur = k & 1;
                            if (k + 1 < 4) {
                                tc::tmem_ld_32x16(trow + TB_ST + 16 * (k + 1), rs[cur ^ 1]);
                                tc::tmem_ld_32x16(trow + TB_DP + 16 * (k + 1), rd[cur ^ 1]);
                            }
                            uint32_t pw[8], dw[8];
                            uint32_t nq[4] = {0, 0, 0, 0};
                            if (MK) {
                                const uint4 rg = tc::lds_u4(regq_a + hb * 64 + k * 16);
                                nq[0] = ne_msb4(rg.x, regj4); nq[1] = ne_msb4(rg.y, regj4);   // byte MSB set <=> the region ids differ
                                nq[2] = ne_msb4(rg.z, regj4); nq[3] = ne_msb4(rg.w, regj4);
                            }
                            {
                                const int rho = hb * QHR + 2 * k;
                                const bool ok0 = 2 * k < nrows && key_real, ok1 = 2 * k + 1 < nrows && key_real;
                                const int2 ko = *reinterpret_cast<const int2*>(koff + rho);
                                const int ri0 = ok0 ? rowbase + ko.x : dummy_row, ri1 = ok1 ? rowbase + ko.y : dummy_row;   // a spare row behind the table / histogram stands in for a missing row
                                bwd_step<WW, MK, F16>(rs[cur], rd[cur], ld_base + rho * 64, tab_a + ri0 * 16, tab_a + ri1 * 16, hist_a + ri0 * 32,
                                                      hist_a + ri1 * 32, ok0, ok1, nq, p.scale_log2, pw, dw);
                            }
                            tc::tmem_st_32x8(trow + TB_ST + 8 * k, pw);      // P^T  (A operand of dV)
                            tc::tmem_st_32x8(trow + TB_DP + 8 * k, dw);      // dS^T (A operand of dK)
                            // dS^T row of this key into the MN-major A tile of dQ = dS K: 16 queries = two 16-byte units.  The tile's
                            // previous contents are still being multiplied (dQ of half-block c - 2 is issued after this half-block's
                            // scores): step 0 keeps its dS in registers and the wait sits behind step 1's arithmetic.
                            if (k == 0) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) dwk[e] = dw[e];
                            } else {
                                if (k == 1) {
                                    const long long x_w0 = xprof ? clock64() : 0;
                                    wait2(&s.ds_free[hb & 1], (ds_use & 1) ^ 1, 12);
                                    if (xprof) x_wd += clock64() - x_w0;
                                    tc::sts_u4(ds_a + sw128_unit(krow, 0), make_uint4(dwk[0], dwk[1], dwk[2], dwk[3]));
                                    tc::sts_u4(ds_a + sw128_unit(krow, 1), make_uint4(dwk[4], dwk[5], dwk[6], dwk[7]));
                                }
                                tc::sts_u4(ds_a + sw128_unit(krow, 2 * k), make_uint4(dw[0], dw[1], dw[2], dw[3]));
                                tc::sts_u4(ds_a + sw128_unit(krow, 2 * k + 1), make_uint4(dw[4], dw[5], dw[6], dw[7]));
                            }
                            if (k + 1 < 4) tc::tmem_ld_wait();
                        }
                    };
                    if (masked) steps(std::true_type{}); else steps(std::false_type{});
                    tc::tmem_st_wait();
                    tc::fence_proxy_async();   // the dS tile is read by the tensor core
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&s.p_full[st]);
                    if (xprof) { x_last = clock64(); x_ws += x_b - x_a; x_wd += x_c - x_b; x_work += x_last - x_c; x_n += 1; }
                }
            }
            // the item's side data (lse / delta, region ids) may be replaced
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_empty[ast]);
        }
        if (xprof) { p.dbg[0] += x_ws; p.dbg[1] += x_work; p.dbg[2] += x_n; p.dbg[3] += x_wd; p.dbg[6] += clock64() - x_t0; p.dbg[4] += x_nepi; p.dbg[5] += x_epi; p.dbg[7] += x_dq; p.dbg[14] += x_aux; p.dbg[15] += x_g0; p.dbg[17] += x_g1; p.dbg[18] += x_g2; p.dbg[19] += x_g3; }
    }

    __syncwarp();
    tc::tc_fence_before();
    __syncthreads();
    // fold the two histogram copies and the per-w_j shifted layout back into the table column (fixed order)
    if (!*p.poison) {
        const int NW = 2 * p.ww - 1;
        float* part = p.dbias_part + ((long long)gi * p.nH + h) * p.L;
        for (int l = threadIdx.x; l < p.L; l += B_THREADS) {
            const int dw = l % NW, r = l / NW;              // r = dd * NHt + dh
            float t = 0.f;
            for (int wj = 0; wj < p.ww; ++wj) {
                const int wi = dw - (p.ww - 1) + wj;        // table index dw = w_i - w_j + ww - 1
                if (wi < 0 || wi >= p.ww) continue;
                const int idx = (wj * p.blk + r) * SLOT + wi + (wj & 1) * 4;
                t += reinterpret_cast<const float*>(base + BO_HIST)[idx];
                t += reinterpret_cast<const float*>(base + BO_HIST + HIST_MAX_BYTES)[idx];
            }
            part[l] = t;
        }
    }
    if (warp == BW_MMA) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

// table rows per w_j block, padded so that consecutive blocks start 4 rows apart mod 8 (bank-conflict-free quarter-warps)
int bwd2_blk(const Geometry& g) {
    int blk = (2 * g.wdc - 1) * (2 * g.wh - 1);
    while (blk % 8 != 4) ++blk;
    return blk;
}
int bwd2_tab_bytes(const Geometry& g) { return g.ww * bwd2_blk(g) * SLOT * 2; }

int bwd2_groups(int B_, int nH, int sms) {
    int g = sms / nH;
    if (g < 1) g = 1;
    if (g > B_) g = B_;
    return g;
}

}  // namespace

size_t tc2_attn_bwd_workspace(int B_, int N, int nH, int hd, int L) {
    (void)N; (void)hd;
    if (nH > kNumSMs) return 0;
    return (size_t)bwd2_groups(B_, nH, kNumSMs) * nH * L * sizeof(float);
}

int tc2_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                 const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, void* dqkv, float* dbias,
                 int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims, int dtype, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
    Geometry g;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > kNumSMs) sms = kNumSMs;   // the workspace size is computed for at most kNumSMs CTAs
    if ((dtype != VSW_BF16 && dtype != VSW_F16) || !geometry_of(N, hd, L, window_dims, &g) || bwd2_tab_bytes(g) + 32 > BTAB_MAX_BYTES || nH > sms || !aligned16(qkv) ||
        !aligned16(out) || !aligned16(dout) || !aligned16(dqkv)) {
        set_error("tcgen05 window attention bwd: needs bf16/fp16, head_dim 32, the configured window (rows of <= 8 tokens) as "
                  "layout hint and N <= 448 a whole number of window rows (hd=%d N=%d L=%d window_dims=0x%x)", hd, N, L, window_dims);
        return VSW_ERR_UNSUPPORTED;
    }
    const int groups = bwd2_groups(B_, nH, sms);
    if (ws_bytes < (size_t)groups * nH * L * sizeof(float)) { set_error("tcgen05 attention bwd: workspace too small"); return VSW_ERR_WORKSPACE; }
    const int C = nH * HD;
    CUtensorMap tmQ, tmKV, tmDO, tmDKV;
    {
        const uint64_t dims[4] = {(uint64_t)3 * C, (uint64_t)g.ww, (uint64_t)g.KR, (uint64_t)B_};
        const uint64_t strides[4] = {1, (uint64_t)3 * C, (uint64_t)g.ww * 3 * C, (uint64_t)N * 3 * C};
        const uint32_t boxq[4] = {HD, SLOT, QHR, 1}, boxk[4] = {HD, 2, KTR, 1};
        constexpr int promo = 64;   // see make_qkv_maps: 64-byte head slices, no over-fetch
        if (!make_tmap_nd_bf16(&tmQ, qkv, 4, dims, strides, boxq, 64, promo) || !make_tmap_nd_bf16(&tmKV, qkv, 4, dims, strides, boxk, 64, promo) ||
            !make_tmap_nd_bf16(&tmDKV, dqkv, 4, dims, strides, boxk, 64)) return VSW_ERR_CUDA;
        const uint64_t dimo[4] = {(uint64_t)C, (uint64_t)g.ww, (uint64_t)g.KR, (uint64_t)B_};
        const uint64_t strido[4] = {1, (uint64_t)C, (uint64_t)g.ww * C, (uint64_t)N * C};
        if (!make_tmap_nd_bf16(&tmDO, dout, 4, dimo, strido, boxq, 64, promo)) return VSW_ERR_CUDA;
    }
    const int blk = bwd2_blk(g), tabb = bwd2_tab_bytes(g);
    const size_t tab_total = (size_t)nH * tabb;
    uint8_t* scratch = scratch_for(st, tab_total + (size_t)nH * 8 + 16);
    if (!scratch) return VSW_ERR_CUDA;
    uint16_t* tabg = (uint16_t*)scratch;
    float* tabstat = (float*)(scratch + tab_total);
    int* poison = (int*)(scratch + tab_total + (size_t)nH * 8);
    cudaMemsetAsync(poison, 0, 4, st);
    if (dtype == VSW_BF16)
        attn2_table_kernel<__nv_bfloat16><<<nH, 1024, 0, st>>>((const __nv_bfloat16*)table, rowcode, colcode, N, nH, L, g.wdc, g.wh, g.ww, 1, blk, tabg, tabstat, poison);
    else
        attn2_table_kernel<__half><<<nH, 1024, 0, st>>>((const __half*)table, rowcode, colcode, N, nH, L, g.wdc, g.wh, g.ww, 1, blk, tabg, tabstat, poison);
    int rc = check_launch("attn2_table");
    if (rc) return rc;
    BwdParams2 p{};
    p.tabg = tabg; p.tabstat = tabstat; p.poison = poison; p.region = region;
    p.out = out; p.dout = dout; p.lse = lse; p.dqkv = dqkv; p.dbias_part = (float*)ws;
    p.B_ = B_; p.nW = nW; p.N = N; p.nH = nH; p.wh = g.wh; p.ww = g.ww; p.KR = g.KR;
    p.nT = (g.KR + KTR - 1) / KTR; p.nhb = (g.KR + QHR - 1) / QHR;
    p.tab_bytes = tabb; p.blk = blk; p.NHt = 2 * g.wh - 1; p.wdc = g.wdc; p.L = L; p.groups = groups;
    p.scale = scale; p.scale_log2 = scale * LOG2E;
    p.dbg = attn2_debug_buffer("bwd");
    cudaError_t e = cudaSuccess;
#define VSW_LAUNCH_BWD2(WWV, F16V)                                                                                        \
    do {                                                                                                                  \
        auto kern = attn2_bwd_kernel<WWV, F16V>;                                                                          \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);                            \
        if (e == cudaSuccess) kern<<<groups * nH, B_THREADS, BWD_SMEM, st>>>(tmQ, tmKV, tmDO, tmDKV, p);                         \
    } while (0)
    if (dtype == VSW_BF16) { if (g.ww == 7) VSW_LAUNCH_BWD2(7, false); else VSW_LAUNCH_BWD2(8, false); }
    else { if (g.ww == 7) VSW_LAUNCH_BWD2(7, true); else VSW_LAUNCH_BWD2(8, true); }
#undef VSW_LAUNCH_BWD2
    if (e != cudaSuccess) { set_error("attn bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
    rc = check_launch("attn2_bwd");
    if (rc) return rc;
    tc2_dbias_reduce_kernel<<<ceil_div((long long)L * nH, 256), 256, 0, st>>>((const float*)ws, groups, nH, L, dbias);
    return check_launch("attn2_dbias_reduce");
}

}  // namespace vsw
