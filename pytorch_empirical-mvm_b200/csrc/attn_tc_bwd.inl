// tcgen05 / TMEM recompute-based backward of fused window attention (included by attn_tc.cu).
// Formulas: SURVEY A5.  Per (window, head) item and per (key block kb of 128, query tile t of 128):
//   S  = Q_t K_kb^T,  dP = dO_t V_kb^T                 tcgen05.mma SS -> TMEM (128 x 128 fp32 each)
//   P  = exp2(S*scale*log2e + bias*log2e [+mask] - lse*log2e),  dS = P (dP - delta)     8 softmax warps
//        P, dS -> bf16 tiles in shared memory (manual 128B swizzle), d(bias table) into PRIVATE per-warp
//        shared-memory histograms (lanes of a warp hit distinct entries -> plain read-modify-write, deterministic)
//   dV_kb += P^T dO_t,  dK_kb += dS^T Q_t               A = P/dS (MN-major), B = dO_t/Q_t (MN-major)
//   dQ_t  += dS K_kb                                    A = dS (K-major),    B = K_kb (MN-major)
// dK/dV of the key block and dQ of all four query tiles stay in TMEM until complete (kb outer, t inner).
// A short last query tile (N % 128 <= 8, e.g. 392 = 3 x 128 + 8) would cost a full softmax pass for 8 useful rows, so it
// is processed TRANSPOSED: S^T = K_kb Q_tail^T and dP^T = V_kb dO_tail^T (128 keys x 16), thread = key, 8 queries each;
// P^T / dS^T go to small un-swizzled K-major tiles (dV += P^T dO_tail, dK += dS^T Q_tail with K = 16), and dS is also
// scattered into rows 0..7 of the regular dS tile for dQ_tail += dS K_kb.
// A CTA is pinned to ONE head so its bias-table column and histograms persist across all its windows; the
// per-CTA histograms are summed by a fixed-order second pass.
namespace vsw {
namespace {

constexpr int BW_KV_OFF = 0;                       // K_kb 8 KB, V_kb 8 KB
constexpr int BW_QD_OFF = 16384;                   // 2 stages x (Q_t 8 KB + dO_t 8 KB)
constexpr int BW_P_OFF = 49152;                    // P  tile: 2 chunks x (128 rows x 128 B)
constexpr int BW_DS_OFF = 81920;                   // dS tile
constexpr int BW_MISC_OFF = 114688;
constexpr int BW_S = 0, BW_DP = 128, BW_DK = 256, BW_DV = 288, BW_DQ = 320;  // TMEM columns
// transposed tail tile: P^T and dS^T (128 keys x 8 queries, 16 B per key) live inside the (then unused) P tile region,
// each followed 4 KB later by a 2 KB zero chunk standing in for queries 8..15 of the K = 16 MMA
constexpr int BW_TAIL_PT = 0, BW_TAIL_DST = 2048, BW_TAIL_LBO = 4096;

struct BwdParams {
    const __nv_bfloat16* table; const int32_t* rowcode; const int32_t* colcode; const uint8_t* region;
    const __nv_bfloat16* out; const __nv_bfloat16* dout; const float* lse;
    __nv_bfloat16* dqkv; float* dbias_part;
    int B_, nW, N, nH, L, Lpad, groups;
    float scale, scale_log2;
    int Npad, nq, nkb;
    int ntail;        // > 0: the last query tile holds only ntail (<= 8) rows and is processed transposed (see below)
    long long* dbg;
    int wd, hw, boxhw;   // key permutation: column c' = hw_index * wd + plane (plane = temporal slab of the window)
    int wh, ww, R1, R2, pad;   // padded on-chip bias / histogram layout (pad == 0: dense); see bias_layout()
};

struct BwdSmem {
    float* tab; float* hist;          // [Lpad], [8][Lpad]
    float* delta; float* lse2;        // [512] each (the two column-half warps of a row write identical values)
    int* rc; int* cc;                 // [512]
    uint8_t* reg[2]; int* masked;     // region ids per item in COLUMN order (double-buffered)
    uint8_t* regq[2];                 // region ids per item in QUERY (natural) order
    int* batched;                     // 1 if the wd keys of one spatial position never share a histogram entry within a warp
    uint64_t *kv_full, *kv_empty, *qd_full, *qd_empty, *s_full, *pds_full, *dkv_full, *dq_full, *aux_full, *aux_empty;
    uint32_t* tmem_slot;
};
__device__ __forceinline__ BwdSmem bwd_carve(uint8_t* base, int Lpad) {
    BwdSmem s;
    uint8_t* p = base + BW_MISC_OFF;
    s.tab = (float*)p; p += (size_t)Lpad * 4;
    s.hist = (float*)p; p += (size_t)8 * Lpad * 4;
    s.delta = (float*)p; p += 512 * 4;
    s.lse2 = (float*)p; p += 512 * 4;
    s.rc = (int*)p; p += 512 * 4;
    s.cc = (int*)p; p += 512 * 4;
    s.kv_full = (uint64_t*)p; p += 8;
    s.kv_empty = (uint64_t*)p; p += 8;
    s.qd_full = (uint64_t*)p; p += 16;
    s.qd_empty = (uint64_t*)p; p += 16;
    s.s_full = (uint64_t*)p; p += 8;
    s.pds_full = (uint64_t*)p; p += 8;
    s.dkv_full = (uint64_t*)p; p += 8;
    s.dq_full = (uint64_t*)p; p += 8;
    s.aux_full = (uint64_t*)p; p += 16;
    s.aux_empty = (uint64_t*)p; p += 16;
    s.masked = (int*)p; p += 8;
    s.tmem_slot = (uint32_t*)p; p += 8;
    s.reg[0] = p; p += 512;
    s.reg[1] = p; p += 512;
    s.regq[0] = p; p += 512;
    s.regq[1] = p; p += 512;
    s.batched = (int*)p; p += 16;
    return s;
}
size_t bwd_smem_bytes(int Lpad) {
    return 1024 + BW_MISC_OFF + (size_t)9 * Lpad * 4 + 2 * 512 * 4 + 2 * 512 * 4 + 15 * 8 + 16 + 2048 + 16 + 64;
}

// byte offset of the 16-byte unit holding keys [8u, 8u+8) of query row i inside a 128B-swizzled chunk
__device__ __forceinline__ uint32_t sw128_off(int row, int unit) { return row * 128 + ((unit ^ (row & 7)) << 4); }

// ---- backward softmax / dS over the full 16-column chunks [cbeg, cfull) of one query row ------------------------------
// MASKED and the histogram batch size G are compile-time so that the unmasked, plane-batched common case carries no
// predicated mask arithmetic and no alternative histogram code.
struct BwdCols {
    uint32_t trow, cc_a, reg_a, tabrow, histrow, pt_a, ds_a, regi4;
    float scale_log2, nl2v, delta_i;
    int row, cbeg, cfull;
    bool valid;
};
template <bool MASKED, int G>
__device__ __forceinline__ void bwd_cols(const BwdCols& a) {
    for (int c = a.cbeg; c < a.cfull; c += 16) {
        uint32_t rs[16], rd[16], nq4[4], cj[16];
        tc::tmem_ld_32x16(a.trow + BW_S + c, rs);
        tc::tmem_ld_32x16(a.trow + BW_DP + c, rd);
        lds16i(a.cc_a + c * 4, cj);
        if (MASKED) neq16(a.reg_a + c, a.regi4, nq4);
        float tb[16], dsv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) tb[e] = tc::lds_f32(a.tabrow + cj[e]);
        tc::tmem_ld_wait();
        uint32_t pw[8], dw[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            float v0 = fmaf(__uint_as_float(rs[e]), a.scale_log2, tb[e]) + a.nl2v;
            float v1 = fmaf(__uint_as_float(rs[e + 1]), a.scale_log2, tb[e + 1]) + a.nl2v;
            if (MASKED) { v0 = mask_add(v0, nq4, e); v1 = mask_add(v1, nq4, e + 1); }
            const float p0 = tc::ex2_approx(v0), p1 = tc::ex2_approx(v1);
            dsv[e] = p0 * (__uint_as_float(rd[e]) - a.delta_i);
            dsv[e + 1] = p1 * (__uint_as_float(rd[e + 1]) - a.delta_i);
            pw[e / 2] = tc::pack_bf16(p0, p1);
            dw[e / 2] = tc::pack_bf16(dsv[e], dsv[e + 1]);
        }
        if (a.valid) {
            // d(bias table): at one column step the 32 rows of the warp hit 32 distinct entries, but (row i, col j) and
            // (row i+1, col j+1) share one, so steps must stay ordered -- except for the G columns of one spatial position
            // (consecutive in the permuted key order), which are a constant plane stride apart and never collide inside a
            // warp (validated at kernel start): those are updated as one batch (loads first, then stores).
#pragma unroll
            for (int g = 0; g < 16; g += G) {
                float hv[G];
#pragma unroll
                for (int e = 0; e < G; ++e) hv[e] = tc::lds_f32(a.histrow + cj[g + e]);
#pragma unroll
                for (int e = 0; e < G; ++e) tc::sts_f32(a.histrow + cj[g + e], hv[e] + dsv[g + e]);
            }
        }
        const int u0 = (c - a.cbeg) / 8;
        tc::sts_u4(a.pt_a + sw128_off(a.row, u0), make_uint4(pw[0], pw[1], pw[2], pw[3]));
        tc::sts_u4(a.pt_a + sw128_off(a.row, u0 + 1), make_uint4(pw[4], pw[5], pw[6], pw[7]));
        tc::sts_u4(a.ds_a + sw128_off(a.row, u0), make_uint4(dw[0], dw[1], dw[2], dw[3]));
        tc::sts_u4(a.ds_a + sw128_off(a.row, u0 + 1), make_uint4(dw[4], dw[5], dw[6], dw[7]));
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmDO, const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    const BwdSmem s = bwd_carve(base, p.Lpad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.nH * HD;
    const int h = blockIdx.x / p.groups, gi = blockIdx.x % p.groups;   // CTA pinned to one head

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmQKV);
        tc::prefetch_tmap(&tmKV);
        tc::prefetch_tmap(&tmDO);
        tc::mbar_init(s.kv_full, 1); tc::mbar_init(s.kv_empty, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&s.qd_full[i], 1); tc::mbar_init(&s.qd_empty[i], 1);
            tc::mbar_init(&s.aux_full[i], 1); tc::mbar_init(&s.aux_empty[i], 8);
        }
        tc::mbar_init(s.s_full, 1); tc::mbar_init(s.pds_full, 8); tc::mbar_init(s.dkv_full, 1); tc::mbar_init(s.dq_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(s.tmem_slot, TMEM_COLS);
    // one-time shared-memory state: codes, bias column (x log2e), zeroed histograms and P/dS tiles
    if (threadIdx.x == 0) *s.batched = 1;
    // padded bias / histogram layout: only if the codes really are this window's dense (dd, dh, dw) codes
    int pad = p.pad;
    if (pad) {
        bool ok = p.rowcode[0] + p.colcode[0] == (p.L - 1) / 2;
        for (int n = threadIdx.x; n < p.N && ok; n += NTHREADS) {
            const int e = (n / (p.wh * p.ww)) * p.R1 + ((n / p.ww) % p.wh) * p.R2 + n % p.ww;
            ok = p.rowcode[n] - p.rowcode[0] == e && p.colcode[0] - p.colcode[n] == e;
        }
        if (!__syncthreads_and(ok)) pad = 0;
    }
    for (int n = threadIdx.x; n < 512; n += NTHREADS) {
        s.rc[n] = n < p.N ? p.rowcode[n] + (pad ? pad * ((n / p.ww) % p.wh) : 0) : 0;
        // column c' of the permuted key order holds key j = plane * hw + spatial index
        const int sp = n / p.wd, pl = n - sp * p.wd;
        const int j = pl * p.hw + sp;
        s.cc[n] = sp < p.hw ? (p.colcode[j] + (pad ? pad * (p.wh - 1 - (j / p.ww) % p.wh) : 0)) * 4 : 0;   // BYTE offsets
    }
    for (int l = threadIdx.x; l < p.L; l += NTHREADS)
        s.tab[pad ? l + pad * ((l % p.R1) / p.R2) : l] = __bfloat162float(p.table[(long long)l * p.nH + h]) * LOG2E;
    for (int l = threadIdx.x; l < 8 * p.Lpad; l += NTHREADS) s.hist[l] = 0.f;
    for (int i = threadIdx.x; i < 65536 / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(base + BW_P_OFF)[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    // Validate the batched histogram update: the wd keys of one spatial position must map to wd entries a
    // constant stride apart, and no two rows of one (aligned) warp may be a multiple of that stride apart.
    if (p.wd > 1) {
        const int stride = (s.cc[1] - s.cc[0]) / 4;   // colcode(plane 1) - colcode(plane 0)
        bool ok = stride != 0;
        for (int n = threadIdx.x; n < p.hw * p.wd && ok; n += NTHREADS) {
            const int sp = n / p.wd, pl = n - sp * p.wd;
            ok = (s.cc[n] - s.cc[sp * p.wd]) == 4 * pl * stride;
        }
        for (int i = threadIdx.x; i < p.N && ok; i += NTHREADS) {
            const int blk_end = min(p.N, (i | 31) + 1);
            for (int i2 = i + 1; i2 < blk_end; ++i2)
                if ((s.rc[i2] - s.rc[i]) % stride == 0) { ok = false; break; }
        }
        if (!ok) *s.batched = 0;
    } else if (threadIdx.x == 0) {
        *s.batched = 0;
    }
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s.tmem_slot;
    // valid columns of key block kb (the block's padded columns are all at its end)
    auto block_cols = [&](int kb) { return (min(p.hw, (kb + 1) * p.boxhw) - kb * p.boxhw) * p.wd; };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {   // elect.sync, not lane == 0: the TMA / MMA operands stay warp-uniform for ptxas
            uint32_t kvn = 0, blk = 0;
            for (int b_ = gi; b_ < p.B_; b_ += p.groups) {
                for (int kb = 0; kb < p.nkb; ++kb, ++kvn) {
                    tc::mbar_wait(s.kv_empty, (kvn & 1) ^ 1);
                    tc::mbar_expect_tx(s.kv_full, 2 * BOX_BYTES);
                    tc::tma_load_4d(&tmKV, s.kv_full, base + BW_KV_OFF, C + h * HD, 0, kb * p.boxhw, b_);
                    tc::tma_load_4d(&tmKV, s.kv_full, base + BW_KV_OFF + BOX_BYTES, 2 * C + h * HD, 0, kb * p.boxhw, b_);
                    for (int t = 0; t < p.nq; ++t, ++blk) {
                        const int st = blk & 1; const uint32_t ph = (blk >> 1) & 1;
                        tc::mbar_wait(&s.qd_empty[st], ph ^ 1);
                        tc::mbar_expect_tx(&s.qd_full[st], 2 * BOX_BYTES);
                        uint8_t* dst = base + BW_QD_OFF + st * 2 * BOX_BYTES;
                        tc::tma_load_3d(&tmQKV, &s.qd_full[st], dst, h * HD, t * QT, b_);
                        tc::tma_load_3d(&tmDO, &s.qd_full[st], dst + BOX_BYTES, h * HD, t * QT, b_);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {   // elect.sync, not lane == 0: the TMA / MMA operands stay warp-uniform for ptxas
            uint32_t kvn = 0, blk = 0, pdsph = 0;
            const uint32_t ka = tc::smem_u32(base + BW_KV_OFF), va = ka + BOX_BYTES;
            const uint32_t pa = tc::smem_u32(base + BW_P_OFF), dsa = tc::smem_u32(base + BW_DS_OFF);
            const uint32_t id_t = tc::idesc_bf16(QT, HD, 1, 1);   // A MN-major (P / dS transposed), B MN-major
            const uint32_t id_q = tc::idesc_bf16(QT, HD, 0, 1);   // A K-major (dS), B MN-major (K)
            for (int b_ = gi; b_ < p.B_; b_ += p.groups) {
                for (int kb = 0; kb < p.nkb; ++kb, ++kvn) {
                    const int nkeys = (block_cols(kb) + 15) & ~15;
                    const uint32_t id_s = tc::idesc_bf16(QT, nkeys, 0, 0);
                    tc::mbar_wait(s.kv_full, kvn & 1);
                    for (int t = 0; t < p.nq; ++t, ++blk) {
                        const int st = blk & 1; const uint32_t ph = (blk >> 1) & 1;
                        tc::mbar_wait(&s.qd_full[st], ph);
                        tc::tc_fence_after();
                        const uint32_t qa = tc::smem_u32(base + BW_QD_OFF + st * 2 * BOX_BYTES), doa = qa + BOX_BYTES;
                        if (p.ntail > 0 && t == p.nq - 1) {
                            // ---- transposed tail tile: S^T = K_kb Q_tail^T, dP^T = V_kb dO_tail^T  (128 keys x 16 queries)
                            const uint32_t id_st = tc::idesc_bf16(QT, 16, 0, 0);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                tc::umma_bf16(tmem + BW_S, tc::smem_desc_sw64(ka + k * 32, 0, 512),
                                              tc::smem_desc_sw64(qa + k * 32, 0, 512), id_st, k);
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                tc::umma_bf16(tmem + BW_DP, tc::smem_desc_sw64(va + k * 32, 0, 512),
                                              tc::smem_desc_sw64(doa + k * 32, 0, 512), id_st, k);
                            tc::umma_commit(s.s_full);
                            tc::mbar_wait(s.pds_full, pdsph); pdsph ^= 1;
                            tc::tc_fence_after();
                            // P^T / dS^T: un-swizzled K-major tiles (8-row core matrices 128 B apart, second K chunk = zeros)
                            tc::umma_bf16(tmem + BW_DV, tc::smem_desc(pa + BW_TAIL_PT, BW_TAIL_LBO, 128, 0),
                                          tc::smem_desc_sw64(doa, 0, 512), id_q, t > 0 ? 1u : 0u);
                            tc::umma_bf16(tmem + BW_DK, tc::smem_desc(pa + BW_TAIL_DST, BW_TAIL_LBO, 128, 0),
                                          tc::smem_desc_sw64(qa, 0, 512), id_q, t > 0 ? 1u : 0u);
                            for (int ks = 0; ks < nkeys / 16; ++ks)   // dQ_tail from rows 0..7 of the regular dS tile
                                tc::umma_bf16(tmem + BW_DQ + t * HD,
                                              tc::smem_desc_sw128(dsa + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                                              tc::smem_desc_sw64(ka + ks * 1024, 0, 512), id_q, (kb > 0 || ks > 0) ? 1u : 0u);
                            tc::umma_commit(&s.qd_empty[st]);
                            continue;
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::umma_bf16(tmem + BW_S, tc::smem_desc_sw64(qa + k * 32, 0, 512),
                                          tc::smem_desc_sw64(ka + k * 32, 0, 512), id_s, k);
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            tc::umma_bf16(tmem + BW_DP, tc::smem_desc_sw64(doa + k * 32, 0, 512),
                                          tc::smem_desc_sw64(va + k * 32, 0, 512), id_s, k);
                        tc::umma_commit(s.s_full);
                        tc::mbar_wait(s.pds_full, pdsph); pdsph ^= 1;
                        tc::tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)   // reduce over the 128 queries of the tile
                            tc::umma_bf16(tmem + BW_DV, tc::smem_desc_sw128(pa + ks * 2048, 16384, 1024),
                                          tc::smem_desc_sw64(doa + ks * 1024, 0, 512), id_t, (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            tc::umma_bf16(tmem + BW_DK, tc::smem_desc_sw128(dsa + ks * 2048, 16384, 1024),
                                          tc::smem_desc_sw64(qa + ks * 1024, 0, 512), id_t, (t > 0 || ks > 0) ? 1u : 0u);
                        for (int ks = 0; ks < nkeys / 16; ++ks)   // reduce over the keys of the block
                            tc::umma_bf16(tmem + BW_DQ + t * HD,
                                          tc::smem_desc_sw128(dsa + (ks >> 2) * 16384 + (ks & 3) * 32, 0, 1024),
                                          tc::smem_desc_sw64(ka + ks * 1024, 0, 512), id_q, (kb > 0 || ks > 0) ? 1u : 0u);
                        tc::umma_commit(&s.qd_empty[st]);
                    }
                    tc::umma_commit(s.dkv_full);
                    tc::umma_commit(s.kv_empty);
                }
                tc::umma_commit(s.dq_full);
            }
        }
    } else if (warp == 2) {
        // ===================== aux: region ids of the item's window =====================
        int it = 0;
        for (int b_ = gi; b_ < p.B_; b_ += p.groups, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            tc::mbar_wait(&s.aux_empty[st], ph ^ 1);
            int diff = 0;
            if (p.region) {
                const uint8_t* rg = p.region + (long long)(b_ % p.nW) * p.N;
                const uint8_t r0 = rg[0];
                for (int n = lane; n < 512; n += 32) {
                    const int sp = n / p.wd, pl = n - sp * p.wd;
                    const uint8_t r = sp < p.hw ? rg[pl * p.hw + sp] : r0;   // column (permuted key) order
                    s.reg[st][n] = r;
                    s.regq[st][n] = n < p.N ? rg[n] : r0;                   // query order
                    diff |= (r != r0);
                }
            }
            diff = __any_sync(0xffffffffu, diff);
            if (lane == 0) s.masked[st] = diff;
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_full[st]);
        }
    } else if (warp >= 4) {
        // ===================== softmax / dS / epilogue warps =====================
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float* hist = s.hist + (size_t)(warp - 4) * p.Lpad;
        float* my_delta = s.delta;
        float* my_lse2 = s.lse2;
        uint8_t* ptile = base + BW_P_OFF + half * 16384;    // this warp's 64-key chunk
        uint8_t* dstile = base + BW_DS_OFF + half * 16384;
        int it = 0; uint32_t sph = 0, dkvph = 0, dqph = 0;
        for (int b_ = gi; b_ < p.B_; b_ += p.groups, ++it) {
            const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
            tc::mbar_wait(&s.aux_full[st], ph);
            const bool masked = s.masked[st] != 0;
            const uint8_t* reg = s.reg[st];
            const bool batched = *s.batched != 0;
            for (int kb = 0; kb < p.nkb; ++kb) {
                const int nv = block_cols(kb);                 // valid columns of this key block
                const int nkeys = (nv + 15) & ~15;
                const int cbeg = half * 64, cend = min(cbeg + 64, nkeys);   // my columns inside the block
                for (int t = 0; t < p.nq; ++t) {
                    if (p.ntail > 0 && t == p.nq - 1) {
                        // ================= transposed tail tile: thread = key column `row`, queries tail0 .. tail0 + ntail =================
                        const int tail0 = t * QT;
                        if (kb == 0 && half == 0) {
                            if (q == 0 && lane < 8) {   // delta / lse of the tail queries (thread = query)
                                const int i = tail0 + lane;
                                float delta_i = 0.f, nl2 = 0.f;
                                if (i < p.N) {
                                    const uint4* op = reinterpret_cast<const uint4*>(p.out + ((long long)b_ * p.N + i) * C + h * HD);
                                    const uint4* dp = reinterpret_cast<const uint4*>(p.dout + ((long long)b_ * p.N + i) * C + h * HD);
#pragma unroll
                                    for (int v = 0; v < 4; ++v) {
                                        const uint4 a = __ldg(op + v), b = __ldg(dp + v);
                                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const float2 x = tc::unpack_bf16(aw[e]), y = tc::unpack_bf16(bw[e]);
                                            delta_i = fmaf(x.x, y.x, delta_i);
                                            delta_i = fmaf(x.y, y.y, delta_i);
                                        }
                                    }
                                    nl2 = -p.lse[((long long)b_ * p.nH + h) * p.N + i] * LOG2E;
                                }
                                s.delta[i] = delta_i;
                                s.lse2[i] = nl2;
                            }
                            tc::named_bar_sync(2, 128);   // the four half-0 warps
                        }
                        tc::mbar_wait(s.s_full, sph); sph ^= 1;
                        tc::tc_fence_after();
                        uint8_t* ptile = base + BW_P_OFF;
                        if (half == 1) {
                            // zero chunks (queries 8..15) of the two K = 16 operands
                            *reinterpret_cast<uint4*>(ptile + BW_TAIL_PT + BW_TAIL_LBO + row * 16) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(ptile + BW_TAIL_DST + BW_TAIL_LBO + row * 16) = make_uint4(0, 0, 0, 0);
                        } else {
                            const int k0 = kb * QT;
                            const int j = k0 + row;                 // column index (permuted key order)
                            const bool key_ok = row < nv;
                            uint32_t rs[16], rd[16];
                            tc::tmem_ld_32x16(tmem + lane_base + BW_S, rs);
                            tc::tmem_ld_32x16(tmem + lane_base + BW_DP, rd);
                            const int ccj = key_ok ? s.cc[j] : 0;   // byte offset
                            const uint8_t regj = (masked && key_ok) ? reg[j] : 0;
                            const uint32_t tab_j = tc::smem_u32(s.tab) + ccj, hist_j = tc::smem_u32(hist) + ccj;
                            uint8_t* dsreg = base + BW_DS_OFF + (row >> 6) * 16384 + (row & 7) * 2;   // regular dS tile, my key column
                            const int unit = (row & 63) >> 3;
                            tc::tmem_ld_wait();
                            uint32_t pw[4], dw[4];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                float pv[2], dv[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int i = tail0 + e + u;
                                    float pe = 0.f, ds = 0.f;
                                    if (key_ok && e + u < p.ntail) {
                                        const int rci = s.rc[i] * 4;
                                        float v = fmaf(__uint_as_float(rs[e + u]), p.scale_log2, tc::lds_f32(tab_j + rci)) + s.lse2[i];
                                        if (masked && s.regq[st][i] != regj) v += MASKV;
                                        pe = tc::ex2_approx(v);
                                        ds = pe * (__uint_as_float(rd[e + u]) - s.delta[i]);
                                        // lanes = 32 distinct keys -> 32 distinct table entries for one query: plain read-modify-write
                                        tc::sts_f32(hist_j + rci, tc::lds_f32(hist_j + rci) + ds);
                                    }
                                    pv[u] = pe;
                                    dv[u] = ds;
                                    // dS[query e+u][my key] into the regular (query-major, 128B-swizzled) dS tile for dQ_tail
                                    *reinterpret_cast<__nv_bfloat16*>(dsreg + sw128_off(e + u, unit)) = __float2bfloat16_rn(ds);
                                }
                                pw[e / 2] = tc::pack_bf16(pv[0], pv[1]);
                                dw[e / 2] = tc::pack_bf16(dv[0], dv[1]);
                            }
                            *reinterpret_cast<uint4*>(ptile + BW_TAIL_PT + row * 16) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                            *reinterpret_cast<uint4*>(ptile + BW_TAIL_DST + row * 16) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                        }
                        tc::fence_proxy_async();
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(s.pds_full);
                        continue;
                    }
                    const int i = t * QT + row;
                    const bool valid = i < p.N;
                    const bool warp_valid = t * QT + q * 32 < p.N;
                    float delta_i, nl2;
                    if (kb == 0) {
                        delta_i = 0.f; nl2 = 0.f;
                        if (valid) {
                            const uint4* op = reinterpret_cast<const uint4*>(p.out + ((long long)b_ * p.N + i) * C + h * HD);
                            const uint4* dp = reinterpret_cast<const uint4*>(p.dout + ((long long)b_ * p.N + i) * C + h * HD);
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const uint4 a = __ldg(op + v), b = __ldg(dp + v);
                                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 x = tc::unpack_bf16(aw[e]), y = tc::unpack_bf16(bw[e]);
                                    delta_i = fmaf(x.x, y.x, delta_i);
                                    delta_i = fmaf(x.y, y.y, delta_i);
                                }
                            }
                            nl2 = -p.lse[((long long)b_ * p.nH + h) * p.N + i] * LOG2E;
                        }
                        my_delta[i] = delta_i;
                        my_lse2[i] = nl2;
                    } else {
                        delta_i = my_delta[i];
                        nl2 = my_lse2[i];
                    }
                    const int ic = valid ? i : p.N - 1;
                    const int rci = s.rc[ic];
                    const uint8_t regi = masked ? s.regq[st][ic] : 0;
                    long long t_a = ACLK();
                    tc::mbar_wait(s.s_full, sph); sph ^= 1;
                    tc::tc_fence_after();
                    long long t_b = ACLK();
                    if (!warp_valid) {
                        // rows beyond the window: contribute zeros to the reductions over queries
                        for (int u = 0; u < 8; ++u) {
                            *reinterpret_cast<uint4*>(ptile + sw128_off(row, u)) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(dstile + sw128_off(row, u)) = make_uint4(0, 0, 0, 0);
                        }
                    } else {
                        const uint32_t tabrow = tc::smem_u32(s.tab + rci), histrow = tc::smem_u32(hist + rci);
                        const uint32_t cc_a = tc::smem_u32(s.cc), reg_a = tc::smem_u32(reg);
                        const uint32_t pt_a = tc::smem_u32(ptile), ds_a = tc::smem_u32(dstile);
                        const uint32_t regi4 = (uint32_t)regi * 0x01010101u;
                        const float nl2v = valid ? nl2 : -INFINITY;          // invalid rows: P = dS = 0
                        const int k0 = kb * QT;
                        const int cfull = min(cend, nv & ~15);   // chunks without padded columns
                        {
                            const BwdCols bc{tmem + lane_base, cc_a + k0 * 4, reg_a + k0, tabrow, histrow, pt_a, ds_a, regi4,
                                             p.scale_log2, nl2v, delta_i, row, cbeg, cfull, valid};
                            const int gmode = batched ? p.wd : 1;
                            if (masked) {
                                if (gmode == 8) bwd_cols<true, 8>(bc); else if (gmode == 4) bwd_cols<true, 4>(bc); else bwd_cols<true, 1>(bc);
                            } else {
                                if (gmode == 8) bwd_cols<false, 8>(bc); else if (gmode == 4) bwd_cols<false, 4>(bc); else bwd_cols<false, 1>(bc);
                            }
                        }
                        for (int c = max(cbeg, cfull); c < cend; c += 16) {   // chunk with columns >= N
                            uint32_t rs[16], rd[16];
                            tc::tmem_ld_32x16(tmem + lane_base + BW_S + c, rs);
                            tc::tmem_ld_32x16(tmem + lane_base + BW_DP + c, rd);
                            tc::tmem_ld_wait();
                            uint32_t pw[8], dw[8];
#pragma unroll
                            for (int e = 0; e < 16; e += 2) {
                                float pv[2], dv[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int j = k0 + c + e + u;   // column index (permuted key order)
                                    float pe = 0.f, ds = 0.f;
                                    if (valid && c + e + u < nv) {
                                        const int off = s.cc[j];
                                        float v = fmaf(__uint_as_float(rs[e + u]), p.scale_log2, tc::lds_f32(tabrow + off)) + nl2;
                                        if (masked && reg[j] != regi) v += MASKV;
                                        pe = tc::ex2_approx(v);
                                        ds = pe * (__uint_as_float(rd[e + u]) - delta_i);
                                        tc::sts_f32(histrow + off, tc::lds_f32(histrow + off) + ds);
                                    }
                                    pv[u] = pe;
                                    dv[u] = ds;
                                }
                                pw[e / 2] = tc::pack_bf16(pv[0], pv[1]);
                                dw[e / 2] = tc::pack_bf16(dv[0], dv[1]);
                            }
                            const int u0 = (c - cbeg) / 8;
                            *reinterpret_cast<uint4*>(ptile + sw128_off(row, u0)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                            *reinterpret_cast<uint4*>(ptile + sw128_off(row, u0 + 1)) = make_uint4(pw[4], pw[5], pw[6], pw[7]);
                            *reinterpret_cast<uint4*>(dstile + sw128_off(row, u0)) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                            *reinterpret_cast<uint4*>(dstile + sw128_off(row, u0 + 1)) = make_uint4(dw[4], dw[5], dw[6], dw[7]);
                        }
                    }
                    tc::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(s.pds_full);
                    if (VSW_ATTN_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 128) {
                        long long t_c = ACLK();
                        p.dbg[0] += t_b - t_a; p.dbg[1] += t_c - t_b; p.dbg[2] += 1;
                    }
                }
                // ---- dK (half 0) / dV (half 1) of this key block
                long long t_d = ACLK();
                tc::mbar_wait(s.dkv_full, dkvph); dkvph ^= 1;
                tc::tc_fence_after();
                if (VSW_ATTN_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 128) { p.dbg[3] += ACLK() - t_d; p.dbg[4] += 1; }
                {
                    // TMEM lane `row` of the block = column c' -> key (plane c' % wd, spatial kb*boxhw + c' / wd)
                    const int key = row < nv ? (row % p.wd) * p.hw + kb * p.boxhw + row / p.wd : p.N;
                    uint32_t lo[16], hi[16];
                    const uint32_t col = half == 0 ? BW_DK : BW_DV;
                    tc::tmem_ld_32x16(tmem + lane_base + col, lo);
                    tc::tmem_ld_32x16(tmem + lane_base + col + 16, hi);
                    tc::tmem_ld_wait();
                    if (key < p.N) {
                        const float sc = half == 0 ? p.scale : 1.0f;
                        uint4* dst = reinterpret_cast<uint4*>(p.dqkv + (((long long)b_ * p.N + key) * 3 + 1 + half) * C + h * HD);
                        uint4 u;
                        u.x = tc::pack_bf16(__uint_as_float(lo[0]) * sc, __uint_as_float(lo[1]) * sc);
                        u.y = tc::pack_bf16(__uint_as_float(lo[2]) * sc, __uint_as_float(lo[3]) * sc);
                        u.z = tc::pack_bf16(__uint_as_float(lo[4]) * sc, __uint_as_float(lo[5]) * sc);
                        u.w = tc::pack_bf16(__uint_as_float(lo[6]) * sc, __uint_as_float(lo[7]) * sc);
                        dst[0] = u;
                        u.x = tc::pack_bf16(__uint_as_float(lo[8]) * sc, __uint_as_float(lo[9]) * sc);
                        u.y = tc::pack_bf16(__uint_as_float(lo[10]) * sc, __uint_as_float(lo[11]) * sc);
                        u.z = tc::pack_bf16(__uint_as_float(lo[12]) * sc, __uint_as_float(lo[13]) * sc);
                        u.w = tc::pack_bf16(__uint_as_float(lo[14]) * sc, __uint_as_float(lo[15]) * sc);
                        dst[1] = u;
                        u.x = tc::pack_bf16(__uint_as_float(hi[0]) * sc, __uint_as_float(hi[1]) * sc);
                        u.y = tc::pack_bf16(__uint_as_float(hi[2]) * sc, __uint_as_float(hi[3]) * sc);
                        u.z = tc::pack_bf16(__uint_as_float(hi[4]) * sc, __uint_as_float(hi[5]) * sc);
                        u.w = tc::pack_bf16(__uint_as_float(hi[6]) * sc, __uint_as_float(hi[7]) * sc);
                        dst[2] = u;
                        u.x = tc::pack_bf16(__uint_as_float(hi[8]) * sc, __uint_as_float(hi[9]) * sc);
                        u.y = tc::pack_bf16(__uint_as_float(hi[10]) * sc, __uint_as_float(hi[11]) * sc);
                        u.z = tc::pack_bf16(__uint_as_float(hi[12]) * sc, __uint_as_float(hi[13]) * sc);
                        u.w = tc::pack_bf16(__uint_as_float(hi[14]) * sc, __uint_as_float(hi[15]) * sc);
                        dst[3] = u;
                    }
                }
                tc::tc_fence_before();
            }
            // ---- dQ of every query tile (16 columns per half)
            tc::mbar_wait(s.dq_full, dqph); dqph ^= 1;
            tc::tc_fence_after();
            for (int t = 0; t < p.nq; ++t) {
                const int i = t * QT + row;
                uint32_t o[16];
                tc::tmem_ld_32x16(tmem + lane_base + BW_DQ + t * HD + half * 16, o);
                tc::tmem_ld_wait();
                if (i < p.N) {
                    uint4* dst = reinterpret_cast<uint4*>(p.dqkv + (((long long)b_ * p.N + i) * 3) * C + h * HD + half * 16);
                    uint4 u;
                    u.x = tc::pack_bf16(__uint_as_float(o[0]) * p.scale, __uint_as_float(o[1]) * p.scale);
                    u.y = tc::pack_bf16(__uint_as_float(o[2]) * p.scale, __uint_as_float(o[3]) * p.scale);
                    u.z = tc::pack_bf16(__uint_as_float(o[4]) * p.scale, __uint_as_float(o[5]) * p.scale);
                    u.w = tc::pack_bf16(__uint_as_float(o[6]) * p.scale, __uint_as_float(o[7]) * p.scale);
                    dst[0] = u;
                    u.x = tc::pack_bf16(__uint_as_float(o[8]) * p.scale, __uint_as_float(o[9]) * p.scale);
                    u.y = tc::pack_bf16(__uint_as_float(o[10]) * p.scale, __uint_as_float(o[11]) * p.scale);
                    u.z = tc::pack_bf16(__uint_as_float(o[12]) * p.scale, __uint_as_float(o[13]) * p.scale);
                    u.w = tc::pack_bf16(__uint_as_float(o[14]) * p.scale, __uint_as_float(o[15]) * p.scale);
                    dst[1] = u;
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s.aux_empty[st]);
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    // flush the eight private histograms of this CTA in a fixed order
    {
        float* part = p.dbias_part + ((long long)gi * p.nH + h) * p.L;
        for (int l = threadIdx.x; l < p.L; l += NTHREADS) {
            float t = 0.f;
#pragma unroll
            const int lp = pad ? l + pad * ((l % p.R1) / p.R2) : l;
            for (int w8 = 0; w8 < 8; ++w8) t += s.hist[(size_t)w8 * p.Lpad + lp];
            part[l] = t;
        }
    }
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

__global__ void tc_dbias_reduce_kernel(const float* __restrict__ part, int groups, int nH, int L, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // out layout (L, nH)
    if (idx >= L * nH) return;
    const int l = idx / nH, h = idx % nH;
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += part[((long long)g * nH + h) * L + l];
    out[idx] = t;
}

int bwd_groups(int B_, int nH) {
    int g = kNumSMs / nH;
    if (g < 1) g = 1;
    if (g > B_) g = B_;
    return g;
}

}  // namespace

size_t tc_attn_bwd_workspace(int B_, int N, int nH, int hd, int L) {
    (void)N; (void)hd;
    if (nH > kNumSMs) return 0;
    return (size_t)bwd_groups(B_, nH) * nH * L * sizeof(float);
}

int tc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, void* dqkv, float* dbias,
                int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims, void* ws, size_t ws_bytes,
                cudaStream_t st) {
    BiasLayout bl = bias_layout(window_dims, L);
    if (bwd_smem_bytes((bl.Lphys + 3) / 4 * 4) > 227 * 1024) bl = bias_layout(0, L);   // no room for the padded copies
    const int Lpad = (bl.Lphys + 3) / 4 * 4;
    const size_t smem = bwd_smem_bytes(Lpad);
    const int planes = window_dims & 0xFF;
    if (hd != HD || N > 448 || N < 1 || nH > kNumSMs || smem > 227 * 1024 || !aligned16(qkv) || !aligned16(out) ||
        !aligned16(dout) || !aligned16(dqkv)) {
        set_error("tcgen05 window attention bwd: needs head_dim 32, N <= 448, bias table fitting shared memory "
                  "(hd=%d N=%d L=%d smem=%zu)", hd, N, L, smem);
        return VSW_ERR_UNSUPPORTED;
    }
    const int groups = bwd_groups(B_, nH);
    if (ws_bytes < (size_t)groups * nH * L * sizeof(float)) { set_error("tcgen05 attention bwd: workspace too small"); return VSW_ERR_WORKSPACE; }
    const int C = nH * HD;
    // key permutation (plane-minor) when the window depth is a power of two that divides N; else natural order
    int wd = (planes == 8 || planes == 4 || planes == 2) && N % planes == 0 ? planes : 1;
    const int hw = N / wd, boxhw = QT / wd;
    CUtensorMap tmQKV, tmKV, tmDO;
    if (!make_tmap_3d_bf16(&tmQKV, qkv, B_, N, 3 * C, 3 * C, (uint64_t)N * 3 * C, QT, HD, 64)) return VSW_ERR_CUDA;
    if (!make_tmap_3d_bf16(&tmDO, dout, B_, N, C, C, (uint64_t)N * C, QT, HD, 64)) return VSW_ERR_CUDA;
    {
        const uint64_t dims[4] = {(uint64_t)3 * C, (uint64_t)wd, (uint64_t)hw, (uint64_t)B_};
        const uint64_t strides[4] = {1, (uint64_t)hw * 3 * C, (uint64_t)3 * C, (uint64_t)N * 3 * C};
        const uint32_t box[4] = {HD, (uint32_t)wd, (uint32_t)boxhw, 1};
        if (!make_tmap_nd_bf16(&tmKV, qkv, 4, dims, strides, box, 64)) return VSW_ERR_CUDA;
    }
    BwdParams p{};
    p.table = (const __nv_bfloat16*)table; p.rowcode = rowcode; p.colcode = colcode; p.region = region;
    p.out = (const __nv_bfloat16*)out; p.dout = (const __nv_bfloat16*)dout; p.lse = lse;
    p.dqkv = (__nv_bfloat16*)dqkv; p.dbias_part = (float*)ws;
    p.B_ = B_; p.nW = nW; p.N = N; p.nH = nH; p.L = L; p.Lpad = Lpad; p.groups = groups;
    p.scale = scale; p.scale_log2 = scale * LOG2E;
    p.wh = bl.wh; p.ww = bl.ww; p.R1 = bl.R1; p.R2 = bl.R2; p.pad = bl.pad;
    p.Npad = (N + 15) / 16 * 16; p.nq = (N + QT - 1) / QT;
    p.ntail = (N % QT != 0 && N % QT <= 8) ? N % QT : 0;
    p.wd = wd; p.hw = hw; p.boxhw = boxhw; p.nkb = (hw + boxhw - 1) / boxhw;
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
            fprintf(stderr, "[vsw attn bwd] blocks=%lld avg cycles: wait_s=%lld softmax=%lld ; kb epilogues=%lld wait_dkv=%lld\n", h[2], h[0]/(h[2]+1), h[1]/(h[2]+1), h[4], h[3]/(h[4]+1));
            cudaMemset(dbg, 0, 64);
        }
    }
    static bool configured[kMaxDevices] = {};   // per device; benign race (the attribute is idempotent)
    const int dev = current_device();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attn bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
        configured[dev] = true;
    }
    attn_bwd_tc_kernel<<<groups * nH, NTHREADS, smem, st>>>(tmQKV, tmKV, tmDO, p);
    int rc = check_launch("attn_bwd_tc");
    if (rc) return rc;
    tc_dbias_reduce_kernel<<<ceil_div((long long)L * nH, 256), 256, 0, st>>>((const float*)ws, groups, nH, L, dbias);
    return check_launch("attn_dbias_reduce_tc");
}

}  // namespace vsw
