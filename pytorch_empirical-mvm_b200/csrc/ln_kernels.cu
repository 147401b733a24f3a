// LayerNorm family: HBM-bound kernels, 16-byte vector loads, sub-warp shuffle reductions.
//   vsw_ln_fwd / vsw_ln_bwd          norm1 (+pad+roll+window_partition gather), norm2, final norm, patch norm
//   vsw_merge_ln_fwd / _bwd          PatchMerging 2x2 gather + LN(4C)
// Reference: visbackbone/video_swin.py:211-229, 248, 273-286, 401-405, 479.
// One row is owned by a group of LANES (<=32) lanes; each lane keeps VPL 16-byte vectors of the row
// in registers, so x is read exactly once.  Column reductions (dgamma/dbeta) run as a separate
// fixed-order two-pass column kernel (deterministic, no atomics).
#include "common.cuh"

namespace vsw {

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// forward.  GROUPS = number of concatenated source rows per output row (1 = plain/gather LN,
// 4 = PatchMerging).  For GROUPS==1 a negative map entry zeroes the OUTPUT row (post-norm pad);
// for GROUPS==4 it zeroes that INPUT group (pre-norm pad).
// ------------------------------------------------------------------------------------------
template <typename T, typename TO, int LANES, int VPL, int GROUPS>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                     const T* __restrict__ beta, const int32_t* __restrict__ map,
                                                     TO* __restrict__ y, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int B, int Tin, int Tout, int C,
                                                     float eps) {
    constexpr int VN = Vec16<T>::N;
    constexpr int RPW = 32 / LANES;  // rows per warp
    const int Cw = C * GROUPS;       // normalised width
    const int nvec = Cw / VN;
    const int vec_per_group = C / VN;
    const int lane = threadIdx.x & 31, sub = lane / LANES, l = lane % LANES;
    const long long nrows = (long long)B * Tout;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long warp_stride = (long long)gridDim.x * (blockDim.x >> 5);
    const float inv_c = 1.0f / (float)Cw;

    for (long long row0 = warp_global * RPW; row0 < nrows; row0 += warp_stride * RPW) {
        const long long row = row0 + sub;
        const bool row_ok = row < nrows;
        const int b = row_ok ? (int)(row / Tout) : 0;
        const int r = row_ok ? (int)(row - (long long)b * Tout) : 0;
        int src1 = r;
        if (GROUPS == 1 && map) src1 = map[r];
        float v[VPL][VN];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int vi = l + k * LANES;
#pragma unroll
            for (int e = 0; e < VN; ++e) v[k][e] = 0.f;
            if (row_ok && vi < nvec) {
                int src = src1, col = vi * VN;
                if (GROUPS > 1) {
                    const int g = vi / vec_per_group;
                    src = map[(long long)r * GROUPS + g];
                    col = (vi - g * vec_per_group) * VN;
                }
                if (src >= 0) load_vec<T>(x + ((long long)b * Tin + src) * C + col, v[k]);
            }
#pragma unroll
            for (int e = 0; e < VN; ++e) s += v[k][e];
        }
        const float mu = group_sum<LANES>(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int vi = l + k * LANES;
            if (vi < nvec) {
#pragma unroll
                for (int e = 0; e < VN; ++e) { const float d = v[k][e] - mu; q += d * d; }
            }
        }
        const float rs = rsqrtf(group_sum<LANES>(q) * inv_c + eps);
        const bool zero_row = (GROUPS == 1) && (src1 < 0);
        if (row_ok) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = l + k * LANES;
                if (vi < nvec) {
                    float gmm[VN], bt[VN], o[VN];
                    load_vec<T>(gamma + vi * VN, gmm);
                    load_vec<T>(beta + vi * VN, bt);
#pragma unroll
                    for (int e = 0; e < VN; ++e) o[e] = zero_row ? 0.f : (v[k][e] - mu) * rs * gmm[e] + bt[e];
                    // TO may be wider than T (fp32 output of a bf16 row): store element groups of TO's vector width
                    constexpr int VO = Vec16<TO>::N;
                    if constexpr (VO == VN) {
                        store_vec<TO>(y + row * Cw + vi * VN, o);
                    } else if constexpr (VO > VN) {   // TO narrower than T (16-bit output of an fp32 row): one 8-byte store
                        store_n<TO, VN>(y + row * Cw + vi * VN, o);
                    } else {
#pragma unroll
                        for (int h = 0; h < VN / VO; ++h) {
                            float oo[VO];
#pragma unroll
                            for (int e = 0; e < VO; ++e) oo[e] = o[h * VO + e];
                            store_vec<TO>(y + row * Cw + vi * VN + h * VO, oo);
                        }
                    }
                }
            }
            if (l == 0 && mean_out) {
                mean_out[row] = zero_row ? 0.f : mu;
                rstd_out[row] = zero_row ? 0.f : rs;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward (row part): dx = dres + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma
// ------------------------------------------------------------------------------------------
// FUSE: also accumulate the column reductions (dgamma = sum dy*xhat, dbeta = sum dy) in registers -- each lane owns
// the same column vectors for every row it visits -- and emit one fixed-order partial per block (part[block][2][Cw]).
template <typename T, typename TDY, int LANES, int VPL, int GROUPS, bool FUSE>
__global__ void __launch_bounds__(256, (VPL == 1 ? 4 : (VPL == 2 ? 2 : 1))) ln_bwd_kernel(const TDY* __restrict__ dy, const T* __restrict__ x,
                                                     const T* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const int32_t* __restrict__ map,
                                                     const T* __restrict__ dres, T* __restrict__ dx, int B, int Tin,
                                                     int Tout, int C, float* __restrict__ part) {
    constexpr int VN = Vec16<T>::N;
    constexpr int RPW = 32 / LANES;
    const int Cw = C * GROUPS;
    const int nvec = Cw / VN;
    const int vec_per_group = C / VN;
    const int lane = threadIdx.x & 31, sub = lane / LANES, l = lane % LANES;
    const long long nrows = (long long)B * Tout;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long warp_stride = (long long)gridDim.x * (blockDim.x >> 5);
    const float inv_c = 1.0f / (float)Cw;
    float ag[FUSE ? VPL : 1][VN], ab[FUSE ? VPL : 1][VN];
#pragma unroll
    for (int k = 0; k < (FUSE ? VPL : 1); ++k)
#pragma unroll
        for (int e = 0; e < VN; ++e) { ag[k][e] = 0.f; ab[k][e] = 0.f; }

    using PackT = Pack<T, VN>;
    const T* res_src = dres ? dres : x;               // always-valid address for the (possibly absent) residual gradient
    for (long long row0 = warp_global * RPW; row0 < nrows; row0 += warp_stride * RPW) {
        const long long row = row0 + sub;
        const bool row_ok = row < nrows;
        const long long rowc = row_ok ? row : 0;
        const int b = (int)(rowc / Tout);
        const int r = (int)(rowc - (long long)b * Tout);
        int src1 = r;
        if (GROUPS == 1 && map) src1 = map[r];
        const float mu = row_ok ? mean[rowc] : 0.f, rs = row_ok ? rstd[rowc] : 0.f;
        // ---- phase 1: predicates and (clamped, always valid) addresses -- so that every global load of the row can be
        //      issued unconditionally and back to back (one exposed memory latency per row instead of one per vector)
        bool act[VPL], has_x[VPL];
        long long off[VPL];
        int vic[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int vi = l + k * LANES;
            act[k] = row_ok && vi < nvec;
            vic[k] = vi < nvec ? vi : 0;
            int src = src1, col = vic[k] * VN;
            if (GROUPS > 1) {
                const int gi = vic[k] / vec_per_group;
                src = map[(long long)r * GROUPS + gi];
                col = (vic[k] - gi * vec_per_group) * VN;
            }
            has_x[k] = act[k] && src >= 0;
            off[k] = has_x[k] ? ((long long)b * Tin + src) * C + col : 0;
        }
        // ---- phase 2: loads
        float xh[VPL][VN], g[VPL][VN];
        PackT rr[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            constexpr int VD = Vec16<TDY>::N;
            if constexpr (VD >= VN) {   // same width, or a 16-bit dy against fp32 rows (one 8-byte load)
                load_n<TDY, VN>(dy + rowc * Cw + vic[k] * VN, g[k]);
            } else {
#pragma unroll
                for (int h = 0; h < VN / VD; ++h) {
                    float dd[VD];
                    load_vec<TDY>(dy + rowc * Cw + vic[k] * VN + h * VD, dd);
#pragma unroll
                    for (int e = 0; e < VD; ++e) g[k][h * VD + e] = dd[e];
                }
            }
            load_vec<T>(x + off[k], xh[k]);
            rr[k] = *reinterpret_cast<const PackT*>(res_src + off[k]);
        }
        // ---- phase 3: row statistics of g = dy * gamma
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            float gm[VN];
            load_vec<T>(gamma + vic[k] * VN, gm);
            // padded INPUT group (GROUPS > 1, src < 0): x = 0 took part in the statistics
            const bool pad_in = GROUPS > 1 && act[k] && !has_x[k];
            const bool fuse_k = FUSE && act[k] && !(GROUPS == 1 && !has_x[k]);   // zeroed (padded) output rows carry no gradient
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                const float d = act[k] ? g[k][e] : 0.f;
                const float xv = has_x[k] ? (xh[k][e] - mu) * rs : (pad_in ? (0.f - mu) * rs : 0.f);
                xh[k][e] = xv;
                if (fuse_k) { ag[k][e] = fmaf(d, xv, ag[k][e]); ab[k][e] += d; }
                const float gg = d * gm[e];
                g[k][e] = gg;
                s1 += gg;
                s2 += gg * xv;
            }
        }
        const float c1 = group_sum<LANES>(s1) * inv_c;
        const float c2 = group_sum<LANES>(s2) * inv_c;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            if (has_x[k]) {
                float o[VN];
#pragma unroll
                for (int e = 0; e < VN; ++e) o[e] = rs * (g[k][e] - c1 - xh[k][e] * c2);
                if (dres) {
#pragma unroll
                    for (int e = 0; e < VN; ++e) o[e] += to_f<T>(rr[k].v[e]);
                }
                if (dx) store_vec<T>(dx + off[k], o);
            }
        }
    }
    if (FUSE) {
        // combine the row sub-groups of a warp, then the 8 warps of the block (fixed order), one partial per block
        __shared__ float red[8][32 * 4 * 8 + 8];
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int vi = l + k * LANES;
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                    float v = pass == 0 ? ag[k][e] : ab[k][e];
#pragma unroll
                    for (int o = LANES; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (sub == 0 && vi < nvec) red[warp][vi * VN + e] = v;
                }
            }
            __syncthreads();
            for (int c = threadIdx.x; c < Cw; c += blockDim.x) {
                float t = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) t += red[w8][c];
                part[((size_t)blockIdx.x * 2 + pass) * Cw + c] = t;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// column reductions: part[split][0][c] = sum_rows dy*xhat, part[split][1][c] = sum_rows dy
// thread = one 16B column vector; blockDim = (TX, TY); rows strided by TY*gridDim.y
// ------------------------------------------------------------------------------------------
template <typename T, typename TDY, int GROUPS>
__global__ void __launch_bounds__(256) ln_colsum_kernel(const TDY* __restrict__ dy, const T* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const int32_t* __restrict__ map, float* __restrict__ part,
                                                        int B, int Tin, int Tout, int C) {
    constexpr int VN = Vec16<T>::N;
    const int Cw = C * GROUPS;
    const int nvec = Cw / VN;
    const int vec_per_group = C / VN;
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    const long long nrows = (long long)B * Tout;
    float a1[VN], a2[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) { a1[e] = 0.f; a2[e] = 0.f; }
    if (vi < nvec) {
        const int gi = GROUPS > 1 ? vi / vec_per_group : 0;
        const int col = GROUPS > 1 ? (vi - gi * vec_per_group) * VN : vi * VN;
        for (long long row = (long long)blockIdx.y * blockDim.y + threadIdx.y; row < nrows;
             row += (long long)gridDim.y * blockDim.y) {
            const int b = (int)(row / Tout);
            const int r = (int)(row - (long long)b * Tout);
            int src = r;
            if (GROUPS > 1) src = map[(long long)r * GROUPS + gi];
            else if (map) src = map[r];
            if (GROUPS == 1 && src < 0) continue;  // zeroed output row: no gradient
            float d[VN];
            constexpr int VD = Vec16<TDY>::N;
            if constexpr (VD >= VN) {
                load_n<TDY, VN>(dy + row * Cw + vi * VN, d);
            } else {
#pragma unroll
                for (int h = 0; h < VN / VD; ++h) {
                    float dd[VD];
                    load_vec<TDY>(dy + row * Cw + vi * VN + h * VD, dd);
#pragma unroll
                    for (int e = 0; e < VD; ++e) d[h * VD + e] = dd[e];
                }
            }
            float xv[VN];
#pragma unroll
            for (int e = 0; e < VN; ++e) xv[e] = 0.f;
            if (src >= 0) load_vec<T>(x + ((long long)b * Tin + src) * C + col, xv);
            const float mu = mean[row], rs = rstd[row];
#pragma unroll
            for (int e = 0; e < VN; ++e) {
                a1[e] += d[e] * (xv[e] - mu) * rs;
                a2[e] += d[e];
            }
        }
    }
    // fixed-order reduction over threadIdx.y through shared memory
    extern __shared__ float sm[];  // [TY][TX][2*VN]
    float* mine = sm + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 2 * VN;
#pragma unroll
    for (int e = 0; e < VN; ++e) { mine[e] = a1[e]; mine[VN + e] = a2[e]; }
    __syncthreads();
    if (threadIdx.y == 0 && vi < nvec) {
        for (int ty = 1; ty < blockDim.y; ++ty) {
            const float* o = sm + ((size_t)ty * blockDim.x + threadIdx.x) * 2 * VN;
#pragma unroll
            for (int e = 0; e < VN; ++e) { a1[e] += o[e]; a2[e] += o[VN + e]; }
        }
        float* p = part + (size_t)blockIdx.y * 2 * Cw;
#pragma unroll
        for (int e = 0; e < VN; ++e) { p[vi * VN + e] = a1[e]; p[Cw + vi * VN + e] = a2[e]; }
    }
}

// blockDim (8, 128): 8 columns (one 32-byte sector per partial row) per block, the partial rows strided over ty,
// fixed-order combine through shared memory.  Many small blocks: the partial matrix is read by the whole chip.
constexpr int kFinTX = 8, kFinTY = 128;
__global__ void __launch_bounds__(kFinTX * kFinTY)
colsum_finish_kernel(const float* __restrict__ part, int splits, int Cw, void* __restrict__ dgamma,
                     void* __restrict__ dbeta, int gdtype) {
    __shared__ float sm[2][kFinTY][kFinTX + 1];
    const int c = blockIdx.x * kFinTX + threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    if (c < Cw)
        for (int s = threadIdx.y; s < splits; s += kFinTY) {
            s1 += part[(size_t)s * 2 * Cw + c];
            s2 += part[(size_t)s * 2 * Cw + Cw + c];
        }
    sm[0][threadIdx.y][threadIdx.x] = s1;
    sm[1][threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y < 2 && c < Cw) {   // row 0 of the block finishes dgamma, row 1 dbeta (serial, fixed order)
        float t = 0.f;
        for (int y = 0; y < kFinTY; ++y) t += sm[threadIdx.y][y][threadIdx.x];
        void* out = threadIdx.y == 0 ? dgamma : dbeta;   // written in the parameters' dtype: no cast kernel behind this one
        if (out) {
            if (gdtype == VSW_F32) reinterpret_cast<float*>(out)[c] = t;
            else if (gdtype == VSW_BF16) reinterpret_cast<__nv_bfloat16*>(out)[c] = __float2bfloat16_rn(t);
            else reinterpret_cast<__half*>(out)[c] = __float2half_rn(t);
        }
    }
}

constexpr int kColSplits = 1184;   // 8 blocks per SM: partial rows of the fused row kernel

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
struct RowCfg { int lanes, vpl; };
static bool pick_row_cfg(int nvec, RowCfg* c) {
    if (nvec <= 0) return false;
    int lanes = 1;
    while (lanes < 32 && lanes < nvec) lanes <<= 1;
    int vpl = (nvec + lanes - 1) / lanes;
    static const int allowed[] = {1, 2, 4, 8, 16, 32};
    for (int a : allowed)
        if (vpl <= a) { c->lanes = lanes; c->vpl = a; return true; }
    return false;
}

#define VSW_ROW_DISPATCH(cfg, ...)                                                                         \
    do {                                                                                                   \
        if (cfg.lanes == 32) {                                                                             \
            switch (cfg.vpl) {                                                                             \
                case 1: { constexpr int LANES = 32, VPL = 1; __VA_ARGS__; } break;                         \
                case 2: { constexpr int LANES = 32, VPL = 2; __VA_ARGS__; } break;                         \
                case 4: { constexpr int LANES = 32, VPL = 4; __VA_ARGS__; } break;                         \
                case 8: { constexpr int LANES = 32, VPL = 8; __VA_ARGS__; } break;                         \
                case 16: { constexpr int LANES = 32, VPL = 16; __VA_ARGS__; } break;                       \
                default: { constexpr int LANES = 32, VPL = 32; __VA_ARGS__; } break;                       \
            }                                                                                              \
        } else if (cfg.lanes == 16) { constexpr int LANES = 16, VPL = 1; __VA_ARGS__; }                    \
        else if (cfg.lanes == 8) { constexpr int LANES = 8, VPL = 1; __VA_ARGS__; }                        \
        else if (cfg.lanes == 4) { constexpr int LANES = 4, VPL = 1; __VA_ARGS__; }                        \
        else if (cfg.lanes == 2) { constexpr int LANES = 2, VPL = 1; __VA_ARGS__; }                        \
        else { constexpr int LANES = 1, VPL = 1; __VA_ARGS__; }                                            \
    } while (0)

static int row_grid(long long nrows, int lanes) {
    const long long rows_per_block = 8LL * (32 / lanes);
    long long blocks = (nrows + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

template <typename T, typename TO, int GROUPS>
static int launch_ln_fwd(const void* x, const void* gamma, const void* beta, const int32_t* map, void* y, float* mean,
                         float* rstd, int B, int Tin, int Tout, int C, float eps, cudaStream_t st) {
    constexpr int VN = Vec16<T>::N;
    RowCfg cfg;
    VSW_REQUIRE((C % VN) == 0 && pick_row_cfg(C * GROUPS / VN, &cfg), VSW_ERR_UNSUPPORTED,
                "layernorm: C=%d must be a multiple of %d and <= %d", C, VN, 32 * 32 * VN / GROUPS);
    const int grid = row_grid((long long)B * Tout, cfg.lanes);
    VSW_ROW_DISPATCH(cfg, (ln_fwd_kernel<T, TO, LANES, VPL, GROUPS><<<grid, 256, 0, st>>>(
                              (const T*)x, (const T*)gamma, (const T*)beta, map, (TO*)y, mean, rstd, B, Tin, Tout, C,
                              eps)));
    return check_launch("ln_fwd");
}

template <typename T, typename TDY, int GROUPS>
static int launch_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                         const int32_t* map, const void* dres, void* dx, void* dgamma, void* dbeta, int B, int Tin,
                         int Tout, int C, void* ws, size_t ws_bytes, cudaStream_t st, int gdtype = VSW_F32) {
    constexpr int VN = Vec16<T>::N;
    RowCfg cfg;
    const int Cw = C * GROUPS;
    VSW_REQUIRE((C % VN) == 0 && pick_row_cfg(Cw / VN, &cfg), VSW_ERR_UNSUPPORTED,
                "layernorm bwd: C=%d must be a multiple of %d", C, VN);
    const bool want_cols = dgamma || dbeta;
    if (want_cols)
        VSW_REQUIRE(ws && ws_bytes >= (size_t)kColSplits * 2 * Cw * sizeof(float), VSW_ERR_WORKSPACE,
                    "layernorm bwd: workspace %zu < %zu", ws_bytes, (size_t)kColSplits * 2 * Cw * sizeof(float));
    // fused column reductions when the per-lane accumulators fit in registers (<= 4 vectors per lane)
    const bool fuse = want_cols && cfg.vpl <= 4 && Cw <= 1024;
    if (dx || fuse) {
        int grid = row_grid((long long)B * Tout, cfg.lanes);
        if (fuse) {
            // persistent sizing: one resident wave (register-limited occupancy x SMs), >= 4 rows per warp, so the
            // per-block column-partial epilogue is amortised and no block waits for a second wave
            VSW_ROW_DISPATCH(cfg, ({
                auto kern = ln_bwd_kernel<T, TDY, LANES, (VPL <= 4 ? VPL : 1), GROUPS, true>;
                static int resident = 0;
                if (!resident) {
                    int nb = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, 0) != cudaSuccess || nb < 1) nb = 1;
                    resident = nb * kNumSMs;
                }
                const long long rows_per_block = 8LL * (32 / LANES) * 4;
                long long g = ((long long)B * Tout + rows_per_block - 1) / rows_per_block;
                if (g > resident) g = resident;
                if (g > kColSplits) g = kColSplits;
                if (g < 1) g = 1;
                grid = (int)g;
                kern<<<grid, 256, 0, st>>>((const TDY*)dy, (const T*)x, (const T*)gamma, mean, rstd, map, (const T*)dres,
                                           (T*)dx, B, Tin, Tout, C, (float*)ws);
            }));
        } else {
            VSW_ROW_DISPATCH(cfg, (ln_bwd_kernel<T, TDY, LANES, VPL, GROUPS, false><<<grid, 256, 0, st>>>(
                                      (const TDY*)dy, (const T*)x, (const T*)gamma, mean, rstd, map, (const T*)dres, (T*)dx,
                                      B, Tin, Tout, C, nullptr)));
        }
        int rc = check_launch("ln_bwd");
        if (rc) return rc;
        if (fuse) {
            colsum_finish_kernel<<<ceil_div(Cw, kFinTX), dim3(kFinTX, kFinTY), 0, st>>>((const float*)ws, grid, Cw, dgamma, dbeta, gdtype);
            return check_launch("ln_colsum_finish");
        }
    }
    if (want_cols) {
        const int nvec = Cw / VN;
        const int TX = nvec >= 32 ? 32 : 16, TY = 256 / TX;
        long long nrows = (long long)B * Tout;
        int splits = (int)((nrows + TY - 1) / TY);
        if (splits > 256) splits = 256;
        if (splits < 1) splits = 1;
        dim3 grid(ceil_div(nvec, TX), splits), block(TX, TY);
        const size_t smem = (size_t)256 * 2 * VN * sizeof(float);
        ln_colsum_kernel<T, TDY, GROUPS><<<grid, block, smem, st>>>((const TDY*)dy, (const T*)x, mean, rstd, map,
                                                                     (float*)ws, B, Tin, Tout, C);
        int rc = check_launch("ln_colsum");
        if (rc) return rc;
        colsum_finish_kernel<<<ceil_div(Cw, kFinTX), dim3(kFinTX, kFinTY), 0, st>>>((const float*)ws, splits, Cw, dgamma, dbeta, gdtype);
        rc = check_launch("ln_colsum_finish");
        if (rc) return rc;
    }
    return VSW_OK;
}

}  // namespace vsw

using namespace vsw;

extern "C" size_t vsw_ln_bwd_workspace(int C) { return (size_t)kColSplits * 2 * (size_t)C * sizeof(float); }

extern "C" int vsw_ln_fwd(const void* x, const void* gamma, const void* beta, const int32_t* map, void* y, float* mean,
                          float* rstd, int B, int Tin, int Tout, int C, float eps, int dtype, int out_dtype,
                          void* stream) {
    VSW_REQUIRE(x && gamma && beta && y && B > 0 && Tin > 0 && Tout > 0 && C > 0, VSW_ERR_ARG, "vsw_ln_fwd: bad args");
    VSW_REQUIRE(map || Tin == Tout, VSW_ERR_ARG, "vsw_ln_fwd: identity map needs Tin == Tout");
    VSW_REQUIRE((mean == nullptr) == (rstd == nullptr), VSW_ERR_ARG, "vsw_ln_fwd: mean/rstd must both be given or NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == dtype) {
        VSW_DISPATCH_DTYPE(dtype, T,
                           return (launch_ln_fwd<T, T, 1>(x, gamma, beta, map, y, mean, rstd, B, Tin, Tout, C, eps, st)));
    }
    if (dtype == VSW_F32) {   // fp32 residual stream, 16-bit branch input (autocast: LN reads fp32, the Linear behind it runs in 16 bit)
        VSW_REQUIRE(out_dtype == VSW_BF16 || out_dtype == VSW_F16, VSW_ERR_DTYPE, "vsw_ln_fwd: bad out_dtype %d", out_dtype);
        if (out_dtype == VSW_BF16) return launch_ln_fwd<float, __nv_bfloat16, 1>(x, gamma, beta, map, y, mean, rstd, B, Tin, Tout, C, eps, st);
        return launch_ln_fwd<float, __half, 1>(x, gamma, beta, map, y, mean, rstd, B, Tin, Tout, C, eps, st);
    }
    VSW_REQUIRE(out_dtype == VSW_F32, VSW_ERR_DTYPE, "vsw_ln_fwd: out_dtype must equal dtype or be fp32");
    VSW_DISPATCH_DTYPE(dtype, T,
                       return (launch_ln_fwd<T, float, 1>(x, gamma, beta, map, y, mean, rstd, B, Tin, Tout, C, eps, st)));
    return VSW_OK;
}

extern "C" int vsw_ln_bwd_ex(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                             const int32_t* map, const void* dres, void* dx, void* dgamma, void* dbeta, int B, int Tin,
                             int Tout, int C, int dtype, int dy_dtype, int dparam_dtype, void* ws, size_t ws_bytes,
                             void* stream) {
    VSW_REQUIRE(dparam_dtype == VSW_F32 || dparam_dtype == VSW_BF16 || dparam_dtype == VSW_F16, VSW_ERR_DTYPE,
                "vsw_ln_bwd: bad dparam_dtype %d", dparam_dtype);
    VSW_REQUIRE(dy && x && gamma && mean && rstd && B > 0 && Tin > 0 && Tout > 0 && C > 0, VSW_ERR_ARG,
                "vsw_ln_bwd: bad args");
    VSW_REQUIRE(map || Tin == Tout, VSW_ERR_ARG, "vsw_ln_bwd: identity map needs Tin == Tout");
    cudaStream_t st = (cudaStream_t)stream;
    if (dy_dtype == dtype) {
        VSW_DISPATCH_DTYPE(dtype, T,
                           return (launch_ln_bwd<T, T, 1>(dy, x, gamma, mean, rstd, map, dres, dx, dgamma, dbeta, B, Tin,
                                                          Tout, C, ws, ws_bytes, st, dparam_dtype)));
    }
    if (dtype == VSW_F32) {   // fp32 x / dres / dx with a 16-bit dy (the autocast case of vsw_ln_fwd above)
        VSW_REQUIRE(dy_dtype == VSW_BF16 || dy_dtype == VSW_F16, VSW_ERR_DTYPE, "vsw_ln_bwd: bad dy_dtype %d", dy_dtype);
        if (dy_dtype == VSW_BF16)
            return launch_ln_bwd<float, __nv_bfloat16, 1>(dy, x, gamma, mean, rstd, map, dres, dx, dgamma, dbeta, B, Tin, Tout, C, ws, ws_bytes, st, dparam_dtype);
        return launch_ln_bwd<float, __half, 1>(dy, x, gamma, mean, rstd, map, dres, dx, dgamma, dbeta, B, Tin, Tout, C, ws, ws_bytes, st, dparam_dtype);
    }
    VSW_REQUIRE(dy_dtype == VSW_F32, VSW_ERR_DTYPE, "vsw_ln_bwd: dy_dtype must equal dtype or be fp32");
    VSW_DISPATCH_DTYPE(dtype, T,
                       return (launch_ln_bwd<T, float, 1>(dy, x, gamma, mean, rstd, map, dres, dx, dgamma, dbeta, B, Tin,
                                                          Tout, C, ws, ws_bytes, st, dparam_dtype)));
    return VSW_OK;
}

extern "C" int vsw_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                          const int32_t* map, const void* dres, void* dx, float* dgamma, float* dbeta, int B, int Tin,
                          int Tout, int C, int dtype, int dy_dtype, void* ws, size_t ws_bytes, void* stream) {
    return vsw_ln_bwd_ex(dy, x, gamma, mean, rstd, map, dres, dx, dgamma, dbeta, B, Tin, Tout, C, dtype, dy_dtype, VSW_F32, ws,
                         ws_bytes, stream);
}

extern "C" int vsw_merge_ln_fwd(const void* x, const void* gamma, const void* beta, const int32_t* map4, void* y,
                                float* mean, float* rstd, int B, int Tin, int Tout, int C, float eps, int dtype,
                                void* stream) {
    VSW_REQUIRE(x && gamma && beta && map4 && y && B > 0 && Tin > 0 && Tout > 0 && C > 0, VSW_ERR_ARG,
                "vsw_merge_ln_fwd: bad args");
    VSW_REQUIRE((mean == nullptr) == (rstd == nullptr), VSW_ERR_ARG, "vsw_merge_ln_fwd: mean/rstd both or none");
    cudaStream_t st = (cudaStream_t)stream;
    VSW_DISPATCH_DTYPE(dtype, T,
                       return (launch_ln_fwd<T, T, 4>(x, gamma, beta, map4, y, mean, rstd, B, Tin, Tout, C, eps, st)));
    return VSW_OK;
}

extern "C" int vsw_merge_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                                const int32_t* map4, void* dx, float* dgamma, float* dbeta, int B, int Tin, int Tout,
                                int C, int dtype, void* ws, size_t ws_bytes, void* stream) {
    VSW_REQUIRE(dy && x && gamma && mean && rstd && map4 && B > 0 && Tin > 0 && Tout > 0 && C > 0, VSW_ERR_ARG,
                "vsw_merge_ln_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    VSW_DISPATCH_DTYPE(dtype, T,
                       return (launch_ln_bwd<T, T, 4>(dy, x, gamma, mean, rstd, map4, nullptr, dx, dgamma, dbeta, B, Tin,
                                                      Tout, C, ws, ws_bytes, st)));
    return VSW_OK;
}
