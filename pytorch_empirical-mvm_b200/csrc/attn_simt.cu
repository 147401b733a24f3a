// CUDA-core fused window attention (fp32 math, storage dtype templated): forward and
// recompute-based backward.  Used for VSW_F32 (parity mode), for head_dim != 32, and as the
// cross-check of the tcgen05 attention kernels.
// Reference: WindowAttention3D.forward, visbackbone/video_swin.py:149-169; backward formulas SURVEY A5.
// Scores never reach HBM: per (window, head) K/V chunks sit in shared memory, each warp owns query
// rows (fwd, dQ pass) or key rows (dK/dV pass) and runs an online softmax.
#include "common.cuh"
#include "attn.cuh"

namespace vsw {

constexpr int HD = 32;        // padded head dim (hd <= 32 supported here)
constexpr int KC = 256;       // keys (or queries) per shared-memory chunk
constexpr int RPW = 8;        // rows per warp; a block of NWARPS warps owns NWARPS*8 rows
constexpr int LD = HD + 1;    // padded smem row

struct AttnArgs {
    const void* qkv; const void* table; const int32_t* rowcode; const int32_t* colcode;
    const uint8_t* region; const void* dmask;
    void* out; float* lse;
    const void* dout; void* dqkv; float* delta; float* dbias_part;
    int B_, nW, N, nH, hd, L; float scale;
    int groups;   // blocks per head in the dQ pass (dbias partial slots)
};

template <typename T>
__device__ __forceinline__ float mask_val(const AttnArgs& a, const uint8_t* reg, int w, int i, int j) {
    if (a.dmask) return to_f<T>(((const T*)a.dmask)[((long long)w * a.N + i) * a.N + j]);
    if (a.region) return reg[i] != reg[j] ? -100.0f : 0.0f;
    return 0.0f;
}

// loads `rows` rows [r0, r0+rows) of operand `which` (0=q,1=k,2=v) of (b_,h) into dst[KC][LD] as fp32
template <typename T>
__device__ __forceinline__ void load_rows_qkv(const AttnArgs& a, int b_, int h, int which, int r0, int rows,
                                              float* dst) {
    const T* base = (const T*)a.qkv;
    for (int idx = threadIdx.x; idx < KC * HD; idx += blockDim.x) {
        const int r = idx / HD, d = idx % HD;
        float v = 0.f;
        if (r < rows && d < a.hd)
            v = to_f<T>(base[(((long long)b_ * a.N + (r0 + r)) * 3 + which) * (a.nH * a.hd) + h * a.hd + d]);
        dst[r * LD + d] = v;
    }
}
template <typename T>
__device__ __forceinline__ void load_rows_bnc(const AttnArgs& a, const void* src, int b_, int h, int r0, int rows,
                                              float* dst) {
    const T* base = (const T*)src;
    for (int idx = threadIdx.x; idx < KC * HD; idx += blockDim.x) {
        const int r = idx / HD, d = idx % HD;
        float v = 0.f;
        if (r < rows && d < a.hd) v = to_f<T>(base[((long long)b_ * a.N + (r0 + r)) * (a.nH * a.hd) + h * a.hd + d]);
        dst[r * LD + d] = v;
    }
}

struct SmemLayout {
    float* A;      // [KC][LD]   K (fwd/dq)  | Q  (dkv)
    float* Bv;     // [KC][LD]   V (fwd/dq)  | dO (dkv)
    float* tab;    // [L]
    float* lse;    // [KC]   (dkv)
    float* dlt;    // [KC]   (dkv)
    float* hist;   // [warps][L] (dq)
    int* rc; int* cc; uint8_t* reg;
};
__device__ __forceinline__ SmemLayout carve(float* sm, int L, int N, int nwarps_hist, bool need_stats) {
    SmemLayout s;
    s.A = sm; sm += KC * LD;
    s.Bv = sm; sm += KC * LD;
    s.tab = sm; sm += L;
    s.lse = sm; s.dlt = sm + KC; if (need_stats) sm += 2 * KC;
    s.hist = sm; sm += (size_t)nwarps_hist * L;
    s.rc = (int*)sm; s.cc = s.rc + N; s.reg = (uint8_t*)(s.cc + N);
    return s;
}
static size_t smem_bytes(int L, int N, int nwarps_hist, bool need_stats) {
    size_t f = 2 * (size_t)KC * LD + L + (need_stats ? 2 * KC : 0) + (size_t)nwarps_hist * L;
    return f * 4 + (size_t)N * 8 + N + 16;
}

template <typename T>
__device__ __forceinline__ void load_common(const AttnArgs& a, const SmemLayout& s, int h, int w) {
    const T* table = (const T*)a.table;
    for (int l = threadIdx.x; l < a.L; l += blockDim.x) s.tab[l] = to_f<T>(table[(long long)l * a.nH + h]);
    for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
        s.rc[n] = a.rowcode[n];
        s.cc[n] = a.colcode[n];
        s.reg[n] = a.region ? a.region[(long long)w * a.N + n] : 0;
    }
}

// ------------------------------------------------------------------------------------------
// forward (MODE 0) and dQ/delta/dbias pass (MODE 1).  grid = (B_*nH or persistent groups, rowgroups)
// ------------------------------------------------------------------------------------------
template <typename T, int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_rows_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    constexpr int RPB = NWARPS * RPW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowgroups = (a.N + RPB - 1) / RPB;
    const SmemLayout s = carve(sm, a.L, a.N, MODE == 1 ? NWARPS : 0, false);
    const int C = a.nH * a.hd;

    // MODE 0: one work item per block.  MODE 1: block (h, g) loops over items (b_, rowgroup) of head h.
    const int h = MODE == 0 ? (int)(blockIdx.x % a.nH) : (int)(blockIdx.x % a.nH);
    const int g = MODE == 0 ? 0 : (int)(blockIdx.x / a.nH);
    const long long items = MODE == 0 ? 1 : (long long)a.B_ * rowgroups;
    const long long item_step = MODE == 0 ? 1 : a.groups;

    if (MODE == 1)
        for (int l = threadIdx.x; l < NWARPS * a.L; l += blockDim.x) s.hist[l] = 0.f;
    int cur_w = -1;
    bool tab_loaded = false;

    for (long long item = (MODE == 0 ? 0 : g); item < items; item += item_step) {
        int b_, rg;
        if (MODE == 0) { b_ = blockIdx.x / a.nH; rg = blockIdx.y; }
        else { b_ = (int)(item / rowgroups); rg = (int)(item % rowgroups); }
        const int w = b_ % a.nW;
        __syncthreads();
        if (!tab_loaded || w != cur_w) { load_common<T>(a, s, h, w); tab_loaded = true; cur_w = w; }
        const int row0 = rg * RPB + warp * RPW;

        float m_[RPW], l_[RPW], acc[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) { m_[r] = -INFINITY; l_[r] = 0.f; acc[r] = 0.f; }

        for (int c0 = 0; c0 < a.N; c0 += KC) {
            const int kc = min(KC, a.N - c0);
            __syncthreads();
            load_rows_qkv<T>(a, b_, h, 1, c0, kc, s.A);
            load_rows_qkv<T>(a, b_, h, 2, c0, kc, s.Bv);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int i = row0 + r;
                if (i >= a.N) continue;  // warp-uniform
                float q[HD];
                {
                    const T* qp = (const T*)a.qkv + (((long long)b_ * a.N + i) * 3 + 0) * C + h * a.hd;
#pragma unroll
                    for (int d = 0; d < HD; ++d) q[d] = d < a.hd ? to_f<T>(qp[d]) * a.scale : 0.f;
                }
                float dO[MODE == 1 ? HD : 1];
                float lse_i = 0.f, delta_i = 0.f;
                if (MODE == 1) {
                    const T* dop = (const T*)a.dout + ((long long)b_ * a.N + i) * C + h * a.hd;
                    const T* op = (const T*)a.out + ((long long)b_ * a.N + i) * C + h * a.hd;
#pragma unroll
                    for (int d = 0; d < HD; ++d) dO[d] = d < a.hd ? to_f<T>(dop[d]) : 0.f;
                    float part = lane < a.hd ? to_f<T>(dop[lane]) * to_f<T>(op[lane]) : 0.f;
                    delta_i = warp_sum(part);
                    lse_i = a.lse[((long long)b_ * a.nH + h) * a.N + i];
                    if (c0 == 0 && lane == 0) a.delta[((long long)b_ * a.nH + h) * a.N + i] = delta_i;
                }
                const int rci = s.rc[i];
                float sc[KC / 32];
                float cmax = -INFINITY;
#pragma unroll
                for (int t = 0; t < KC / 32; ++t) {
                    const int jj = lane + 32 * t;
                    float v = -INFINITY;
                    if (jj < kc) {
                        const int j = c0 + jj;
                        float dot = 0.f;
#pragma unroll
                        for (int d = 0; d < HD; ++d) dot = fmaf(q[d], s.A[jj * LD + d], dot);
                        v = dot + s.tab[rci + s.cc[j]] + mask_val<T>(a, s.reg, w, i, j);
                    }
                    sc[t] = v;
                    cmax = fmaxf(cmax, v);
                }
                if (MODE == 0) {
                    cmax = warp_max(cmax);
                    const float mnew = fmaxf(m_[r], cmax);
                    const float corr = __expf(m_[r] - mnew);
                    float ps = 0.f;
#pragma unroll
                    for (int t = 0; t < KC / 32; ++t) { sc[t] = __expf(sc[t] - mnew); ps += sc[t]; }
                    ps = warp_sum(ps);
                    l_[r] = l_[r] * corr + ps;
                    acc[r] *= corr;
                    m_[r] = mnew;
                    const int tmax = (kc + 31) / 32;
                    for (int t = 0; t < tmax; ++t)
#pragma unroll
                        for (int src = 0; src < 32; ++src) {
                            const float pj = __shfl_sync(0xffffffffu, sc[t], src);
                            acc[r] = fmaf(pj, s.Bv[(src + 32 * t) * LD + lane], acc[r]);
                        }
                } else {
                    // ds = p * (dO.v - delta); accumulate dq and the per-warp bias histogram
#pragma unroll
                    for (int t = 0; t < KC / 32; ++t) {
                        const int jj = lane + 32 * t;
                        float ds = 0.f;
                        if (jj < kc) {
                            const float p = __expf(sc[t] - lse_i);
                            float dp = 0.f;
#pragma unroll
                            for (int d = 0; d < HD; ++d) dp = fmaf(dO[d], s.Bv[jj * LD + d], dp);
                            ds = p * (dp - delta_i);
                            s.hist[warp * a.L + rci + s.cc[c0 + jj]] += ds;  // distinct j -> distinct slot
                        }
                        sc[t] = ds;
                    }
                    __syncwarp();   // the next query row of this warp may hit the same histogram slots from other lanes
                    const int tmax = (kc + 31) / 32;
                    for (int t = 0; t < tmax; ++t)
#pragma unroll
                        for (int src = 0; src < 32; ++src) {
                            const float dsj = __shfl_sync(0xffffffffu, sc[t], src);
                            acc[r] = fmaf(dsj, s.A[(src + 32 * t) * LD + lane], acc[r]);
                        }
                }
            }
        }
        // write results of this item
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int i = row0 + r;
            if (i >= a.N) continue;
            if (MODE == 0) {
                if (lane < a.hd)
                    ((T*)a.out)[((long long)b_ * a.N + i) * C + h * a.hd + lane] = from_f<T>(acc[r] / l_[r]);
                if (lane == 0) a.lse[((long long)b_ * a.nH + h) * a.N + i] = m_[r] + __logf(l_[r]);
            } else {
                if (lane < a.hd)
                    ((T*)a.dqkv)[(((long long)b_ * a.N + i) * 3 + 0) * C + h * a.hd + lane] =
                        from_f<T>(acc[r] * a.scale);
            }
        }
    }
    if (MODE == 1) {
        __syncthreads();
        float* part = a.dbias_part + ((long long)g * a.nH + h) * a.L;
        for (int l = threadIdx.x; l < a.L; l += blockDim.x) {
            float t = 0.f;
            for (int wv = 0; wv < NWARPS; ++wv) t += s.hist[wv * a.L + l];
            part[l] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------
// dK / dV pass: warp owns key rows, chunks over queries.  grid = (B_*nH, keygroups)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_dkv_kernel(AttnArgs a) {
    extern __shared__ float sm[];
    constexpr int NWARPS = 8, RPB = NWARPS * RPW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SmemLayout s = carve(sm, a.L, a.N, 0, true);
    const int C = a.nH * a.hd;
    const int b_ = blockIdx.x / a.nH, h = blockIdx.x % a.nH, w = b_ % a.nW;
    load_common<T>(a, s, h, w);
    const int row0 = blockIdx.y * RPB + warp * RPW;

    float dk[RPW], dv[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) { dk[r] = 0.f; dv[r] = 0.f; }

    for (int c0 = 0; c0 < a.N; c0 += KC) {
        const int qc = min(KC, a.N - c0);
        __syncthreads();
        load_rows_qkv<T>(a, b_, h, 0, c0, qc, s.A);          // Q chunk
        load_rows_bnc<T>(a, a.dout, b_, h, c0, qc, s.Bv);    // dO chunk
        for (int i = threadIdx.x; i < KC; i += blockDim.x) {
            const bool ok = i < qc;
            s.lse[i] = ok ? a.lse[((long long)b_ * a.nH + h) * a.N + c0 + i] : 0.f;
            s.dlt[i] = ok ? a.delta[((long long)b_ * a.nH + h) * a.N + c0 + i] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int j = row0 + r;
            if (j >= a.N) continue;
            float kreg[HD], vreg[HD];
            {
                const T* kp = (const T*)a.qkv + (((long long)b_ * a.N + j) * 3 + 1) * C + h * a.hd;
                const T* vp = kp + C;
#pragma unroll
                for (int d = 0; d < HD; ++d) {
                    kreg[d] = d < a.hd ? to_f<T>(kp[d]) * a.scale : 0.f;
                    vreg[d] = d < a.hd ? to_f<T>(vp[d]) : 0.f;
                }
            }
            const int ccj = s.cc[j];
            float pv[KC / 32], dsv[KC / 32];
#pragma unroll
            for (int t = 0; t < KC / 32; ++t) {
                const int ii = lane + 32 * t;
                float p = 0.f, ds = 0.f;
                if (ii < qc) {
                    const int i = c0 + ii;
                    float dot = 0.f, dp = 0.f;
#pragma unroll
                    for (int d = 0; d < HD; ++d) {
                        dot = fmaf(s.A[ii * LD + d], kreg[d], dot);
                        dp = fmaf(s.Bv[ii * LD + d], vreg[d], dp);
                    }
                    const float sc = dot + s.tab[s.rc[i] + ccj] + mask_val<T>(a, s.reg, w, i, j);
                    p = __expf(sc - s.lse[ii]);
                    ds = p * (dp - s.dlt[ii]);
                }
                pv[t] = p;
                dsv[t] = ds;
            }
            const int tmax = (qc + 31) / 32;
            for (int t = 0; t < tmax; ++t)
#pragma unroll
                for (int src = 0; src < 32; ++src) {
                    const float pi = __shfl_sync(0xffffffffu, pv[t], src);
                    const float di = __shfl_sync(0xffffffffu, dsv[t], src);
                    const int ii = src + 32 * t;
                    dv[r] = fmaf(pi, s.Bv[ii * LD + lane], dv[r]);
                    dk[r] = fmaf(di, s.A[ii * LD + lane], dk[r]);
                }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int j = row0 + r;
        if (j >= a.N || lane >= a.hd) continue;
        T* base = (T*)a.dqkv + (((long long)b_ * a.N + j) * 3) * C + h * a.hd + lane;
        base[C] = from_f<T>(dk[r] * a.scale);
        base[2 * C] = from_f<T>(dv[r]);
    }
}

__global__ void dbias_reduce_kernel(const float* __restrict__ part, int groups, int nH, int L,
                                    float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over L*nH, out layout (L, nH)
    if (idx >= L * nH) return;
    const int l = idx / nH, h = idx % nH;
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += part[((long long)g * nH + h) * L + l];
    out[idx] = t;
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static int hist_warps(int L, int N) {
    if (smem_bytes(L, N, 8, false) <= 200 * 1024) return 8;
    if (smem_bytes(L, N, 4, false) <= 200 * 1024) return 4;
    return 2;
}
static int dq_groups(int B_, int N, int nH, int L) {
    const int rpb = hist_warps(L, N) * RPW;
    const long long items = (long long)B_ * ((N + rpb - 1) / rpb);
    long long g = (2LL * kNumSMs + nH - 1) / nH;
    if (g > items) g = items;
    if (g < 1) g = 1;
    return (int)g;
}

size_t simt_attn_bwd_workspace(int B_, int N, int nH, int hd, int L) {
    (void)hd;
    return (size_t)B_ * nH * N * sizeof(float) + (size_t)dq_groups(B_, N, nH, L) * nH * L * sizeof(float);
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) { set_error("window attention: needs %zu bytes of shared memory", bytes); return VSW_ERR_UNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
    return VSW_OK;
}

template <typename T>
static int run_fwd(const AttnArgs& a, cudaStream_t st) {
    const size_t sb = smem_bytes(a.L, a.N, 0, false);
    int rc = set_smem(attn_rows_kernel<T, 0, 8>, sb);
    if (rc) return rc;
    dim3 grid(a.B_ * a.nH, (a.N + 8 * RPW - 1) / (8 * RPW));
    attn_rows_kernel<T, 0, 8><<<grid, 256, sb, st>>>(a);
    return check_launch("attn_fwd_simt");
}

template <typename T>
static int run_bwd(AttnArgs a, float* dbias, cudaStream_t st) {
    a.groups = dq_groups(a.B_, a.N, a.nH, a.L);
    // per-warp private histograms (deterministic): 8 warps when they fit, else 4 / 2
    int rc;
    const dim3 grid1(a.nH * a.groups);
    const int hw = hist_warps(a.L, a.N);
    size_t sb8 = smem_bytes(a.L, a.N, 8, false), sb4 = smem_bytes(a.L, a.N, 4, false),
           sb2 = smem_bytes(a.L, a.N, 2, false);
    if (hw == 8) {
        if ((rc = set_smem(attn_rows_kernel<T, 1, 8>, sb8))) return rc;
        attn_rows_kernel<T, 1, 8><<<grid1, 256, sb8, st>>>(a);
    } else if (hw == 4) {
        if ((rc = set_smem(attn_rows_kernel<T, 1, 4>, sb4))) return rc;
        attn_rows_kernel<T, 1, 4><<<grid1, 128, sb4, st>>>(a);
    } else {
        if ((rc = set_smem(attn_rows_kernel<T, 1, 2>, sb2))) return rc;
        attn_rows_kernel<T, 1, 2><<<grid1, 64, sb2, st>>>(a);
    }
    if ((rc = check_launch("attn_dq_simt"))) return rc;
    const size_t sbk = smem_bytes(a.L, a.N, 0, true);
    if ((rc = set_smem(attn_dkv_kernel<T>, sbk))) return rc;
    dim3 grid2(a.B_ * a.nH, (a.N + 8 * RPW - 1) / (8 * RPW));
    attn_dkv_kernel<T><<<grid2, 256, sbk, st>>>(a);
    if ((rc = check_launch("attn_dkv_simt"))) return rc;
    dbias_reduce_kernel<<<ceil_div((long long)a.L * a.nH, 256), 256, 0, st>>>(a.dbias_part, a.groups, a.nH, a.L, dbias);
    return check_launch("attn_dbias_reduce");
}

int simt_attn_fwd(const void* qkv, const void* table, const int32_t* rowcode, const int32_t* colcode,
                  const uint8_t* region, const void* dmask, void* out, float* lse, int B_, int nW, int N, int nH,
                  int hd, int L, float scale, int dtype, cudaStream_t st) {
    VSW_REQUIRE(hd >= 1 && hd <= HD, VSW_ERR_UNSUPPORTED, "window attention: head_dim %d > %d not supported", hd, HD);
    AttnArgs a{};
    a.qkv = qkv; a.table = table; a.rowcode = rowcode; a.colcode = colcode; a.region = region; a.dmask = dmask;
    a.out = out; a.lse = lse; a.B_ = B_; a.nW = nW; a.N = N; a.nH = nH; a.hd = hd; a.L = L; a.scale = scale;
    VSW_DISPATCH_DTYPE(dtype, T, return run_fwd<T>(a, st));
    return VSW_OK;
}

int simt_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const void* table,
                  const int32_t* rowcode, const int32_t* colcode, const uint8_t* region, const void* dmask, void* dqkv,
                  float* dbias, int B_, int nW, int N, int nH, int hd, int L, float scale, int dtype, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
    VSW_REQUIRE(hd >= 1 && hd <= HD, VSW_ERR_UNSUPPORTED, "window attention: head_dim %d > %d not supported", hd, HD);
    VSW_REQUIRE(ws && ws_bytes >= simt_attn_bwd_workspace(B_, N, nH, hd, L), VSW_ERR_WORKSPACE,
                "window attention bwd: workspace too small");
    AttnArgs a{};
    a.qkv = qkv; a.table = table; a.rowcode = rowcode; a.colcode = colcode; a.region = region; a.dmask = dmask;
    a.out = const_cast<void*>(out); a.lse = const_cast<float*>(lse); a.dout = dout; a.dqkv = dqkv;
    a.delta = (float*)ws;
    a.dbias_part = (float*)ws + (size_t)B_ * nH * N;
    a.B_ = B_; a.nW = nW; a.N = N; a.nH = nH; a.hd = hd; a.L = L; a.scale = scale;
    VSW_DISPATCH_DTYPE(dtype, T, return run_bwd<T>(a, dbias, st));
    return VSW_OK;
}

}  // namespace vsw
