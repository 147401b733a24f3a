// EncVideo tail: what consumes the Swin output every step.
// Reference: model.py:57-70 -- after `fc` (8E -> hidden) the tokens of each frame get a class row in front,
// the position / frame-length (or frame-order) embeddings are added and the result is LayerNorm-ed:
//     pre[b,t,p,:] = (p == 0 ? emb_cls : f[b,t,p-1,:]) + emb_pos[p,:] + (odr[b][t] == t ? emb_len[t,:] : emb_odr)
//     out[b, t*P + p, :] = LN(pre[b,t,p,:]) * gamma + beta                    P = 1 + h*w
//     m_img[b, t*P + p]  = vt_mask ? vt_mask[b,t,p] : 1                        (model.py:72-76)
// One kernel forward (cat + 2 adds + LayerNorm + mask = 6 eager kernels and 4 HBM round trips in the reference),
// one row kernel + two small fixed-order column reductions backward.  HBM-bound, one warp per row, the row lives in
// registers, 8/16-byte vector accesses.  The embeddings and gamma/beta are fp32 (tiny, L2-resident); f / out / dy
// carry the activation dtypes.  Deterministic: no atomics, fixed reduction orders.
#include "common.cuh"

namespace vsw {

constexpr int kTailMaxK = 8;         // chunks of 4 channels per lane: C <= 32 * 4 * 8 = 1024 (kernels are templated on KCH <= 8)
constexpr int kTailBwdOcc = 2;       // persistent blocks per SM in the backward row kernel (<= 128 registers x 256 threads)
constexpr int kTailMaxBlocks = 148 * kTailBwdOcc;

template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ p, float (&o)[4]) {
    Pack<T, 4> pk = *reinterpret_cast<const Pack<T, 4>*>(p);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = to_f<T>(pk.v[e]);
}
template <typename T>
__device__ __forceinline__ void store4(T* __restrict__ p, const float (&v)[4]) {
    Pack<T, 4> pk;
#pragma unroll
    for (int e = 0; e < 4; ++e) pk.v[e] = from_f<T>(v[e]);
    *reinterpret_cast<Pack<T, 4>*>(p) = pk;
}

// pre-LayerNorm row (b,t,p) into registers; returns its sum
template <typename T, int KCH>
__device__ __forceinline__ float tail_row(const T* __restrict__ f, const float* __restrict__ cls,
                                          const float* __restrict__ pos, const float* __restrict__ sel, long long bt,
                                          int p, int hw, int C, int lane, float (&v)[KCH][4]) {
    const T* frow = f + (bt * hw + (p > 0 ? p - 1 : 0)) * (long long)C;
    const float* prow = pos + (long long)p * C;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int col = (lane + 32 * k) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[k][e] = 0.f;
        if (col < C) {
            float a[4], b4[4], c4[4];
            if (p > 0) load4<T>(frow + col, a);
            else load4<float>(cls + col, a);
            load4<float>(prow + col, b4);
            load4<float>(sel + col, c4);
#pragma unroll
            for (int e = 0; e < 4; ++e) { v[k][e] = a[e] + b4[e] + c4[e]; s += v[k][e]; }
        }
    }
    return s;
}

template <typename T, typename TO, int KCH>
__global__ void __launch_bounds__(256, 4) enc_tail_fwd_kernel(const T* __restrict__ f, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, const float* __restrict__ len,
                                                           const float* __restrict__ odr_emb,
                                                           const int32_t* __restrict__ odr,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const int64_t* __restrict__ vt_mask, TO* __restrict__ out,
                                                           int64_t* __restrict__ m_out, float* __restrict__ mean_out,
                                                           float* __restrict__ rstd_out, int B, int Tn, int hw, int C,
                                                           float eps) {
    const int P = hw + 1;
    const long long nrows = (long long)B * Tn * P;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const float inv_c = 1.0f / (float)C;
    for (long long row = warp0; row < nrows; row += wstride) {
        const long long bt = row / P;
        const int p = (int)(row - bt * P);
        const int t = (int)(bt % Tn);
        const float* sel = (odr && odr[bt] != t) ? odr_emb : len + (long long)t * C;
        float v[KCH][4];
        const float mu = warp_sum(tail_row<T, KCH>(f, cls, pos, sel, bt, p, hw, C, lane, v)) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < KCH; ++k)
            if ((lane + 32 * k) * 4 < C) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = v[k][e] - mu; q += d * d; }
            }
        const float rs = rsqrtf(warp_sum(q) * inv_c + eps);
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            const int col = (lane + 32 * k) * 4;
            if (col < C) {
                float g4[4], b4[4], o[4];
                load4<float>(gamma + col, g4);
                load4<float>(beta + col, b4);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (v[k][e] - mu) * rs * g4[e] + b4[e];
                store4<TO>(out + row * C + col, o);
            }
        }
        if (lane == 0) {
            if (mean_out) { mean_out[row] = mu; rstd_out[row] = rs; }
            if (m_out) m_out[row] = vt_mask ? vt_mask[row] : 1;
        }
    }
}

// backward, row part: dpre = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.  dpre goes to the fp32
// workspace (all rows; the embedding gradients are column sums of row subsets of it) and, for p > 0, to df in the
// feature dtype.  dgamma / dbeta: every warp accumulates into its PRIVATE shared-memory slice acc[warp][2][C] (float4
// read-modify-write, lanes on consecutive vectors, no synchronisation) -- register accumulators had put the kernel at 171
// registers = one 8-warp block per SM (12.5 % warps active, 85 us); the slices are combined in fixed order at the end,
// one partial per block.
template <typename T, typename TDY, int KCH>
__global__ void __launch_bounds__(256, kTailBwdOcc) enc_tail_bwd_kernel(
    const TDY* __restrict__ dy, const T* __restrict__ f, const float* __restrict__ cls, const float* __restrict__ pos,
    const float* __restrict__ len, const float* __restrict__ odr_emb, const int32_t* __restrict__ odr,
    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ df,
    float* __restrict__ dpre, float* __restrict__ part, int B, int Tn, int hw, int C) {
    extern __shared__ __align__(16) float acc[];          // [8 warps][2][C]
    const int P = hw + 1;
    const long long nrows = (long long)B * Tn * P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const float inv_c = 1.0f / (float)C;
    float* acc_g = acc + (size_t)warp * 2 * C;
    float* acc_b = acc_g + C;
    for (int col = lane * 4; col < C; col += 128) {
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        store4<float>(acc_g + col, z);
        store4<float>(acc_b + col, z);
    }
    __syncwarp();

    for (long long row = warp0; row < nrows; row += wstride) {
        const long long bt = row / P;
        const int p = (int)(row - bt * P);
        const int t = (int)(bt % Tn);
        const float* sel = (odr && odr[bt] != t) ? odr_emb : len + (long long)t * C;
        float xh[KCH][4], g[KCH][4];
        (void)tail_row<T, KCH>(f, cls, pos, sel, bt, p, hw, C, lane, xh);
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            const int col = (lane + 32 * k) * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) g[k][e] = 0.f;
            if (col < C) {
                float d4[4], g4[4], a4[4], b4[4];
                load4<TDY>(dy + row * C + col, d4);
                load4<float>(gamma + col, g4);
                load4<float>(acc_g + col, a4);
                load4<float>(acc_b + col, b4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xv = (xh[k][e] - mu) * rs;
                    xh[k][e] = xv;
                    a4[e] = fmaf(d4[e], xv, a4[e]);
                    b4[e] += d4[e];
                    const float gg = d4[e] * g4[e];
                    g[k][e] = gg;
                    s1 += gg;
                    s2 += gg * xv;
                }
                store4<float>(acc_g + col, a4);
                store4<float>(acc_b + col, b4);
            }
        }
        const float c1 = warp_sum(s1) * inv_c, c2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            const int col = (lane + 32 * k) * 4;
            if (col < C) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = rs * (g[k][e] - c1 - xh[k][e] * c2);
                store4<float>(dpre + row * C + col, o);
                if (df && p > 0) store4<T>(df + (bt * hw + p - 1) * (long long)C + col, o);
            }
        }
    }
    // block partial of dgamma (pass 0) / dbeta (pass 1): the 8 warp slices combined in fixed order
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int pass = i / C, c = i - pass * C;
        float tsum = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tsum += acc[((size_t)w8 * 2 + pass) * C + c];
        part[((size_t)blockIdx.x * 2 + pass) * C + c] = tsum;
    }
}

// Fixed-order column sums over row subsets.  blockDim (32, 16): 32 columns x 16 row lanes, smem combine in ty order.
//   STAGE 0 (grid.x = pos_rows + B*Tn):  j <  pos_rows : dpos[j] = sum_bt dpre[bt, j]  (0 for j >= P); j == 0 also -> dcls
//                                        j >= pos_rows : S[bt]   = sum_p  dpre[bt, p]
//   STAGE 1 (grid.x = len_rows + 3):     j <  len_rows : dlen[j] = sum_b [odr[b,j] == j] S[b,j]  (0 for j >= Tn)
//                                        j == len_rows : dodr    = sum_bt [odr[bt] != t] S[bt]
//                                        j == len_rows + 1 / + 2 : dgamma / dbeta = sum_blocks part[block][0 / 1]
template <int STAGE>
__global__ void __launch_bounds__(512) enc_tail_reduce_kernel(const float* __restrict__ dpre, float* __restrict__ S,
                                                              const float* __restrict__ part, int nparts,
                                                              const int32_t* __restrict__ odr, float* __restrict__ dcls,
                                                              float* __restrict__ dpos, float* __restrict__ dlen,
                                                              float* __restrict__ dodr, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int B, int Tn, int hw, int C,
                                                              int pos_rows, int len_rows) {
    const int P = hw + 1;
    const int j = blockIdx.x;
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int ty = threadIdx.y, TY = blockDim.y;
    const bool col_ok = c < C;
    float acc = 0.f;
    float* dst = nullptr;
    float* dst2 = nullptr;
    if (STAGE == 0) {
        if (j < pos_rows) {
            dst = dpos + (size_t)j * C;
            if (j == 0) dst2 = dcls;
            if (j < P && col_ok)
                for (int bt = ty; bt < B * Tn; bt += TY) acc += dpre[((size_t)bt * P + j) * C + c];
        } else {
            const int bt = j - pos_rows;
            dst = S + (size_t)bt * C;
            if (col_ok)
                for (int p = ty; p < P; p += TY) acc += dpre[((size_t)bt * P + p) * C + c];
        }
    } else {
        if (j < len_rows) {
            dst = dlen + (size_t)j * C;
            if (j < Tn && col_ok)
                for (int b = ty; b < B; b += TY) {
                    const int bt = b * Tn + j;
                    if (!odr || odr[bt] == j) acc += S[(size_t)bt * C + c];
                }
        } else if (j == len_rows) {
            dst = dodr;
            if (odr && col_ok)
                for (int bt = ty; bt < B * Tn; bt += TY)
                    if (odr[bt] != bt % Tn) acc += S[(size_t)bt * C + c];
        } else {
            const int pass = j - len_rows - 1;
            dst = pass == 0 ? dgamma : dbeta;
            if (col_ok)
                for (int s = ty; s < nparts; s += TY) acc += part[((size_t)s * 2 + pass) * C + c];
        }
    }
    __shared__ float sm[16][33];
    sm[ty][threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0 && col_ok) {
        float tsum = 0.f;
        for (int y = 0; y < TY; ++y) tsum += sm[y][threadIdx.x];
        if (dst) dst[c] = tsum;
        if (dst2) dst2[c] = tsum;
    }
}

static int tail_bwd_grid(long long nrows) {
    long long g = (nrows + 8 * 4 - 1) / (8 * 4);   // >= 4 rows per warp
    if (g > kTailMaxBlocks) g = kTailMaxBlocks;
    return (int)(g < 1 ? 1 : g);
}

template <typename T, typename TO>
static int launch_tail_fwd(const void* f, const float* cls, const float* pos, const float* len, const float* odr_emb,
                           const int32_t* odr, const float* gamma, const float* beta, const int64_t* vt_mask, void* out,
                           int64_t* m_out, float* mean, float* rstd, int B, int Tn, int hw, int C, float eps,
                           cudaStream_t st) {
    const long long nrows = (long long)B * Tn * (hw + 1);
    long long grid = (nrows + 7) / 8;
    if (grid > (long long)kNumSMs * 16) grid = (long long)kNumSMs * 16;
#define VSW_TAIL_FWD(K)                                                                                              \
    enc_tail_fwd_kernel<T, TO, K><<<(int)grid, 256, 0, st>>>((const T*)f, cls, pos, len, odr_emb, odr, gamma, beta, vt_mask, \
                                                             (TO*)out, m_out, mean, rstd, B, Tn, hw, C, eps)
    if (C <= 256) VSW_TAIL_FWD(2);
    else if (C <= 512) VSW_TAIL_FWD(4);
    else if (C <= 768) VSW_TAIL_FWD(6);
    else VSW_TAIL_FWD(8);
#undef VSW_TAIL_FWD
    return check_launch("enc_video_tail_fwd");
}

template <typename T, typename TDY>
static int launch_tail_bwd(const void* dy, const void* f, const float* cls, const float* pos, const float* len,
                           const float* odr_emb, const int32_t* odr, const float* gamma, const float* mean,
                           const float* rstd, void* df, float* dcls, float* dpos, float* dlen, float* dodr,
                           float* dgamma, float* dbeta, int B, int Tn, int hw, int C, int pos_rows, int len_rows,
                           float* ws, cudaStream_t st) {
    const int P = hw + 1;
    const long long nrows = (long long)B * Tn * P;
    float* dpre = ws;
    float* S = dpre + (size_t)nrows * C;
    float* part = S + (size_t)B * Tn * C;
    const int grid = tail_bwd_grid(nrows);
    const size_t smem = (size_t)8 * 2 * C * sizeof(float);   // <= 64 KB at C = 1024: two blocks per SM fit
#define VSW_TAIL_BWD(K)                                                                                              \
    do {                                                                                                             \
        auto kern = enc_tail_bwd_kernel<T, TDY, K>;                                                                  \
        if (smem > 48 * 1024 &&                                                                                      \
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {     \
            set_error("enc_video_tail_bwd: cannot reserve %zu bytes of shared memory", smem);                        \
            return VSW_ERR_CUDA;                                                                                     \
        }                                                                                                            \
        kern<<<grid, 256, smem, st>>>((const TDY*)dy, (const T*)f, cls, pos, len, odr_emb, odr, gamma, mean, rstd,   \
                                      (T*)df, dpre, part, B, Tn, hw, C);                                             \
    } while (0)
    if (C <= 256) VSW_TAIL_BWD(2);
    else if (C <= 512) VSW_TAIL_BWD(4);
    else if (C <= 768) VSW_TAIL_BWD(6);
    else VSW_TAIL_BWD(8);
#undef VSW_TAIL_BWD
    int rc = check_launch("enc_video_tail_bwd");
    if (rc) return rc;
    const dim3 blk(32, 16);
    enc_tail_reduce_kernel<0><<<dim3(pos_rows + B * Tn, ceil_div(C, 32)), blk, 0, st>>>(
        dpre, S, part, grid, odr, dcls, dpos, dlen, dodr, dgamma, dbeta, B, Tn, hw, C, pos_rows, len_rows);
    rc = check_launch("enc_video_tail_reduce0");
    if (rc) return rc;
    enc_tail_reduce_kernel<1><<<dim3(len_rows + 3, ceil_div(C, 32)), blk, 0, st>>>(
        dpre, S, part, grid, odr, dcls, dpos, dlen, dodr, dgamma, dbeta, B, Tn, hw, C, pos_rows, len_rows);
    return check_launch("enc_video_tail_reduce1");
}

}  // namespace vsw

using namespace vsw;

#define VSW_TAIL_COMMON_CHECKS(name)                                                                                   \
    VSW_REQUIRE(B > 0 && Tn > 0 && hw > 0 && C > 0, VSW_ERR_ARG, name ": bad sizes");                                   \
    VSW_REQUIRE((C % 4) == 0 && C <= 32 * 4 * kTailMaxK, VSW_ERR_UNSUPPORTED,                                           \
                name ": hidden size %d must be a multiple of 4 and <= %d", C, 32 * 4 * kTailMaxK);                      \
    VSW_REQUIRE(pos_rows >= hw + 1, VSW_ERR_ARG, name ": emb_pos has %d rows, needs %d (1 + h*w)", pos_rows, hw + 1);   \
    VSW_REQUIRE(len_rows >= Tn, VSW_ERR_ARG, name ": emb_len has %d rows, needs %d frames", len_rows, Tn)

extern "C" int vsw_enc_video_tail_fwd(const void* f, const float* emb_cls, const float* emb_pos, const float* emb_len,
                                      const float* emb_odr, const int32_t* odr, const float* gamma, const float* beta,
                                      const int64_t* vt_mask, void* out, int64_t* m_img, float* mean, float* rstd, int B,
                                      int Tn, int hw, int C, int pos_rows, int len_rows, float eps, int dtype,
                                      int out_dtype, void* stream) {
    VSW_REQUIRE(f && emb_cls && emb_pos && emb_len && gamma && beta && out, VSW_ERR_ARG, "vsw_enc_video_tail_fwd: null pointer");
    VSW_REQUIRE(!odr || emb_odr, VSW_ERR_ARG, "vsw_enc_video_tail_fwd: odr given without emb_odr");
    VSW_REQUIRE((mean == nullptr) == (rstd == nullptr), VSW_ERR_ARG, "vsw_enc_video_tail_fwd: mean/rstd both or none");
    VSW_TAIL_COMMON_CHECKS("vsw_enc_video_tail_fwd");
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == dtype) {
        VSW_DISPATCH_DTYPE(dtype, T, return (launch_tail_fwd<T, T>(f, emb_cls, emb_pos, emb_len, emb_odr, odr, gamma, beta,
                                                                  vt_mask, out, m_img, mean, rstd, B, Tn, hw, C, eps, st)));
    }
    VSW_REQUIRE(out_dtype == VSW_F32, VSW_ERR_DTYPE, "vsw_enc_video_tail_fwd: out_dtype must equal dtype or be fp32");
    VSW_DISPATCH_DTYPE(dtype, T, return (launch_tail_fwd<T, float>(f, emb_cls, emb_pos, emb_len, emb_odr, odr, gamma, beta,
                                                                  vt_mask, out, m_img, mean, rstd, B, Tn, hw, C, eps, st)));
    return VSW_OK;
}

extern "C" size_t vsw_enc_video_tail_bwd_workspace(int B, int Tn, int hw, int C) {
    if (B <= 0 || Tn <= 0 || hw <= 0 || C <= 0) return 0;
    const size_t nrows = (size_t)B * Tn * (hw + 1);
    return (nrows * C + (size_t)B * Tn * C + (size_t)kTailMaxBlocks * 2 * C) * sizeof(float);
}

extern "C" int vsw_enc_video_tail_bwd(const void* dy, const void* f, const float* emb_cls, const float* emb_pos,
                                      const float* emb_len, const float* emb_odr, const int32_t* odr, const float* gamma,
                                      const float* mean, const float* rstd, void* df, float* demb_cls, float* demb_pos,
                                      float* demb_len, float* demb_odr, float* dgamma, float* dbeta, int B, int Tn, int hw,
                                      int C, int pos_rows, int len_rows, int dtype, int dy_dtype, void* ws,
                                      size_t ws_bytes, void* stream) {
    VSW_REQUIRE(dy && f && emb_cls && emb_pos && emb_len && gamma && mean && rstd, VSW_ERR_ARG,
                "vsw_enc_video_tail_bwd: null pointer");
    VSW_REQUIRE(demb_cls && demb_pos && demb_len && demb_odr && dgamma && dbeta, VSW_ERR_ARG,
                "vsw_enc_video_tail_bwd: every parameter-gradient pointer is required");
    VSW_REQUIRE(!odr || emb_odr, VSW_ERR_ARG, "vsw_enc_video_tail_bwd: odr given without emb_odr");
    VSW_TAIL_COMMON_CHECKS("vsw_enc_video_tail_bwd");
    const size_t need = vsw_enc_video_tail_bwd_workspace(B, Tn, hw, C);
    VSW_REQUIRE(ws && ws_bytes >= need, VSW_ERR_WORKSPACE, "vsw_enc_video_tail_bwd: workspace %zu < %zu", ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (dy_dtype == dtype) {
        VSW_DISPATCH_DTYPE(dtype, T, return (launch_tail_bwd<T, T>(dy, f, emb_cls, emb_pos, emb_len, emb_odr, odr, gamma, mean,
                                                                  rstd, df, demb_cls, demb_pos, demb_len, demb_odr, dgamma,
                                                                  dbeta, B, Tn, hw, C, pos_rows, len_rows, (float*)ws, st)));
    }
    VSW_REQUIRE(dy_dtype == VSW_F32, VSW_ERR_DTYPE, "vsw_enc_video_tail_bwd: dy_dtype must equal dtype or be fp32");
    VSW_DISPATCH_DTYPE(dtype, T, return (launch_tail_bwd<T, float>(dy, f, emb_cls, emb_pos, emb_len, emb_odr, odr, gamma, mean,
                                                                  rstd, df, demb_cls, demb_pos, demb_len, demb_odr, dgamma,
                                                                  dbeta, B, Tn, hw, C, pos_rows, len_rows, (float*)ws, st)));
    return VSW_OK;
}
