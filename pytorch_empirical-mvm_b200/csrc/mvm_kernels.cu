// MVM (masked visual modelling) kernels on either side of the Swin encoder (the parts that
// touch the encoder's input and output):
//   vsw_block_mask_apply   main_pretrain.py:355-362 -- zero the masked 32x32 patches of the clip that goes INTO the student
//                          encoder (`img[i] *= 1 - cov`) and, optionally, materialise the full-resolution `mvm_mask`;
//   vsw_masked_l1_fwd/bwd  main_pretrain.py:520-522 (3d_feature; same formula :427-428 pixel, :534-535 2d_feature) -- the
//                          masked L1 between the prediction and the TEACHER encoder's output tokens:
//                              loss = sum_{r,c} |pred[r,c] - target[r,c]| * m[r] / (sum_r m[r] + 1e-5) / in_c
// Both are HBM-bound element passes: 16-byte vector accesses, grid sized to the SM count; the loss only reads rows whose
// weight is non-zero (about 1 - p_mask of the rows are skipped).  Deterministic: fixed-order two-stage reduction.
#include "common.cuh"

namespace vsw {

// ------------------------------------------------------------------------------------------------------------------
// patch mask: thread = one 16-byte vector along W (never straddles a patch: ps % VN == 0, W % VN == 0)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) block_mask_apply_kernel(const T* __restrict__ img, const uint8_t* __restrict__ cov,
                                                               T* __restrict__ out, float* __restrict__ mask,
                                                               long long nvec, int Cin, int H, int W, int ps) {
    constexpr int VN = Vec16<T>::N;
    const int h = H / ps, w = W / ps;
    const int vec_per_row = W / VN;
    for (long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec; vi += (long long)gridDim.x * blockDim.x) {
        const int xv = (int)(vi % vec_per_row);
        const long long rowi = vi / vec_per_row;          // (frame, channel, y)
        const int y = (int)(rowi % H);
        const long long frame = rowi / ((long long)H * Cin);
        const float c = (float)cov[(frame * h + y / ps) * w + (xv * VN) / ps];
        const float keep = 1.0f - c;                        // the reference multiplies: img *= (1.0 - cov)
        if (out) {
            float v[VN];
            load_vec<T>(img + vi * VN, v);
#pragma unroll
            for (int e = 0; e < VN; ++e) v[e] *= keep;
            store_vec<T>(out + vi * VN, v);
        }
        if (mask) {
#pragma unroll
            for (int q = 0; q < VN / 4; ++q) {
                const float m4[4] = {c, c, c, c};
                store_vec<float>(mask + vi * VN + q * 4, m4);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// masked L1
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ld4(const T* __restrict__ p, float (&o)[4]) {
    Pack<T, 4> pk = *reinterpret_cast<const Pack<T, 4>*>(p);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = to_f<T>(pk.v[e]);
}

constexpr int kL1MaxBlocks = 148 * 4;

// one warp per row; part[block] = (sum |d| * m, sum m) over the rows of the block, combined in fixed order
template <typename T, typename TT>
__global__ void __launch_bounds__(256) masked_l1_partial_kernel(const T* __restrict__ pred, const TT* __restrict__ target,
                                                                const float* __restrict__ m, float2* __restrict__ part,
                                                                long long rows, int C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * 8 + warp, wstride = (long long)gridDim.x * 8;
    float acc = 0.f, macc = 0.f;
    for (long long r = warp0; r < rows; r += wstride) {
        const float mr = m[r];
        macc += mr;
        if (mr == 0.f) continue;                           // masked-out rows are never read
        float s = 0.f;
        for (int col = lane * 4; col < C; col += 128) {
            float a[4], b[4];
            ld4<T>(pred + r * C + col, a);
            ld4<TT>(target + r * C + col, b);
#pragma unroll
            for (int e = 0; e < 4; ++e) s += fabsf(a[e] - b[e]);
        }
        acc += warp_sum(s) * mr;
    }
    __shared__ float2 red[8];
    if (lane == 0) red[warp] = make_float2(acc, macc);
    __syncthreads();
    if (threadIdx.x == 0) {
        float2 t = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { t.x += red[i].x; t.y += red[i].y; }
        part[blockIdx.x] = t;
    }
}

__global__ void masked_l1_finish_kernel(const float2* __restrict__ part, int nparts, float in_c, float* __restrict__ loss,
                                        float* __restrict__ msum) {
    // one warp, lane-strided partial sums then a shuffle tree: fixed order
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 32) { a += part[i].x; b += part[i].y; }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
        *loss = a / (b + 1e-5f) / in_c;
        *msum = b;
    }
}

template <typename T, typename TT>
__global__ void __launch_bounds__(256) masked_l1_bwd_kernel(const T* __restrict__ pred, const TT* __restrict__ target,
                                                            const float* __restrict__ m, const float* __restrict__ msum,
                                                            const float* __restrict__ dloss, T* __restrict__ dpred,
                                                            long long rows, int C, float in_c) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
    const float k = dloss[0] / (msum[0] + 1e-5f) / in_c;
    for (long long r = warp0; r < rows; r += wstride) {
        const float g = m[r] * k;
        for (int col = lane * 4; col < C; col += 128) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            if (g != 0.f) {
                float a[4], b[4];
                ld4<T>(pred + r * C + col, a);
                ld4<TT>(target + r * C + col, b);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = a[e] - b[e];
                    o[e] = d > 0.f ? g : (d < 0.f ? -g : 0.f);     // sign(0) = 0, as torch's l1_loss backward
                }
            }
            Pack<T, 4> pk;
#pragma unroll
            for (int e = 0; e < 4; ++e) pk.v[e] = from_f<T>(o[e]);
            *reinterpret_cast<Pack<T, 4>*>(dpred + r * C + col) = pk;
        }
    }
}

static int l1_grid(long long rows) {
    long long g = (rows + 7) / 8;
    if (g > kL1MaxBlocks) g = kL1MaxBlocks;
    return (int)(g < 1 ? 1 : g);
}

template <typename T, typename TT>
static int launch_l1_fwd(const void* pred, const void* target, const float* m, float* loss, float* msum, void* ws,
                         long long rows, int C, float in_c, cudaStream_t st) {
    const int grid = l1_grid(rows);
    masked_l1_partial_kernel<T, TT><<<grid, 256, 0, st>>>((const T*)pred, (const TT*)target, m, (float2*)ws, rows, C);
    int rc = check_launch("masked_l1_partial");
    if (rc) return rc;
    masked_l1_finish_kernel<<<1, 32, 0, st>>>((const float2*)ws, grid, in_c, loss, msum);
    return check_launch("masked_l1_finish");
}

template <typename T, typename TT>
static int launch_l1_bwd(const void* pred, const void* target, const float* m, const float* msum, const float* dloss,
                         void* dpred, long long rows, int C, float in_c, cudaStream_t st) {
    long long grid = (rows + 7) / 8;
    if (grid > (long long)kNumSMs * 8) grid = (long long)kNumSMs * 8;
    masked_l1_bwd_kernel<T, TT><<<(int)grid, 256, 0, st>>>((const T*)pred, (const TT*)target, m, msum, dloss, (T*)dpred, rows,
                                                           C, in_c);
    return check_launch("masked_l1_bwd");
}

}  // namespace vsw

using namespace vsw;

extern "C" int vsw_block_mask_apply(const void* img, const uint8_t* cov, void* img_out, float* mvm_mask, int frames,
                                    int Cin, int H, int W, int ps, int dtype, void* stream) {
    VSW_REQUIRE(cov && frames > 0 && Cin > 0 && H > 0 && W > 0 && ps > 0, VSW_ERR_ARG, "vsw_block_mask_apply: bad args");
    VSW_REQUIRE((img && img_out) || (!img_out && mvm_mask), VSW_ERR_ARG,
                "vsw_block_mask_apply: give img + img_out (may alias) and/or mvm_mask");
    VSW_REQUIRE(H % ps == 0 && W % ps == 0, VSW_ERR_ARG,
                "vsw_block_mask_apply: frame %dx%d is not a multiple of the patch size %d", H, W, ps);
    cudaStream_t st = (cudaStream_t)stream;
    VSW_DISPATCH_DTYPE(dtype, T, {
        constexpr int VN = Vec16<T>::N;
        VSW_REQUIRE(ps % VN == 0, VSW_ERR_UNSUPPORTED, "vsw_block_mask_apply: patch size %d must be a multiple of %d", ps, VN);
        const long long nvec = (long long)frames * Cin * H * W / VN;
        long long grid = (nvec + 255) / 256;
        if (grid > (long long)kNumSMs * 16) grid = (long long)kNumSMs * 16;
        block_mask_apply_kernel<T><<<(int)grid, 256, 0, st>>>((const T*)img, cov, (T*)img_out, mvm_mask, nvec, Cin, H, W, ps);
        return check_launch("block_mask_apply");
    });
    return VSW_OK;
}

extern "C" size_t vsw_masked_l1_workspace(void) { return (size_t)kL1MaxBlocks * sizeof(float2); }

#define VSW_L1_DISPATCH(fn, ...)                                                                                \
    if (target_dtype == dtype) { VSW_DISPATCH_DTYPE(dtype, T, return (fn<T, T>(__VA_ARGS__))); }                \
    VSW_REQUIRE(target_dtype == VSW_F32, VSW_ERR_DTYPE, "masked_l1: target dtype must equal dtype or be fp32"); \
    VSW_DISPATCH_DTYPE(dtype, T, return (fn<T, float>(__VA_ARGS__)))

extern "C" int vsw_masked_l1_fwd(const void* pred, const void* target, const float* row_weight, float* loss, float* msum,
                                 long long rows, int C, float in_c, int dtype, int target_dtype, void* ws, size_t ws_bytes,
                                 void* stream) {
    VSW_REQUIRE(pred && target && row_weight && loss && msum && rows > 0 && C > 0 && in_c > 0.f, VSW_ERR_ARG,
                "vsw_masked_l1_fwd: bad args");
    VSW_REQUIRE(C % 4 == 0, VSW_ERR_UNSUPPORTED, "vsw_masked_l1_fwd: C=%d must be a multiple of 4", C);
    VSW_REQUIRE(ws && ws_bytes >= vsw_masked_l1_workspace(), VSW_ERR_WORKSPACE, "vsw_masked_l1_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    VSW_L1_DISPATCH(launch_l1_fwd, pred, target, row_weight, loss, msum, ws, rows, C, in_c, st);
    return VSW_OK;
}

extern "C" int vsw_masked_l1_bwd(const void* pred, const void* target, const float* row_weight, const float* msum,
                                 const float* dloss, void* dpred, long long rows, int C, float in_c, int dtype,
                                 int target_dtype, void* stream) {
    VSW_REQUIRE(pred && target && row_weight && msum && dloss && dpred && rows > 0 && C > 0 && in_c > 0.f, VSW_ERR_ARG,
                "vsw_masked_l1_bwd: bad args");
    VSW_REQUIRE(C % 4 == 0, VSW_ERR_UNSUPPORTED, "vsw_masked_l1_bwd: C=%d must be a multiple of 4", C);
    cudaStream_t st = (cudaStream_t)stream;
    VSW_L1_DISPATCH(launch_l1_bwd, pred, target, row_weight, msum, dloss, dpred, rows, C, in_c, st);
    return VSW_OK;
}
