// Integer index maps of the shifted-window machinery.  Bit-exact with the reference
// (visbackbone/video_swin.py:84-93 window_partition/reverse, :220-241 roll, :292-307 compute_mask,
//  :123-137 relative_position_index, :276-284 PatchMerging slicing).  Closed forms: SURVEY Appendix A.
#include "common.cuh"
#include "geometry.cuh"

namespace vsw {

__global__ void window_maps_kernel(WinGeom g, int32_t* __restrict__ gather, uint8_t* __restrict__ region) {
    const int total = g.nW * g.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int w = i / g.N, n = i - w * g.N;
        if (gather) gather[i] = g.source_token(w, n);
        if (region) region[i] = (uint8_t)g.region_id(w, n);
    }
}

__global__ void rel_pos_index_kernel(int wd, int wh, int ww, int64_t* __restrict__ out) {
    const int N = wd * wh * ww;
    const long long total = (long long)N * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / N), j = (int)(t - (long long)i * N);
        const int di = i / (wh * ww), hi = (i / ww) % wh, wi = i % ww;
        const int dj = j / (wh * ww), hj = (j / ww) % wh, wj = j % ww;
        out[t] = (long long)(di - dj + wd - 1) * (2 * wh - 1) * (2 * ww - 1) +
                 (long long)(hi - hj + wh - 1) * (2 * ww - 1) + (wi - wj + ww - 1);
    }
}

template <typename T>
__global__ void shift_mask_kernel(const uint8_t* __restrict__ region, int nW, int N, T* __restrict__ out) {
    const long long total = (long long)nW * N * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % N);
        const long long wi = t / N;           // w*N + i
        const int w = (int)(wi / N);
        const uint8_t ri = region[wi], rj = region[(long long)w * N + j];
        out[t] = from_f<T>(ri != rj ? -100.0f : 0.0f);
    }
}

__global__ void merge_map_kernel(int D, int H, int W, int32_t* __restrict__ out) {
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const int total = D * H2 * W2 * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i & 3, r = i >> 2;
        const int w2 = r % W2, h2 = (r / W2) % H2, d = r / (W2 * H2);
        const int h = 2 * h2 + (g & 1), w = 2 * w2 + (g >> 1);  // g: (dh,dw) = (0,0),(1,0),(0,1),(1,1)
        out[i] = (h < H && w < W) ? (d * H + h) * W + w : -1;
    }
}

}  // namespace vsw

using namespace vsw;

extern "C" int vsw_window_maps(int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw,
                               int32_t* gather, uint8_t* region, void* stream) {
    WinGeom g;
    VSW_REQUIRE(make_win_geom(D, H, W, wd, wh, ww, sd, sh, sw, &g), VSW_ERR_ARG,
                "vsw_window_maps: bad geometry grid(%d,%d,%d) window(%d,%d,%d) shift(%d,%d,%d)", D, H, W, wd, wh, ww,
                sd, sh, sw);
    const int total = g.nW * g.N;
    window_maps_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(g, gather, region);
    return check_launch("vsw_window_maps");
}

extern "C" int vsw_rel_pos_index(int wd, int wh, int ww, int64_t* out, void* stream) {
    VSW_REQUIRE(wd > 0 && wh > 0 && ww > 0 && out, VSW_ERR_ARG, "vsw_rel_pos_index: bad args");
    const long long total = (long long)wd * wh * ww * wd * wh * ww;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    rel_pos_index_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(wd, wh, ww, out);
    return check_launch("vsw_rel_pos_index");
}

extern "C" int vsw_shift_mask(const uint8_t* region, int nW, int N, void* out, int dtype, void* stream) {
    VSW_REQUIRE(region && out && nW > 0 && N > 0, VSW_ERR_ARG, "vsw_shift_mask: bad args");
    const long long total = (long long)nW * N * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    VSW_DISPATCH_DTYPE(dtype, T,
                       (shift_mask_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(region, nW, N, (T*)out)));
    return check_launch("vsw_shift_mask");
}

extern "C" int vsw_merge_map(int D, int H, int W, int32_t* out, void* stream) {
    VSW_REQUIRE(D > 0 && H > 0 && W > 0 && out, VSW_ERR_ARG, "vsw_merge_map: bad args");
    const int total = D * ((H + 1) / 2) * ((W + 1) / 2) * 4;
    merge_map_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(D, H, W, out);
    return check_launch("vsw_merge_map");
}

// ---------------------------------------------------------------------------------------------
// residual add with window-reverse scatter:  out[b, map[r], :] = x[b, map[r], :] + scale[b] * y[b, r, :]
// (x / out in TX, the branch output y in TY).  Used when the residual stream is kept in a wider type than the branch
// (torch.autocast: fp32 stream, 16-bit Linear outputs -- video_swin.py:256, 261 under autocast promote to fp32).
// ---------------------------------------------------------------------------------------------
namespace vsw {
template <typename TX, typename TY>
__global__ void __launch_bounds__(256) residual_add_kernel(const TX* __restrict__ x, const TY* __restrict__ y,
                                                           const int32_t* __restrict__ map, const float* __restrict__ scale,
                                                           TX* __restrict__ out, int B, int R, int T, int C) {
    const int vpr = C / 4;
    const long long total = (long long)B * R * vpr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / vpr;
        const int v = (int)(i - row * vpr);
        const int b = (int)(row / R), r = (int)(row - (long long)b * R);
        const int t = map ? __ldg(map + r) : r;
        if (t < 0) continue;
        const float sc = scale ? __ldg(scale + b) : 1.f;
        const TY* yp = y + row * C + v * 4;
        const long long o = ((long long)b * T + t) * C + v * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) out[o + e] = from_f<TX>(fmaf(sc, to_f<TY>(yp[e]), to_f<TX>(x[o + e])));
    }
}
}  // namespace vsw

extern "C" int vsw_residual_add(const void* x, const void* y, const int32_t* rowmap, const float* rowscale, void* out, int B,
                                int rows_per_batch, int dst_rows_per_batch, int C, int dtype, int y_dtype, void* stream) {
    using namespace vsw;
    VSW_REQUIRE(x && y && out && B > 0 && rows_per_batch > 0 && dst_rows_per_batch > 0 && C > 0 && C % 4 == 0, VSW_ERR_ARG,
                "vsw_residual_add: bad args");
    VSW_REQUIRE(rowmap || rows_per_batch == dst_rows_per_batch, VSW_ERR_ARG, "vsw_residual_add: identity map needs equal row counts");
    const long long total = (long long)B * rows_per_batch * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
    cudaStream_t st = (cudaStream_t)stream;
    VSW_DISPATCH_DTYPE(dtype, TX,
        VSW_DISPATCH_DTYPE(y_dtype, TY,
            (residual_add_kernel<TX, TY><<<(int)blocks, 256, 0, st>>>((const TX*)x, (const TY*)y, rowmap, rowscale, (TX*)out, B,
                                                                     rows_per_batch, dst_rows_per_batch, C))));
    return check_launch("residual_add");
}

// ---------------------------------------------------------------------------------------------
// stochastic-depth factors (video_swin.py:46-54):  out[b] = floor(keep + u[b]) / keep, the sum rounded to the dtype of u
// exactly as `keep_prob + torch.rand(shape, dtype=x.dtype)` rounds it.  One launch instead of add / floor / cast / div.
// ---------------------------------------------------------------------------------------------
namespace vsw {
template <typename T>
__global__ void drop_path_scale_kernel(const T* __restrict__ u, float keep, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = floorf(to_f<T>(from_f<T>(to_f<T>(u[i]) + keep))) / keep;
}
}  // namespace vsw

extern "C" int vsw_drop_path_scale(const void* u, float keep, float* out, int n, int dtype, void* stream) {
    using namespace vsw;
    VSW_REQUIRE(u && out && n > 0 && keep > 0.f, VSW_ERR_ARG, "vsw_drop_path_scale: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    VSW_DISPATCH_DTYPE(dtype, T, (drop_path_scale_kernel<T><<<(n + 127) / 128, 128, 0, st>>>((const T*)u, keep, out, n)));
    return check_launch("drop_path_scale");
}
