// PatchEmbed3D gather kernels (visbackbone/video_swin.py:390-400): im2col / col2im for the
// Conv3d(k=(pd,ph,pw), stride=(1,ph,pw)) with right/bottom zero padding and one appended zero frame.
// HBM-bound: consecutive threads walk (dx, token) so both the clip reads and the col writes coalesce.
#include "common.cuh"

namespace vsw {

template <typename TX, typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const TX* __restrict__ x, T* __restrict__ col, int B, int Cin,
                                                     int D, int H, int W, int pd, int ph, int pw, int Dout, int Hp,
                                                     int Wp) {
    // one block per (b, d, h') row of tokens; smem tile [Wp][Kvol] so that global reads run along w and
    // global writes run along the col row
    extern __shared__ __align__(16) unsigned char smraw[];
    T* tile = reinterpret_cast<T*>(smraw);
    const int Kvol = Cin * pd * ph * pw;
    const int rowid = blockIdx.x;  // over B*Dout*Hp
    const int hp = rowid % Hp, d = (rowid / Hp) % Dout, b = rowid / (Hp * Dout);
    const int wspan = Wp * pw;
    const int outer = Cin * pd * ph;
    for (int idx = threadIdx.x; idx < outer * wspan; idx += blockDim.x) {
        const int wx = idx % wspan;          // pixel column
        const int o = idx / wspan;           // (c, dt, dy)
        const int dy = o % ph, dt = (o / ph) % pd, c = o / (ph * pd);
        const int f = d + dt, yy = hp * ph + dy;
        float v = 0.f;
        if (f < D && yy < H && wx < W) v = to_f<TX>(x[(((long long)b * Cin + c) * D + f) * H * W + (long long)yy * W + wx]);
        const int wp = wx / pw, dx = wx % pw;
        tile[wp * Kvol + o * pw + dx] = from_f<T>(v);
    }
    __syncthreads();
    T* dst = col + (long long)rowid * Wp * Kvol;
    for (int idx = threadIdx.x; idx < Wp * Kvol; idx += blockDim.x) dst[idx] = tile[idx];
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, TX* __restrict__ dx, int B, int Cin,
                                                     int D, int H, int W, int pd, int ph, int pw, int Dout, int Hp,
                                                     int Wp) {
    const int Kvol = Cin * pd * ph * pw;
    const long long total = (long long)B * Cin * D * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int wx = (int)(i % W);
        const int yy = (int)((i / W) % H);
        const int f = (int)((i / ((long long)W * H)) % D);
        const int c = (int)((i / ((long long)W * H * D)) % Cin);
        const int b = (int)(i / ((long long)W * H * D * Cin));
        const int wp = wx / pw, dxx = wx % pw, hp = yy / ph, dy = yy % ph;
        float s = 0.f;
        for (int dt = 0; dt < pd; ++dt) {
            const int d = f - dt;
            if (d < 0 || d >= Dout) continue;
            const long long row = (((long long)b * Dout + d) * Hp + hp) * Wp + wp;
            s += to_f<T>(dcol[row * Kvol + ((c * pd + dt) * ph + dy) * pw + dxx]);
        }
        dx[i] = from_f<TX>(s);
    }
}

}  // namespace vsw
using namespace vsw;

#define VSW_PE_DISPATCH(x_dtype, dtype, ...)                                            \
    VSW_DISPATCH_DTYPE(x_dtype, TX, VSW_DISPATCH_DTYPE(dtype, T, __VA_ARGS__))

extern "C" int vsw_patch_im2col(const void* x, void* col, int B, int Cin, int D, int H, int W, int pd, int ph, int pw,
                                int x_dtype, int dtype, void* stream) {
    VSW_REQUIRE(x && col && B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0 && pd > 0 && ph > 0 && pw > 0, VSW_ERR_ARG,
                "vsw_patch_im2col: bad args");
    const int Dout = D + 2 - pd, Hp = (H + ph - 1) / ph, Wp = (W + pw - 1) / pw;
    VSW_REQUIRE(Dout > 0, VSW_ERR_ARG, "vsw_patch_im2col: D=%d too short for temporal patch %d", D, pd);
    const int Kvol = Cin * pd * ph * pw;
    const size_t esz = dtype == VSW_F32 ? 4 : 2;
    const size_t smem = (size_t)Wp * Kvol * esz;
    VSW_REQUIRE(smem <= 200 * 1024, VSW_ERR_UNSUPPORTED, "vsw_patch_im2col: row tile of %zu bytes too large", smem);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * Dout * Hp;
    VSW_REQUIRE(rows < (1LL << 31), VSW_ERR_UNSUPPORTED, "vsw_patch_im2col: too many rows");
    VSW_PE_DISPATCH(x_dtype, dtype, {
        auto k = im2col_kernel<TX, T>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(unsigned)rows, 256, smem, st>>>((const TX*)x, (T*)col, B, Cin, D, H, W, pd, ph, pw, Dout, Hp, Wp);
    });
    return check_launch("vsw_patch_im2col");
}

extern "C" int vsw_patch_col2im(const void* dcol, void* dx, int B, int Cin, int D, int H, int W, int pd, int ph,
                                int pw, int x_dtype, int dtype, void* stream) {
    VSW_REQUIRE(dcol && dx && B > 0 && Cin > 0 && D > 0 && H > 0 && W > 0 && pd > 0 && ph > 0 && pw > 0, VSW_ERR_ARG,
                "vsw_patch_col2im: bad args");
    const int Dout = D + 2 - pd, Hp = (H + ph - 1) / ph, Wp = (W + pw - 1) / pw;
    VSW_REQUIRE(Dout > 0, VSW_ERR_ARG, "vsw_patch_col2im: D too short");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)B * Cin * D * H * W;
    int blocks = (int)((total + 255) / 256 < (long long)kNumSMs * 32 ? (total + 255) / 256 : (long long)kNumSMs * 32);
    VSW_PE_DISPATCH(x_dtype, dtype,
                    (col2im_kernel<TX, T><<<blocks, 256, 0, st>>>((const T*)dcol, (TX*)dx, B, Cin, D, H, W, pd, ph, pw,
                                                                  Dout, Hp, Wp)));
    return check_launch("vsw_patch_col2im");
}
