// CUDA-core GEMM (fp32 accumulate, storage dtype templated).  See gemm_simt.cuh.
#include "gemm_simt.cuh"

namespace vsw {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename T, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) simt_gemm_kernel(SimtGemmParams p) {
    __shared__ float As[BK][BM + PAD];
    __shared__ float Bs[BK][BN + PAD];
    const T* __restrict__ A = (const T*)p.A;
    const T* __restrict__ Bm = (const T*)p.B;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int split = blockIdx.z;
    const int kbeg = split * p.k_per_split;
    const int kend = min(p.K, kbeg + p.k_per_split);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- A tile (BM x BK) ----
#pragma unroll
        for (int it = 0; it < (BM * BK) / 256; ++it) {
            const int idx = tid + it * 256;
            const int mm = A_KC ? idx / BK : idx % BM;
            const int kk = A_KC ? idx % BK : idx / BM;
            const int m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < p.M && k < kend) {
                long long arow = m;
                float sc = 1.f;
                bool ok = true;
                if (p.a_rowmap || p.a_rowscale) {
                    const int b = m / p.rows_per_batch;
                    const int r = m - b * p.rows_per_batch;
                    if (p.a_rowmap) {
                        const int s = p.a_rowmap[r];
                        ok = s >= 0;
                        arow = (long long)b * p.src_rows_per_batch + (ok ? s : 0);
                    }
                    if (p.a_rowscale) sc = p.a_rowscale[b];
                }
                if (ok) v = sc * to_f<T>(A[arow * p.sam + (long long)k * p.sak]);
                if (p.a_out && blockIdx.x == 0) ((T*)p.a_out)[(long long)m * p.K + k] = from_f<T>(v);
            }
            As[kk][mm] = v;
        }
        // ---- B tile (BN x BK) ----
#pragma unroll
        for (int it = 0; it < (BN * BK) / 256; ++it) {
            const int idx = tid + it * 256;
            const int nn = B_KC ? idx / BK : idx % BN;
            const int kk = B_KC ? idx % BK : idx / BN;
            const int n = n0 + nn, k = k0 + kk;
            float v = 0.f;
            if (n < p.N && k < kend) v = to_f<T>(Bm[(long long)n * p.sbn + (long long)k * p.sbk]);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        long long drow = m;
        float rsc = 1.f;
        bool row_ok = true;
        if (p.epi == SE_RESIDUAL) {
            const int b = m / p.rows_per_batch;
            const int r = m - b * p.rows_per_batch;
            int d = r;
            if (p.rowmap) { d = p.rowmap[r]; row_ok = d >= 0; }
            drow = (long long)b * p.dst_rows_per_batch + d;
            if (p.rowscale) rsc = p.rowscale[b];
        }
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j];
            if (p.epi == SE_PARTIAL) {
                p.partial[((long long)split * p.M + m) * p.N + n] = v;
                continue;
            }
            if (p.bias) v += to_f<T>(((const T*)p.bias)[n]);
            const long long o = drow * p.ldc + n;
            if (p.epi == SE_GELU) {
                if (p.aux_out) ((T*)p.aux_out)[o] = from_f<T>(p.aux_is_grad ? gelu_grad_f(v) : v);
                v = gelu_f(v);
            } else if (p.epi == SE_RESIDUAL) {
                v = to_f<T>(((const T*)p.res)[o]) + rsc * v;
            } else if (p.epi == SE_DGRAD) {
                if (p.gelu_pre) {
                    const float u = to_f<T>(((const T*)p.gelu_pre)[o]);
                    v *= p.aux_is_grad ? u : gelu_grad_f(u);
                }
            }
            ((T*)p.C)[o] = from_f<T>(v);
        }
    }
}

int launch_simt_gemm(const SimtGemmParams& p, bool a_kc, bool b_kc, int dtype, cudaStream_t st) {
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), p.ksplit > 0 ? p.ksplit : 1);
    VSW_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, VSW_ERR_UNSUPPORTED, "simt gemm: grid too large (M=%d)", p.M);
    SimtGemmParams q = p;
    if (q.ksplit <= 0) { q.ksplit = 1; q.k_per_split = q.K; }
#define VSW_LAUNCH_SIMT(AKC, BKC) \
    VSW_DISPATCH_DTYPE(dtype, T, (simt_gemm_kernel<T, AKC, BKC><<<grid, 256, 0, st>>>(q)))
    if (a_kc && b_kc) { VSW_LAUNCH_SIMT(true, true); }
    else if (a_kc && !b_kc) { VSW_LAUNCH_SIMT(true, false); }
    else if (!a_kc && !b_kc) { VSW_LAUNCH_SIMT(false, false); }
    else { VSW_LAUNCH_SIMT(false, true); }
#undef VSW_LAUNCH_SIMT
    return check_launch("simt_gemm");
}

template <typename TO>
__global__ void partial_reduce_kernel(const float* __restrict__ partial, int splits, long long elems,
                                      TO* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < elems;
         i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += partial[(long long)k * elems + i];
        out[i] = from_f<TO>(s);
    }
}

int launch_partial_reduce(const float* partial, int splits, long long elems, void* out, int out_dtype,
                          cudaStream_t st) {
    int blocks = (int)((elems + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    VSW_DISPATCH_DTYPE(out_dtype, TO,
                       (partial_reduce_kernel<TO><<<blocks, 256, 0, st>>>(partial, splits, elems, (TO*)out)));
    return check_launch("partial_reduce");
}

// ---- column sum (bias gradient): HBM-bound, 16-byte loads, fixed-order two-pass reduction ----
constexpr int kColsumSplits = 296;   // 2 row-splits per SM
size_t colsum_ws_bytes(int N) { return (size_t)kColsumSplits * N * sizeof(float); }

// blockDim (TX, TY): thread (tx, ty) owns the 16-byte column vector tx of the block's column tile and every
// (TY * gridDim.y)-th row; partials are combined over ty through shared memory in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dy, int M, int N, float* __restrict__ part) {
    constexpr int VN = Vec16<T>::N;
    extern __shared__ float sm[];   // [TY][TX][VN]
    const int nvec = N / VN;
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    float a[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) a[e] = 0.f;
    if (vi < nvec) {
        const long long stride = (long long)gridDim.y * blockDim.y;
        long long m = (long long)blockIdx.y * blockDim.y + threadIdx.y;
        // 4 independent loads in flight per thread
        for (; m + 3 * stride < M; m += 4 * stride) {
            float v0[VN], v1[VN], v2[VN], v3[VN];
            load_vec<T>(dy + m * N + (long long)vi * VN, v0);
            load_vec<T>(dy + (m + stride) * N + (long long)vi * VN, v1);
            load_vec<T>(dy + (m + 2 * stride) * N + (long long)vi * VN, v2);
            load_vec<T>(dy + (m + 3 * stride) * N + (long long)vi * VN, v3);
#pragma unroll
            for (int e = 0; e < VN; ++e) a[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]);
        }
        for (; m < M; m += stride) {
            float v0[VN];
            load_vec<T>(dy + m * N + (long long)vi * VN, v0);
#pragma unroll
            for (int e = 0; e < VN; ++e) a[e] += v0[e];
        }
    }
    float* mine = sm + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * VN;
#pragma unroll
    for (int e = 0; e < VN; ++e) mine[e] = a[e];
    __syncthreads();
    if (threadIdx.y == 0 && vi < nvec) {
        for (int ty = 1; ty < blockDim.y; ++ty) {
            const float* o = sm + ((size_t)ty * blockDim.x + threadIdx.x) * VN;
#pragma unroll
            for (int e = 0; e < VN; ++e) a[e] += o[e];
        }
        float* p = part + (size_t)blockIdx.y * N + (size_t)vi * VN;
#pragma unroll
        for (int e = 0; e < VN; ++e) p[e] = a[e];
    }
}

// scalar fallback for N not a multiple of the vector width
template <typename T>
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const T* __restrict__ dy, int M, int N, float* __restrict__ part) {
    __shared__ float sm[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (n < N)
        for (long long m = (long long)blockIdx.y * 8 + threadIdx.y; m < M; m += (long long)gridDim.y * 8)
            s += to_f<T>(dy[m * N + n]);
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) t += sm[y][threadIdx.x];
        part[(long long)blockIdx.y * N + n] = t;
    }
}

int launch_colsum(const void* dy, int M, int N, void* db, int dtype, int out_dtype, void* ws, cudaStream_t st) {
    const int vn = dtype == VSW_F32 ? 4 : 8;
    int rc;
    int splits;
    if (N % vn == 0) {
        const int nvec = N / vn;
        const int TX = nvec >= 32 ? 32 : 16, TY = 256 / TX;
        const int xtiles = ceil_div(nvec, TX);
        splits = ceil_div(M, TY * 8);
        const int want = (2 * kNumSMs + xtiles - 1) / xtiles;
        if (splits > want) splits = want;
        if (splits > kColsumSplits) splits = kColsumSplits;
        if (splits < 1) splits = 1;
        dim3 grid(xtiles, splits), block(TX, TY);
        const size_t smem = (size_t)256 * vn * sizeof(float);
        VSW_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, block, smem, st>>>((const T*)dy, M, N, (float*)ws)));
    } else {
        splits = ceil_div(M, 8 * 16);
        if (splits > kColsumSplits) splits = kColsumSplits;
        if (splits < 1) splits = 1;
        dim3 grid(ceil_div(N, 32), splits), block(32, 8);
        VSW_DISPATCH_DTYPE(dtype, T, (colsum_scalar_kernel<T><<<grid, block, 0, st>>>((const T*)dy, M, N, (float*)ws)));
    }
    rc = check_launch("colsum");
    if (rc) return rc;
    return launch_partial_reduce((const float*)ws, splits, N, db, out_dtype, st);
}

}  // namespace vsw
