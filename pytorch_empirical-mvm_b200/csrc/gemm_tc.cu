// tcgen05 / TMEM / TMA bf16 GEMM family for sm_100a.
//
// One persistent, warp-specialised kernel template covers the three GEMM forms of a Linear layer:
//   forward  y  = x  w^T        A = x  (K-major),  B = w  (K-major)
//   dgrad    dx = dy w          A = dy (K-major),  B = w  (MN-major: w is (N,K) row-major, reduced over N)
//   wgrad    dw = dy^T x        A = dy (MN-major), B = x  (MN-major), reduced over the M rows, split across
//                               CTAs into fp32 partials + fixed-order second pass (deterministic)
// Roles (18 warps, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
// warps 2..17 = epilogue (TMEM -> registers -> fused epilogue -> smem transposition -> coalesced global stores; in the wgrad
// variant four of them add up the dy tiles for the bias gradient instead).  The accumulator is double-buffered
// in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the MMAs of tile i+1; operands flow
// through a STAGES-deep TMA ring with 128-byte swizzle (no bank conflicts, no padding).
// Fused epilogues (compile-time): bias | bias+GELU(+pre-activation | +GELU') | bias+drop-path scale+row scatter+residual |
//                  dgrad (* GELU'(pre) | * saved GELU') | fp32 split partial.
// CTA-pair mode (cta_group::2, 256 x 256 tiles, each CTA loads half of B) for the fwd / dgrad GEMMs with 256-wide tiles.
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <mutex>
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace vsw {

// ---------------------------------------------------------------------------------------------
// host: tensor map encoder
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    });
    return fn;
}

bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                       uint32_t box_rows, uint32_t box_cols) {
    auto fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows,
                  box_cols);
        return false;
    }
    return true;
}

static CUtensorMapL2promotion promo_of(int bytes) {
    return bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

bool make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows, uint32_t box_cols,
                       int swizzle_bytes, int l2_promotion_bytes) {
    auto fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t gdim[3] = {cols, rows, batch};
    cuuint64_t gstride[2] = {row_stride_elems * 2, batch_stride_elems * 2};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    promo_of(l2_promotion_bytes), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d) batch=%llu rows=%llu cols=%llu", (int)r,
                  (unsigned long long)batch, (unsigned long long)rows, (unsigned long long)cols);
        return false;
    }
    return true;
}

bool make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                       const uint32_t* box, int swizzle_bytes, int l2_promotion_bytes) {
    auto fn = get_encode_fn();
    if (!fn || rank < 1 || rank > 5) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 1; i < rank; ++i) gstride[i - 1] = strides_elems[i] * 2;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstride, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    promo_of(l2_promotion_bytes), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r); return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
namespace {

// phase counters (printed under VSW_GEMM_DEBUG=1) exist only in builds with -DVSW_GEMM_PROF=1 (VSW_NVCC_EXTRA of build.py)
#ifndef VSW_GEMM_PROF
#define VSW_GEMM_PROF 0
#endif
#if VSW_GEMM_PROF
#define GCLK() clock64()
#else
#define GCLK() 0LL
#endif

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int NUM_EPI_WARPS = 16;   // 4 per TMEM lane quadrant: the fused epilogues are latency/MUFU bound, not issue bound
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

// TE_RESIDUAL_ID: the residual epilogue whose output rows ARE the tile's rows (fc2: no window-reverse row map), so that both the
// residual tile and the output can go through the tensor maps
enum TcEpi { TE_BIAS = 0, TE_GELU = 1, TE_RESIDUAL = 2, TE_DGRAD = 3, TE_PARTIAL = 4, TE_DGRAD_GELU = 5, TE_GELU_GRAD = 6, TE_DGRAD_MUL = 7,
             TE_RESIDUAL_ID = 8 };

struct TcParams {
    int M, N, K;                 // output M x N, reduction length K
    int n_tiles_m, n_tiles_n, splits, k_per_split;
    int epi;
    const __nv_bfloat16* bias; __nv_bfloat16* out; __nv_bfloat16* aux_out; const __nv_bfloat16* res;
    const int32_t* rowmap; const float* rowscale; int rows_per_batch, dst_rows_per_batch;
    const __nv_bfloat16* gelu_pre; float* partial;
    float* colsum;               // COLSUM kernels: per-split row sums of the A operand, [splits][M]
    long long ldc;
    long long* dbg;   // optional phase counters (VSW_GEMM_DEBUG)
};

// GELU(x) = x Phi(x) with Phi(x) ~= 0.5 + 0.5 tanh(x (a + b x^2 + c x^4)): a three-term fit to the exact (erf) normal CDF
// (max |GELU error| 3.7e-5, max |GELU' error| 9.3e-5 over all x; MUFU.TANH adds <= 2.4e-4 |x|) -- an order of magnitude
// below bf16 output resolution.  One MUFU op and ~7 FMA-pipe instructions per element instead of erff's ~25 / 3 MUFU,
// which made the GELU epilogues issue-bound.  (The fp32 SIMT path and the oracle use erff.)
constexpr float GELU_A = 7.97422827e-01f, GELU_B = 3.70038147e-02f, GELU_C = -3.47516169e-04f;
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
    const float s = fminf(x * x, 64.0f);   // beyond |x| = 8 tanh is saturated; the clamp keeps the quartic monotone
    const float t = tanh_approx(x * fmaf(fmaf(GELU_C, s, GELU_B), s, GELU_A));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
    const float s = fminf(x * x, 64.0f);
    const float t = tanh_approx(x * fmaf(fmaf(GELU_C, s, GELU_B), s, GELU_A));
    const float du = fmaf(fmaf(5.0f * GELU_C, s, 3.0f * GELU_B), s, GELU_A);   // d/dx [x P(x^2)]
    const float w = fmaf(-t, t, 1.0f) * (0.5f * x);
    return fmaf(w, du, fmaf(0.5f, t, 0.5f));
}

// y = GELU(x) and g = GELU'(x) from one tanh (training-form fc1 epilogue)
__device__ __forceinline__ void gelu_both_fast(float x, float& y, float& g) {
    const float s = fminf(x * x, 64.0f);
    const float t = tanh_approx(x * fmaf(fmaf(GELU_C, s, GELU_B), s, GELU_A));
    const float du = fmaf(fmaf(5.0f * GELU_C, s, 3.0f * GELU_B), s, GELU_A);
    const float phi = fmaf(0.5f, t, 0.5f);
    const float w = fmaf(-t, t, 1.0f) * (0.5f * x);
    y = x * phi;
    g = fmaf(w, du, phi);
}

// The kernels move 16-bit elements; F16 selects IEEE half instead of bfloat16 for the conversions and the MMA operand format
// (pointers stay typed __nv_bfloat16* = "a 16-bit element").
template <bool F16> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (F16) { const __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<const uint32_t*>(&v); }
    else return tc::pack_bf16(lo, hi);
}
template <bool F16> __device__ __forceinline__ float2 unpack2(uint32_t u) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    else return tc::unpack_bf16(u);
}
template <bool F16> __device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]) {
    float2 a = unpack2<F16>(u.x), b = unpack2<F16>(u.y), c = unpack2<F16>(u.z), d = unpack2<F16>(u.w);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}
template <bool F16> __device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&o)[8]) {
    unpack8<F16>(__ldg(reinterpret_cast<const uint4*>(p)), o);
}
template <bool F16> __device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack2<F16>(v[0], v[1]); u.y = pack2<F16>(v[2], v[3]);
    u.z = pack2<F16>(v[4], v[5]); u.w = pack2<F16>(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// EPI is a COMPILE-TIME parameter: with a run-time switch the compiler if-converts the branches and every tile pays
// for the GELU / GELU' math (MUFU-bound) whatever the epilogue is.
//
// COLSUM (wgrad only, A MN-major): four of the sixteen epilogue warps become "column-sum" warps that add up the A
// tiles (= dy^T) of the first N-tile's work items straight from the TMA-filled shared-memory stages, so the bias
// gradient sum_m dy[m, n] costs no second pass over dy.  They hold each stage (extra arrivals on `empty`) until read.
//
// PAIR: two CTAs of a cluster (one TPC) work on one 256 x BN tile with cta_group::2 MMAs of M = 256: each CTA loads its own
// 128 A rows and only HALF of the B tile (the tensor core reads the other half from the peer's shared memory), which cuts
// the bytes every SM pulls through its L2 port per k-block from 48 KB to 32 KB -- the single-CTA 128 x 256 main loop is
// bound by that port (~64 B/clk/SM), not by the tensor pipe.  Each CTA keeps its own 128 accumulator lanes, so the
// epilogue is unchanged.  Barriers: `full` lives in the leader (both producers arrive + their bytes), `empty` / `tfull`
// are multicast commits to both CTAs, `tempty` of the leader collects the epilogue warps of both CTAs.
template <int BN, bool A_MN, bool B_MN, int STAGES, int EPI, bool COLSUM = false, bool PAIR = false, bool F16 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, const TcParams p) {
    static_assert(!COLSUM || (A_MN && EPI == TE_PARTIAL), "COLSUM is a wgrad-only variant");
    static_assert(!(PAIR && COLSUM), "the column-sum warps need a CTA-local `full` barrier");
    // outputs whose rows are the tile's rows leave through tensor-map stores from the per-warp staging buffer (one instruction per
    // 32 x 32 chunk instead of 4 LDS + 12 SHFL + 4 predicated 16-byte stores per lane: fc1 + GELU + GELU' at M = 802816, N = 512,
    // K = 128 went from 0.49 to 0.38 ms, the plain epilogue from 0.25 to 0.19 = cuBLAS); the row-scattering residual epilogue and
    // the fp32 split-K partials keep their own paths
    // (epilogues with a side tensor use the staging buffer twice per chunk -- the wait for the bulk store's read in between made
    // them 8 % slower, measured -- and keep the per-lane stores too)
    // SIDE2 (dgrad x GELU' forms, residual with identity rows): the side tensor tile of a chunk arrives by a tensor-map LOAD into a second per-warp buffer (one
    // TMA stage fewer), issued one chunk ahead -- instead of 4 x (2 SHFL + LDG + STS) per lane on the chunk's critical path --
    // and the output leaves by a tensor-map store like the side-less epilogues
    constexpr bool SIDE2 = (EPI == TE_DGRAD_MUL || EPI == TE_DGRAD_GELU || EPI == TE_RESIDUAL_ID);
    constexpr bool TMA_OUT = (EPI == TE_BIAS || EPI == TE_GELU || EPI == TE_GELU_GRAD || EPI == TE_DGRAD || SIDE2);
    constexpr int CS_WARPS = COLSUM ? 4 : 0;
    constexpr int EPI_WARPS = NUM_EPI_WARPS - CS_WARPS;
    constexpr int BNL = PAIR ? BN / 2 : BN;           // B rows this CTA loads
    constexpr int B_BYTES = BNL * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // power of two for BN in {64,128,256}
    constexpr uint32_t IDESC = tc::idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0) & (F16 ? ~((1u << 7) | (1u << 10)) : ~0u);   // A / B format field: 1 = bf16, 0 = f16
    const uint32_t cta_rank = PAIR ? tc::cluster_ctarank() : 0u;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int TILE_M = PAIR ? 2 * BM : BM;        // rows of a work item

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    constexpr uint32_t STG_BYTES = 32 * 64;   // per-epilogue-warp transposition buffer (32 rows x 64 B, XOR-swizzled 16-B units)
    const uint32_t stage_base = tc::smem_u32(smem + STAGES * STAGE_BYTES + 256);
    uint64_t* side_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + 256 + 2 * NUM_EPI_WARPS * STG_BYTES);   // SIDE2: one per epilogue warp

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.n_tiles_m * p.n_tiles_n;
    const int total = tiles * p.splits;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], PAIR ? 2 : 1); tc::mbar_init(&empty[s], 1 + CS_WARPS); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], (PAIR ? 2 : 1) * EPI_WARPS); }
        if constexpr (SIDE2) { for (int w8 = 0; w8 < NUM_EPI_WARPS; ++w8) tc::mbar_init(&side_bar[w8], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tc::tmem_alloc_2cta(tmem_slot, TMEM_COLS); else tc::tmem_alloc(tmem_slot, TMEM_COLS); }
    tc::tc_fence_before();
    if constexpr (PAIR) tc::cluster_sync_all(); else __syncthreads();   // PAIR: the peer's barriers are initialised too
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (tc::elect_one()) {   // elect.sync, not lane == 0: the TMA / MMA operands stay warp-uniform for ptxas
            int stage = 0; uint32_t phase = 0;
            for (int w = unit; w < total; w += n_units) {
                const int split = w / tiles, t = w - split * tiles;
                const int m0 = (t / p.n_tiles_n) * TILE_M + (int)cta_rank * BM;       // this CTA's A rows
                const int n0 = (t % p.n_tiles_n) * BN + (int)cta_rank * (BN - BNL);   // this CTA's share of the B rows
                const int kbeg = split * p.k_per_split;
                const int kend = min(p.K, kbeg + p.k_per_split);
                for (int k0 = kbeg; k0 < kend; k0 += BK) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * STAGE_BYTES;
                    uint8_t* sB = sA + A_BYTES;
                    if constexpr (PAIR) {
                        // both CTAs load into their own smem; all completion bytes are signalled on the LEADER's barrier
                        // (count 2: the leader's arrive + expect of both CTAs' bytes, the peer's plain remote arrive)
                        const uint32_t fb = tc::mapa_u32(tc::smem_u32(&full[stage]), 0);
                        if (cta_rank == 0) tc::mbar_expect_tx(&full[stage], 2 * STAGE_BYTES); else tc::mbar_arrive_cluster(fb);
                        if (!A_MN) {
                            tc::tma_load_2d_2sm(&tmA, fb, sA, k0, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j) tc::tma_load_2d_2sm(&tmA, fb, sA + j * 8192, m0 + 64 * j, k0);
                        }
                        if (!B_MN) {
                            tc::tma_load_2d_2sm(&tmB, fb, sB, k0, n0);                     // box 64(k) x BN/2(n)
                        } else {
#pragma unroll
                            for (int j = 0; j < BNL / 64; ++j) tc::tma_load_2d_2sm(&tmB, fb, sB + j * 8192, n0 + 64 * j, k0);
                        }
                    } else {
                    tc::mbar_expect_tx(&full[stage], STAGE_BYTES);
                    if (!A_MN) {
                        tc::tma_load_2d(&tmA, &full[stage], sA, k0, m0);          // box 64(k) x 128(m)
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)                          // box 64(m) x 64(k)
                            tc::tma_load_2d(&tmA, &full[stage], sA + j * 8192, m0 + 64 * j, k0);
                    }
                    if (!B_MN) {
                        tc::tma_load_2d(&tmB, &full[stage], sB, k0, n0);          // box 64(k) x BN(n)
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)                          // box 64(n) x 64(k)
                            tc::tma_load_2d(&tmB, &full[stage], sB + j * 8192, n0 + 64 * j, k0);
                    }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (cta_rank == 0 && tc::elect_one()) {   // elect.sync: ptxas then keeps the MMA operands warp-uniform (no R2UR.BROADCAST loop per MMA);   // PAIR: the leader issues for both CTAs
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = unit; w < total; w += n_units) {
                const int split = w / tiles;
                const int kbeg = split * p.k_per_split;
                const int kend = min(p.K, kbeg + p.k_per_split);
                long long m_a = GCLK();
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                long long m_b = GCLK(), m_wait_full = 0;
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int k0 = kbeg; k0 < kend; k0 += BK) {
                    long long f_a = GCLK();
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    m_wait_full += GCLK() - f_a;
                    const uint32_t aaddr = tc::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t baddr = aaddr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = A_MN ? tc::smem_desc_sw128(aaddr + k * 2048, 8192, 1024)
                                                 : tc::smem_desc_sw128(aaddr + k * 32, 0, 1024);
                        const uint64_t bd = B_MN ? tc::smem_desc_sw128(baddr + k * 2048, 8192, 1024)
                                                 : tc::smem_desc_sw128(baddr + k * 32, 0, 1024);
                        if constexpr (PAIR) tc::umma_bf16_2cta(d_tmem, ad, bd, IDESC, accumulate);
                        else tc::umma_bf16(d_tmem, ad, bd, IDESC, accumulate);
                        accumulate = 1;
                    }
                    // smem slot reusable once these MMAs retire (PAIR: in both CTAs)
                    if constexpr (PAIR) tc::umma_commit_2cta(&empty[stage]); else tc::umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue
                if constexpr (PAIR) tc::umma_commit_2cta(&tfull[acc]); else tc::umma_commit(&tfull[acc]);
                if (VSW_GEMM_PROF && p.dbg && blockIdx.x == 0) { p.dbg[0] += m_b - m_a; p.dbg[1] += m_wait_full; p.dbg[2] += GCLK() - m_b; p.dbg[3] += 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (COLSUM && warp >= 2 + EPI_WARPS) {
        // ================= column-sum warps (COLSUM variant) =================
        // A stage holds BM/64 boxes of 64 reduction rows x 128 B (64 bf16 of the A-row index, 128B-swizzled).  Thread
        // (box bj, physical 16-byte chunk pch, row class rcl) reads rows rcl + 8 i: they all hold the same logical chunk
        // pch ^ rcl, so the thread accumulates 8 fixed A-rows; 8 threads per (box, chunk) are combined at the end.
        const int te = threadIdx.x - (2 + EPI_WARPS) * 32;   // 0..127
        const int bj = te >> 6, pch = te & 7, rcl = (te >> 3) & 7;
        float* scr = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // [128][8] (epilogue staging is unused by TE_PARTIAL)
        int stage = 0; uint32_t phase = 0;
        for (int w = unit; w < total; w += n_units) {
            const int split = w / tiles, t = w - split * tiles;
            const int m0 = (t / p.n_tiles_n) * BM;
            const bool mine = (t % p.n_tiles_n) == 0;
            const int kbeg = split * p.k_per_split;
            const int kend = min(p.K, kbeg + p.k_per_split);
            float cs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] = 0.f;
            for (int k0 = kbeg; k0 < kend; k0 += BK) {
                tc::mbar_wait(&full[stage], phase);
                if (mine) {
                    const uint32_t sA = tc::smem_u32(smem + stage * STAGE_BYTES) + bj * 8192 + pch * 16;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 v = tc::lds_u4(sA + (rcl + 8 * i) * 128);
                        const float2 a = unpack2<F16>(v.x), b = unpack2<F16>(v.y), c = unpack2<F16>(v.z), d = unpack2<F16>(v.w);
                        cs[0] += a.x; cs[1] += a.y; cs[2] += b.x; cs[3] += b.y; cs[4] += c.x; cs[5] += c.y; cs[6] += d.x; cs[7] += d.y;
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (mine) {
#pragma unroll
                for (int e = 0; e < 8; ++e) scr[te * 8 + e] = cs[e];
            }
            tc::named_bar_sync(1, CS_WARPS * 32);
            if (mine) {
                // A-row n of the tile: box n / 64, logical chunk (n % 64) / 8, element n % 8; fixed summation order
                const int nb = te >> 6, q8 = (te & 63) >> 3, e = te & 7;
                float tsum = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) tsum += scr[(nb * 64 + c * 8 + (q8 ^ c)) * 8 + e];
                if (m0 + te < p.M) p.colsum[(long long)split * p.M + m0 + te] = tsum;
            }
            tc::named_bar_sync(1, CS_WARPS * 32);
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;       // which interleaved 32-column chunks it owns (0..EPI_WARPS/4-1)
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t side_phase = 0;                 // SIDE2: parity of this warp's side-tile barrier
        const uint32_t stg_w = stage_base + (uint32_t)(warp - 2) * STG_BYTES;
        const uint32_t sstg_w = stg_w + NUM_EPI_WARPS * STG_BYTES;      // SIDE2: the side tile's own buffer
        uint64_t* sbar = &side_bar[warp - 2];
        // SIDE2: tensor-map load of the 32 x 32 side tile of chunk c of tile (m0, n0) (rows >= M / columns >= N arrive as zeros)
        auto issue_side = [&](int m0, int n0, int c) {
            if constexpr (SIDE2) {
                if (lane == 0) {
                    tc::mbar_expect_tx(sbar, STG_BYTES);
                    tc::tma_load_2d(&tmAux, sbar, smem + (sstg_w - tc::smem_u32(smem)), n0 + c * 32, m0 + q * 32);
                }
            }
        };
        for (int w = unit; w < total; w += n_units) {
            const int split = w / tiles, t = w - split * tiles;
            const int m0 = (t / p.n_tiles_n) * TILE_M + (int)cta_rank * BM, n0 = (t % p.n_tiles_n) * BN;
            long long e_a = GCLK(), dbg_ld = 0, dbg_math = 0, dbg_st = 0;
            issue_side(m0, n0, half);            // in flight while the accumulator is still being computed
            // the bias of this warp's first chunk is fetched BEFORE the wait for the accumulator (an L2 round trip of ~400
            // cycles otherwise sits in front of every chunk's arithmetic); later chunks fetch theirs behind the TMEM load
            uint4 bq[4];
            auto fetch_bias = [&](int c) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = n0 + c * 32 + g * 8;
                    bq[g] = (p.bias && col < p.N) ? __ldg(reinterpret_cast<const uint4*>(p.bias + col)) : make_uint4(0, 0, 0, 0);
                }
            };
            if (EPI != TE_PARTIAL) fetch_bias(half);
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tc_fence_after();
            long long e_b = GCLK();
            const int row = m0 + q * 32 + lane;
            const bool row_in = row < p.M;
            long long drow = row;
            float rsc = 1.f;
            bool row_ok = row_in;
            if (EPI == TE_RESIDUAL_ID && row_in) {
                if (p.rowscale) rsc = __ldg(p.rowscale + row / p.rows_per_batch);
            }
            if (EPI == TE_RESIDUAL && row_in) {
                const int b = row / p.rows_per_batch;
                const int r = row - b * p.rows_per_batch;
                int d = r;
                if (p.rowmap) { d = __ldg(p.rowmap + r); row_ok = d >= 0; }
                drow = (long long)b * p.dst_rows_per_batch + d;
                if (p.rowscale) rsc = __ldg(p.rowscale + b);
            }
            // Per 32-column chunk: TMEM -> registers (thread = row) -> fused epilogue math -> bf16 -> per-warp smem
            // transposition buffer -> global stores in which 4 consecutive lanes write one 64-byte row segment (full
            // sectors), instead of 32 lanes writing 16 bytes to 32 different rows.
            const uint32_t stg = stage_base + (uint32_t)(warp - 2) * STG_BYTES;
#pragma unroll 1
            for (int c = half; c < BN / 32; c += EPI_WARPS / 4) {
                uint32_t r32[32];
                if constexpr (TMA_OUT) { if (lane == 0) tc::bulk_wait_read_all(); }   // the previous chunk's store has read the staging buffer
                __syncwarp();
                long long c_a = GCLK();
                tc::tmem_ld_32x32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), r32);
                const int col0 = n0 + c * 32;
                if (EPI != TE_PARTIAL && c != half) fetch_bias(c);
                // Side tensor of the epilogue (residual / GELU' / pre-activation), same (row, column) footprint as the output:
                // loaded COALESCED (4 lanes per 64-byte row segment, rows via the owning lanes' registers) into the warp's
                // staging buffer while the TMEM load is in flight, then re-read in the thread-per-row layout.
                constexpr bool HAS_SIDE = (EPI == TE_RESIDUAL || EPI == TE_RESIDUAL_ID || EPI == TE_DGRAD_MUL || EPI == TE_DGRAD_GELU);
                uint4 side[HAS_SIDE ? 4 : 1];
                if constexpr (SIDE2) {
                    tc::mbar_wait(sbar, side_phase);
                    side_phase ^= 1;
                    const uint32_t rowa = sstg_w + lane * 64;    // the TMA unit's 64-byte swizzle is keyed by the absolute address
#pragma unroll
                    for (int g = 0; g < 4; ++g) side[g] = tc::lds_u4(rowa + ((g ^ ((rowa >> 7) & 3)) << 4));
                    tc::fence_proxy_async();                     // our reads of the buffer come before the next tile load into it
                    __syncwarp();
                    if (c + EPI_WARPS / 4 < BN / 32) issue_side(m0, n0, c + EPI_WARPS / 4);
                } else if constexpr (HAS_SIDE) {
                    const __nv_bfloat16* sp = (EPI == TE_RESIDUAL) ? p.res : p.gelu_pre;
                    const int unit = lane & 3;
                    const int colu = col0 + unit * 8;
                    uint4 ld[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rl = k * 8 + (lane >> 2);
                        const long long dr = __shfl_sync(0xffffffffu, drow, rl);
                        const int okr = __shfl_sync(0xffffffffu, (int)row_ok, rl);
                        ld[k] = (okr && colu < p.N) ? __ldg(reinterpret_cast<const uint4*>(sp + dr * p.ldc + colu)) : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rl = k * 8 + (lane >> 2);
                        tc::sts_u4(stg + rl * 64 + ((unit ^ ((rl >> 1) & 3)) << 4), ld[k]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < 4; ++g) side[g] = tc::lds_u4(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4));
                }
                tc::tmem_ld_wait();
                long long c_b = GCLK();
                if constexpr (EPI == TE_PARTIAL) {
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int col = col0 + g * 8;
                            if (col < p.N) {
                                float* dst = p.partial + ((long long)split * p.M + row) * p.N + col;
                                *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(r32[g * 8]), __uint_as_float(r32[g * 8 + 1]),
                                                                              __uint_as_float(r32[g * 8 + 2]), __uint_as_float(r32[g * 8 + 3]));
                                *reinterpret_cast<float4*>(dst + 4) = make_float4(__uint_as_float(r32[g * 8 + 4]), __uint_as_float(r32[g * 8 + 5]),
                                                                                  __uint_as_float(r32[g * 8 + 6]), __uint_as_float(r32[g * 8 + 7]));
                            }
                        }
                    }
                } else {
                // ---- epilogue math in the thread-per-row layout; results packed to bf16 (16 words per thread)
                uint32_t packed[16], packed_aux[16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int col = col0 + g * 8;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r32[g * 8 + e]);
                    if (col < p.N) {
                        if (p.bias) {
                            float bb[8];
                            unpack8<F16>(bq[g], bb);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] += bb[e];
                        }
                        if constexpr (EPI == TE_GELU) {
                            if (p.aux_out) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) packed_aux[g * 4 + e] = pack2<F16>(v[2 * e], v[2 * e + 1]);
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = gelu_fast(v[e]);
                        } else if constexpr (EPI == TE_GELU_GRAD) {
                            float gd[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) gelu_both_fast(v[e], v[e], gd[e]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) packed_aux[g * 4 + e] = pack2<F16>(gd[2 * e], gd[2 * e + 1]);
                        } else if constexpr (HAS_SIDE) {
                            float u[8];
                            unpack8<F16>(side[g], u);
                            if constexpr (EPI == TE_DGRAD_MUL) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] *= u[e];
                            } else if constexpr (EPI == TE_RESIDUAL || EPI == TE_RESIDUAL_ID) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = fmaf(rsc, v[e], u[e]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] *= gelu_grad_fast(u[e]);
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) packed[g * 4 + e] = pack2<F16>(v[2 * e], v[2 * e + 1]);
                }
                long long c_c = GCLK();
                // ---- transposed, coalesced stores (one or two outputs)
                const int npass = ((EPI == TE_GELU && p.aux_out) || EPI == TE_GELU_GRAD) ? 2 : 1;
                for (int pass = 0; pass < npass; ++pass) {
                    const uint32_t* src = (npass == 2 && pass == 0) ? packed_aux : packed;
                    __nv_bfloat16* outp = (npass == 2 && pass == 0) ? p.aux_out : p.out;
                    if constexpr (TMA_OUT) {
                        if (pass > 0) { if (lane == 0) tc::bulk_wait_read_all(); }
                        __syncwarp();
                        const uint32_t rowa = stg + lane * 64;          // 64-byte swizzle keyed by the ABSOLUTE address, as the TMA unit applies it
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            tc::sts_u4(rowa + ((g ^ ((rowa >> 7) & 3)) << 4), make_uint4(src[g * 4], src[g * 4 + 1], src[g * 4 + 2], src[g * 4 + 3]));
                        tc::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tc::tma_store_2d((npass == 2 && pass == 0) ? &tmAux : &tmOut, stg, col0, m0 + q * 32);   // rows >= M, columns >= N are clipped
                            tc::bulk_commit_group();
                        }
                    } else {
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < 4; ++g)   // row `lane`, 16-byte unit g XOR-swizzled by the row pair: conflict-free both ways
                        tc::sts_u4(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4), make_uint4(src[g * 4], src[g * 4 + 1], src[g * 4 + 2], src[g * 4 + 3]));
                    __syncwarp();
                    const int unit = lane & 3;
                    const int colu = col0 + unit * 8;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rl = k * 8 + (lane >> 2);            // local row 0..31
                        const uint4 val = tc::lds_u4(stg + rl * 64 + ((unit ^ ((rl >> 1) & 3)) << 4));
                        // destination row of local row rl: fetched from the owning lane's registers
                        const long long dr = __shfl_sync(0xffffffffu, drow, rl);
                        const int okr = __shfl_sync(0xffffffffu, (int)row_ok, rl);
                        if (okr && colu < p.N) *reinterpret_cast<uint4*>(outp + dr * p.ldc + colu) = val;
                    }
                    }
                }
                if (VSW_GEMM_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { dbg_ld += c_b - c_a; dbg_math += c_c - c_b; dbg_st += GCLK() - c_c; }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) tc::mbar_arrive_cluster(tc::mapa_u32(tc::smem_u32(&tempty[acc]), 0));   // the leader's barrier
                else tc::mbar_arrive(&tempty[acc]);
            }
            if (VSW_GEMM_PROF && p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[4] += e_b - e_a; p.dbg[5] += GCLK() - e_b; p.dbg[6] += 1; p.dbg[8] += dbg_ld; p.dbg[9] += dbg_math; p.dbg[10] += dbg_st; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    if constexpr (TMA_OUT) { if (warp >= 2 && lane == 0) tc::bulk_wait_all(); }   // this thread's output tile stores
    // Reconverge the producer / MMA warps first: their lanes 1..31 must not sit in the (warp-aligned, blocking) cluster barrier
    // while lane 0 is still working -- that would starve lane 0 of issue slots.
    __syncwarp();
    tc::tc_fence_before();
    if constexpr (PAIR) tc::cluster_sync_all(); else __syncthreads();   // PAIR: no CTA leaves while its peer may still signal it
    if (warp == 1) {
        tc::tc_fence_after();
        if constexpr (PAIR) tc::tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// rows gather + per-batch scale:  out[m,:] = scale[b] * in[b*src_rows + map[r], :]  (zero row if map < 0)
template <bool F16>
__global__ void __launch_bounds__(256) gather_rows_kernel(const __nv_bfloat16* __restrict__ in,
                                                          __nv_bfloat16* __restrict__ out, const int32_t* __restrict__ map,
                                                          const float* __restrict__ scale, int M, int C,
                                                          int rows_per_batch, int src_rows_per_batch) {
    const int vec_per_row = C / 8;
    const long long total = (long long)M * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / vec_per_row), vcol = (int)(i - (long long)m * vec_per_row);
        const int b = m / rows_per_batch, r = m - b * rows_per_batch;
        const int s = map ? __ldg(map + r) : r;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (s >= 0) {
            ld8<F16>(in + ((long long)b * src_rows_per_batch + s) * C + vcol * 8, v);
            if (scale) {
                const float sc = __ldg(scale + b);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= sc;
            }
        }
        st8<F16>(out + (long long)m * C + vcol * 8, v);
    }
}

long long* gemm_dbg_buffer(int bn, bool amn, bool bmn, const TcParams& p) {
    static long long* dbg = nullptr;
    static bool init = false;
    static char last[128] = {0};
    if (!init) {
        init = true;
        if (VSW_GEMM_PROF && getenv("VSW_GEMM_DEBUG")) { cudaMalloc(&dbg, 1024); cudaMemset(dbg, 0, 1024); }
    }
    if (!dbg) return nullptr;
    long long h[128];
    cudaMemcpy(h, dbg, 1024, cudaMemcpyDeviceToHost);
    if (h[3]) fprintf(stderr, "[vsw gemm %s] tiles/cta=%lld  MMA thread: wait_tmem_empty=%lld wait_tma=%lld issue+rest=%lld | epilogue warp: wait_acc=%lld work=%lld (cycles per tile)\n",
                      last, h[3], h[0] / h[3], h[1] / h[3], (h[2] - h[1]) / h[3], h[4] / (h[6] + 1), h[5] / (h[6] + 1));
    if (h[3]) fprintf(stderr, "      epilogue split per tile: tmem_ld=%lld math=%lld stores=%lld\n", h[8] / (h[6] + 1), h[9] / (h[6] + 1), h[10] / (h[6] + 1));
    cudaMemset(dbg, 0, 1024);
    snprintf(last, sizeof(last), "BN=%d A_MN=%d B_MN=%d M=%d N=%d K=%d epi=%d", bn, (int)amn, (int)bmn, p.M, p.N, p.K, p.epi);
    return dbg;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool COLSUM = false, bool PAIR = false, bool F16 = false>
int launch_tc_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st, const CUtensorMap* outs) {
    constexpr int STAGE = A_BYTES + (PAIR ? BN / 2 : BN) * BK * 2;
    constexpr bool SIDE2 = (EPI == TE_DGRAD_MUL || EPI == TE_DGRAD_GELU || EPI == TE_RESIDUAL_ID);   // second per-warp staging buffer (+ 16 barriers) instead of one TMA stage
    constexpr int STAGES = (PAIR ? 6 : (BN <= 128 ? 5 : 4)) - (SIDE2 ? 1 : 0);
    constexpr size_t SMEM = (size_t)STAGES * STAGE + 1024 + 256 + (SIDE2 ? 2 : 1) * NUM_EPI_WARPS * 32 * 64 + (SIDE2 ? 128 : 0);
    static_assert(SMEM <= 227 * 1024, "GEMM shared-memory layout exceeds 227 KB");
    auto kern = tc_gemm_kernel<BN, A_MN, B_MN, STAGES, EPI, COLSUM, PAIR, F16>;
    const int dev = current_device();
    static bool configured[kMaxDevices] = {};  // per device (function attributes are); benign race: the attribute is idempotent
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) { set_error("tc gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
        configured[dev] = true;
    }
    const int total = p.n_tiles_m * p.n_tiles_n * p.splits;
    TcParams q = p;
    q.dbg = gemm_dbg_buffer(BN, A_MN, B_MN, p);
    if constexpr (PAIR) {
        // one CTA pair (cluster of 2 = one TPC) per work item slot
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // a persistent grid must be fully resident: not every TPC has both SMs enabled, so ask how many pairs fit
        static int max_pairs_dev[kMaxDevices] = {};   // cached per device
        if (!max_pairs_dev[dev]) {
            cfg.gridDim = dim3(kNumSMs);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) n = kNumSMs / 4;
            max_pairs_dev[dev] = n < kNumSMs / 2 ? n : kNumSMs / 2;
        }
        const int max_pairs = max_pairs_dev[dev];
        const int pairs = total < max_pairs ? total : max_pairs;
        cfg.gridDim = dim3(2 * pairs);
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, outs[0], outs[1], q);
        if (e != cudaSuccess) { set_error("tc gemm (pair): launch: %s", cudaGetErrorString(e)); return VSW_ERR_CUDA; }
        return check_launch("tc_gemm_pair");
    } else {
        const int grid = total < kNumSMs ? total : kNumSMs;
        kern<<<grid, NUM_THREADS, SMEM, st>>>(tmA, tmB, outs[0], outs[1], q);
        return check_launch("tc_gemm");
    }
}

// PAIR = CTA-pair (cta_group::2) 256 x 256 tiles; wgrad keeps single-CTA tiles (its column-sum warps need a local barrier)
template <int BN, bool A_MN, bool B_MN, bool PAIR, bool F16>
int launch_tc_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, cudaStream_t st, const CUtensorMap* outs) {
    if constexpr (A_MN && B_MN) {
        return p.colsum ? launch_tc_epi<BN, true, true, TE_PARTIAL, true, false, F16>(tmA, tmB, p, st, outs)
                        : launch_tc_epi<BN, true, true, TE_PARTIAL, false, false, F16>(tmA, tmB, p, st, outs);
    } else if constexpr (B_MN) {
        if (p.gelu_pre && p.epi == TE_DGRAD_MUL) return launch_tc_epi<BN, false, true, TE_DGRAD_MUL, false, PAIR, F16>(tmA, tmB, p, st, outs);
        return p.gelu_pre ? launch_tc_epi<BN, false, true, TE_DGRAD_GELU, false, PAIR, F16>(tmA, tmB, p, st, outs)
                          : launch_tc_epi<BN, false, true, TE_DGRAD, false, PAIR, F16>(tmA, tmB, p, st, outs);
    } else {
        switch (p.epi) {
            case TE_GELU: return launch_tc_epi<BN, false, false, TE_GELU, false, PAIR, F16>(tmA, tmB, p, st, outs);
            case TE_GELU_GRAD: return launch_tc_epi<BN, false, false, TE_GELU_GRAD, false, PAIR, F16>(tmA, tmB, p, st, outs);
            case TE_RESIDUAL: return launch_tc_epi<BN, false, false, TE_RESIDUAL, false, PAIR, F16>(tmA, tmB, p, st, outs);
            case TE_RESIDUAL_ID: return launch_tc_epi<BN, false, false, TE_RESIDUAL_ID, false, PAIR, F16>(tmA, tmB, p, st, outs);
            default: return launch_tc_epi<BN, false, false, TE_BIAS, false, PAIR, F16>(tmA, tmB, p, st, outs);
        }
    }
}
// f16: IEEE half operands (the same kernels with the other operand format and conversions)
template <int BN, bool A_MN, bool B_MN, bool PAIR = false>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, bool f16, cudaStream_t st,
              const CUtensorMap* outs = nullptr) {
    static const CUtensorMap none[2] = {};   // kernels whose epilogue does not use tensor-map stores never touch these
    if (!outs) outs = none;
    return f16 ? launch_tc_t<BN, A_MN, B_MN, PAIR, true>(tmA, tmB, p, st, outs) : launch_tc_t<BN, A_MN, B_MN, PAIR, false>(tmA, tmB, p, st, outs);
}

// tensor maps of the output (and the optional second output) for the epilogue's 32 x 32 chunk stores: 64-byte swizzle
static bool make_out_maps(CUtensorMap* outs, const void* out, const void* aux, int M, int N, long long ld) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, strides[2] = {1, (uint64_t)ld};
    const uint32_t box[2] = {32, 32};
    if (!make_tmap_nd_bf16(&outs[0], out, 2, dims, strides, box, 64, 128)) return false;
    if (aux) return make_tmap_nd_bf16(&outs[1], aux, 2, dims, strides, box, 64, 128);
    outs[1] = outs[0];
    return true;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// N-tile: 256 halves the A-operand re-reads from L2 (the main loop is L2-bandwidth bound at 128x128), as long as
// the tile count still fills the GPU
static int pick_bn(int M, int N) {
    if (N % 256 != 0) return 128;
    const long long tiles256 = (long long)ceil_div(M, BM) * (N / 256);
    return tiles256 >= kNumSMs ? 256 : 128;
}

// CTA-pair tiles when the 256-wide tile is in use, the reduction is long enough for the main loop to matter and there
// are enough 256 x 256 tiles for the 74 pairs
static bool use_pair(int M, int N, int K, int BN) {
    if (BN != 256) return false;
    return K >= 256 && (long long)ceil_div(M, 2 * BM) * (N / 256) >= kNumSMs / 2;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
int tc_linear(const TcLinearArgs& a, cudaStream_t st) {
    if ((a.K % 8) || (a.N % 8) || !aligned16(a.x) || !aligned16(a.w) || !aligned16(a.y) ||
        (a.bias && !aligned16(a.bias)) || (a.res && !aligned16(a.res)) || (a.aux_out && !aligned16(a.aux_out))) {
        set_error("tcgen05 linear: needs K %% 8 == 0, N %% 8 == 0 and 16-byte aligned pointers (M=%d N=%d K=%d)", a.M,
                  a.N, a.K);
        return VSW_ERR_UNSUPPORTED;
    }
    CUtensorMap tmA, tmB;
    const int BN = pick_bn(a.M, a.N);
    const bool pair = use_pair(a.M, a.N, a.K, BN);
    if (!make_tmap_2d_bf16(&tmA, a.x, a.M, a.K, a.K, BM, BK)) return VSW_ERR_CUDA;
    if (!make_tmap_2d_bf16(&tmB, a.w, a.N, a.K, a.K, pair ? BN / 2 : BN, BK)) return VSW_ERR_CUDA;
    TcParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.n_tiles_m = ceil_div(a.M, pair ? 2 * BM : BM); p.n_tiles_n = ceil_div(a.N, BN); p.splits = 1; p.k_per_split = ceil_div(a.K, BK) * BK;
    p.epi = a.epi == VSW_EPI_BIAS ? TE_BIAS : (a.epi == VSW_EPI_GELU ? TE_GELU : (a.epi == VSW_EPI_GELU_GRAD ? TE_GELU_GRAD : TE_RESIDUAL));
    p.bias = (const __nv_bfloat16*)a.bias; p.out = (__nv_bfloat16*)a.y; p.aux_out = (__nv_bfloat16*)a.aux_out;
    p.res = (const __nv_bfloat16*)a.res; p.rowmap = a.rowmap; p.rowscale = a.rowscale;
    p.rows_per_batch = a.rows_per_batch; p.dst_rows_per_batch = a.dst_rows_per_batch; p.ldc = a.N;
    const bool f16 = a.dtype == VSW_F16;
    // residual epilogue without a row map (fc2 + residual, video_swin.py:261): output rows = tile rows
    if (p.epi == TE_RESIDUAL && !a.rowmap && a.rows_per_batch == a.dst_rows_per_batch) p.epi = TE_RESIDUAL_ID;
    CUtensorMap outs[2];
    if (!make_out_maps(outs, a.y, p.epi == TE_RESIDUAL_ID ? a.res : a.aux_out, a.M, a.N, a.N)) return VSW_ERR_CUDA;
    if (pair) return launch_tc<256, false, false, true>(tmA, tmB, p, f16, st, outs);
    return BN == 256 ? launch_tc<256, false, false>(tmA, tmB, p, f16, st, outs) : launch_tc<128, false, false>(tmA, tmB, p, f16, st, outs);
}

// ---------------------------------------------------------------------------------------------
// dgrad: dx (M x K) = A (M x N) w (N x K), A = optional gather/scale of dy
// ---------------------------------------------------------------------------------------------
int tc_dgrad(const TcDgradArgs& a, cudaStream_t st) {
    if ((a.K % 8) || (a.N % 8) || !aligned16(a.dy) || !aligned16(a.w) || !aligned16(a.dx) ||
        (a.gelu_pre && !aligned16(a.gelu_pre)) || (a.a_out && !aligned16(a.a_out))) {
        set_error("tcgen05 dgrad: needs K %% 8 == 0, N %% 8 == 0 and 16-byte aligned pointers");
        return VSW_ERR_UNSUPPORTED;
    }
    const void* A = a.dy;
    const bool f16 = a.dtype == VSW_F16;
    if (a.a_rowmap || a.a_rowscale) {
        if (!a.a_out) { set_error("tcgen05 dgrad: gathered A needs the a_out buffer"); return VSW_ERR_UNSUPPORTED; }
        const long long total = (long long)a.M * (a.N / 8);
        int blocks = (int)((total + 255) / 256 < (long long)kNumSMs * 16 ? (total + 255) / 256 : (long long)kNumSMs * 16);
        if (f16) gather_rows_kernel<true><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)a.dy, (__nv_bfloat16*)a.a_out, a.a_rowmap,
                                                                 a.a_rowscale, a.M, a.N, a.rows_per_batch, a.src_rows_per_batch);
        else gather_rows_kernel<false><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)a.dy, (__nv_bfloat16*)a.a_out, a.a_rowmap,
                                                              a.a_rowscale, a.M, a.N, a.rows_per_batch, a.src_rows_per_batch);
        int rc = check_launch("gather_rows");
        if (rc) return rc;
        A = a.a_out;
    }
    CUtensorMap tmA, tmB;
    if (!make_tmap_2d_bf16(&tmA, A, a.M, a.N, a.N, BM, BK)) return VSW_ERR_CUDA;       // K-major over N
    if (!make_tmap_2d_bf16(&tmB, a.w, a.N, a.K, a.K, 64, 64)) return VSW_ERR_CUDA;     // MN-major: rows = reduction
    TcParams p{};
    p.M = a.M; p.N = a.K; p.K = a.N;
    const int BN = pick_bn(a.M, a.K);
    const bool pair = use_pair(a.M, a.K, a.N, BN);
    p.n_tiles_m = ceil_div(a.M, pair ? 2 * BM : BM); p.n_tiles_n = ceil_div(a.K, BN); p.splits = 1; p.k_per_split = ceil_div(a.N, BK) * BK;
    p.epi = (a.gelu_pre && a.pre_is_grad) ? TE_DGRAD_MUL : TE_DGRAD;
    p.out = (__nv_bfloat16*)a.dx; p.gelu_pre = (const __nv_bfloat16*)a.gelu_pre; p.ldc = a.K;
    CUtensorMap outs[2];
    if (!make_out_maps(outs, a.dx, a.gelu_pre, a.M, a.K, a.K)) return VSW_ERR_CUDA;   // second map: the side tensor (GELU' or pre-activation), same shape as dx
    if (pair) return launch_tc<256, false, true, true>(tmA, tmB, p, f16, st, outs);
    return BN == 256 ? launch_tc<256, false, true>(tmA, tmB, p, f16, st, outs) : launch_tc<128, false, true>(tmA, tmB, p, f16, st, outs);
}

// ---------------------------------------------------------------------------------------------
// wgrad: dw (N x K) = dy (M x N)^T x (M x K)
// ---------------------------------------------------------------------------------------------
static int wgrad_bn(int K) { return K % 256 == 0 ? 256 : 128; }
static void tc_wgrad_split(int M, int N, int K, int* splits, int* m_per_split) {
    // work items = output tiles x splits of the reduction over the M rows.  Pick the split count (each split >= 512
    // rows, at most ~4 items per SM) that wastes the fewest SM-slots in the last wave of the persistent grid.
    const long long tiles = (long long)ceil_div(N, BM) * ceil_div(K, wgrad_bn(K));
    long long smax = (M + 511) / 512;
    const long long cap = (4LL * kNumSMs + tiles - 1) / tiles;
    if (smax > cap) smax = cap;
    if (smax < 1) smax = 1;
    // efficiency of the persistent grid = items / (waves x SMs).  Among the split counts within 1.5 % of the best efficiency
    // take the SMALLEST one that still gives every CTA two items (so one epilogue overlaps a main loop): the fp32 partials
    // written by the epilogue and re-read by the reduce kernel grow linearly with the split count.
    auto eff = [&](long long s) {
        const long long items = tiles * s;
        const long long waves = (items + kNumSMs - 1) / kNumSMs;
        return (double)items / (double)(waves * kNumSMs);
    };
    double best_eff = -1.0;
    for (long long s = 1; s <= smax; ++s) best_eff = eff(s) > best_eff ? eff(s) : best_eff;
    long long best = smax;
    for (long long s = 1; s <= smax; ++s) {
        if (eff(s) < best_eff - 0.015) continue;
        if (tiles * s < 2LL * kNumSMs - kNumSMs / 8 && s < smax) continue;   // fewer than ~2 waves
        best = s;
        break;
    }
    int mps = (int)((M + best - 1) / best);
    mps = (mps + BK - 1) / BK * BK;
    *m_per_split = mps;
    *splits = (M + mps - 1) / mps;
}

// fixed-order reduction of the split partials: dw (elemsW values) and, optionally, db (elemsB values)
template <typename TO>
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long long elemsW, TO* __restrict__ dw,
                                    const float* __restrict__ cpart, int elemsB, TO* __restrict__ db) {
    const long long total = elemsW + elemsB;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        if (i < elemsW) {
            for (int k = 0; k < splits; ++k) s += partial[(long long)k * elemsW + i];
            dw[i] = from_f<TO>(s);
        } else {
            const long long j = i - elemsW;
            for (int k = 0; k < splits; ++k) s += cpart[(long long)k * elemsB + j];
            db[j] = from_f<TO>(s);
        }
    }
}

size_t tc_wgrad_workspace(int M, int N, int K) {
    int s, mps;
    tc_wgrad_split(M, N, K, &s, &mps);
    return (size_t)s * N * K * sizeof(float) + (size_t)s * N * sizeof(float);
}

// dw = dy^T x and (db != nullptr) db = column sums of dy, both from ONE pass over dy
int tc_wgrad(const void* dy, const void* x, void* dw, void* db, int M, int N, int K, int dtype, int grad_dtype, void* ws,
             size_t ws_bytes, cudaStream_t st) {
    if ((K % 8) || (N % 8) || !aligned16(dy) || !aligned16(x) || !aligned16(ws)) {
        set_error("tcgen05 wgrad: needs K %% 8 == 0, N %% 8 == 0 and 16-byte aligned pointers");
        return VSW_ERR_UNSUPPORTED;
    }
    int splits, mps;
    tc_wgrad_split(M, N, K, &splits, &mps);
    if (ws_bytes < tc_wgrad_workspace(M, N, K)) { set_error("tcgen05 wgrad: workspace too small"); return VSW_ERR_WORKSPACE; }
    CUtensorMap tmA, tmB;
    if (!make_tmap_2d_bf16(&tmA, dy, M, N, N, 64, 64)) return VSW_ERR_CUDA;  // MN-major A: rows = reduction (m)
    if (!make_tmap_2d_bf16(&tmB, x, M, K, K, 64, 64)) return VSW_ERR_CUDA;   // MN-major B
    TcParams p{};
    p.M = N; p.N = K; p.K = M;
    const int BN = wgrad_bn(K);
    p.n_tiles_m = ceil_div(N, BM); p.n_tiles_n = ceil_div(K, BN); p.splits = splits; p.k_per_split = mps;
    p.epi = TE_PARTIAL; p.partial = (float*)ws; p.ldc = K;
    float* cpart = (float*)ws + (size_t)splits * N * K;
    p.colsum = db ? cpart : nullptr;
    const bool f16 = dtype == VSW_F16;
    int rc = BN == 256 ? launch_tc<256, true, true>(tmA, tmB, p, f16, st) : launch_tc<128, true, true>(tmA, tmB, p, f16, st);
    if (rc) return rc;
    const long long elemsW = (long long)N * K, total = elemsW + (db ? N : 0);
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    VSW_DISPATCH_DTYPE(grad_dtype, TO,
                       (wgrad_reduce_kernel<TO><<<blocks, 256, 0, st>>>((const float*)ws, splits, elemsW, (TO*)dw, cpart,
                                                                         db ? N : 0, (TO*)db)));
    return check_launch("wgrad_reduce");
}

}  // namespace vsw
