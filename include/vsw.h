/* vsw.h -- C ABI of libvsw_b200.so: hand-written sm_100a kernels for the Video-Swin 3D hot path
 * of tsujuifu/pytorch_empirical-mvm (reference: visbackbone/video_swin.py; citations below are
 * file:line relative to the reference tree).
 *
 * The reference has no FFI layer: its "operator interface" for this path is the set of torch ops
 * inside visbackbone/video_swin.py.  Every entry point below replaces one group of those ops and is
 * what a binding for this path has to bind (INTEGRATION.md shows the ctypes stub).  The last two
 * sections cover what sits directly on either side of the encoder: the EncVideo
 * tail (model.py:57-76) and the MVM patch masking / masked-L1 loss (main_pretrain.py:355-362, 520-522).
 *
 * Conventions (SURVEY.md section 8b "Lower"):
 *   - plain C symbols, raw device pointers + sizes, no torch types;
 *   - return 0 (VSW_OK) or a negative vsw_status; never throws, never aborts; the message of the
 *     last failure on the calling thread is available through vsw_last_error();
 *   - the CALLER owns every buffer (including workspaces), the library never allocates device
 *     memory, never synchronises and never changes the current device;
 *   - every launch goes to the caller's stream (a cudaStream_t passed as void*);
 *   - re-entrant, callable from any host thread (autograd engine thread included);
 *   - deterministic: no floating-point atomics, fixed reduction orders.
 *   - "dtype" is the storage type of activations/weights (vsw_dtype); arithmetic is always fp32
 *     (or 16-bit x 16-bit -> fp32 on the tensor cores for VSW_BF16 and VSW_F16 GEMM / attention tiles);
 *     statistics (mean/rstd/lse) and index maps are fp32 / int32 / uint8 regardless of dtype.
 *
 * Token layout: channels-last.  An activation is (B, T, C) with T = D*H*W tokens in row-major
 * (d,h,w) order; a windowed activation is (B, nW*N, C) with windows row-major over
 * (Dp/wd, Hp/wh, Wp/ww) and slots row-major over (wd, wh, ww)  [video_swin.py:84-88].
 */
#ifndef VSW_H_
#define VSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum { VSW_F32 = 0, VSW_BF16 = 1, VSW_F16 = 2 } vsw_dtype;

typedef enum {
    VSW_OK = 0,
    VSW_ERR_ARG = -1,      /* bad shape / null pointer / misaligned pointer */
    VSW_ERR_DTYPE = -2,    /* dtype not supported by this entry point */
    VSW_ERR_UNSUPPORTED = -3, /* geometry not supported by this kernel (caller must not fall back silently) */
    VSW_ERR_CUDA = -4,     /* a CUDA runtime/driver call failed */
    VSW_ERR_WORKSPACE = -5 /* workspace too small */
} vsw_status;

/* epilogues of vsw_linear_fwd */
typedef enum {
    VSW_EPI_BIAS = 0,       /* y = acc + bias                                        (qkv, fc-like)          */
    VSW_EPI_GELU = 1,       /* y = gelu_erf(acc + bias); optional pre-activation out  [video_swin.py:76-77]   */
    VSW_EPI_RESIDUAL = 2,   /* y[dst] = res[dst] + rowscale[b] * (acc + bias), dst via optional row map
                               [proj + window_reverse + roll back + residual, video_swin.py:170,233-241,256;
                                fc2 + residual, video_swin.py:261]                                            */
    VSW_EPI_GELU_GRAD = 3   /* y = gelu_erf(acc + bias) and aux_out = gelu_erf'(acc + bias): the training form of the
                               fc1 epilogue -- the backward then needs one multiply per element
                               (vsw_linear_dgrad_mul) instead of re-evaluating the GELU derivative             */
} vsw_epilogue;

/* GEMM back ends (vsw_set_gemm_backend): the CUDA-core fp32 kernel is the only one for VSW_F32. */
typedef enum { VSW_GEMM_AUTO = 0, VSW_GEMM_SIMT = 1, VSW_GEMM_TCGEN05 = 2 } vsw_gemm_backend;

int vsw_version(void);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its length */
int vsw_last_error(char* buf, size_t n);
/* selects the kernel family for VSW_BF16 GEMMs/attention; returns the previous value */
int vsw_set_gemm_backend(int backend);
int vsw_get_gemm_backend(void);
/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
long long vsw_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Index maps -- integer domain, bit-exact with the reference.
 * ---------------------------------------------------------------------------------------------- */

/* Shift + window-partition gather map and shift-mask region ids.
 * Replaces torch.roll + window_partition (video_swin.py:84-88, 220-229) and the region counter of
 * compute_mask (video_swin.py:294-302).
 *   grid (D,H,W): unpadded token grid; (wd,wh,ww)/(sd,sh,sw): EFFECTIVE window/shift (after
 *   get_window_size, video_swin.py:95-108).  Padded grid = ceil to window multiples.
 *   gather [nW*N] int32: flat (d*H+h)*W+w index into the UNPADDED grid of the token read by window
 *                        slot (w,n), or -1 where the slot lies in the post-norm zero padding.
 *   region [nW*N] uint8: cnt = 9*rd + 3*rh + rw of the slot in the shifted frame.
 * Either output pointer may be NULL. */
int vsw_window_maps(int D, int H, int W, int wd, int wh, int ww, int sd, int sh, int sw,
                    int32_t* gather, uint8_t* region, void* stream);

/* relative_position_index buffer (N x N int64), video_swin.py:123-137 */
int vsw_rel_pos_index(int wd, int wh, int ww, int64_t* out, void* stream);

/* dense additive mask (nW,N,N) of {0,-100} from region ids, video_swin.py:303-307 */
int vsw_shift_mask(const uint8_t* region, int nW, int N, void* out, int dtype, void* stream);

/* PatchMerging gather map (video_swin.py:276-284): out [D*ceil(H/2)*ceil(W/2)*4] int32, flat token
 * index in the (D,H,W) grid of channel group g in order (dh,dw)=(0,0),(1,0),(0,1),(1,1); -1 = pad */
int vsw_merge_map(int D, int H, int W, int32_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm family (HBM-bound, vectorised, warp-shuffle).  eps = 1e-5 in the reference.
 * ---------------------------------------------------------------------------------------------- */

/* y[b,r,:] = LN(x[b, map[r], :]) * gamma + beta  for r < Tout;  map == NULL -> identity (Tout==Tin);
 * map[r] < 0 -> y row = 0 (zero padding AFTER the norm, video_swin.py:211-217).
 * mean/rstd [B*Tout] fp32 are written when non-NULL (needed by the backward).
 * norm1 + pad + roll + window_partition: video_swin.py:211-229; norm2: :248; final norm: :479;
 * patch_embed.norm: :401-405.  out_dtype may differ from dtype (final norm under autocast). */
int vsw_ln_fwd(const void* x, const void* gamma, const void* beta, const int32_t* map,
               void* y, float* mean, float* rstd,
               int B, int Tin, int Tout, int C, float eps, int dtype, int out_dtype, void* stream);

/* Backward of vsw_ln_fwd for the rows r < Tout:
 *   dx[b, map[r], :] = (dres ? dres[b, map[r], :] : 0) + LN'(dy[b,r,:]) ;   rows with map[r] < 0 skipped.
 * dy_dtype is the storage type of dy (fp32 for the final norm under autocast).
 * dgamma/dbeta (fp32 [C]) get the column reductions sum_r dy*xhat / sum_r dy, computed through the
 * fp32 workspace `ws` of at least vsw_ln_bwd_workspace(C) bytes (two-pass, fixed order). */
size_t vsw_ln_bwd_workspace(int C);
int vsw_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
               const int32_t* map, const void* dres, void* dx, float* dgamma, float* dbeta,
               int B, int Tin, int Tout, int C, int dtype, int dy_dtype, void* ws, size_t ws_bytes,
               void* stream);

/* Stochastic-depth factors of DropPath (video_swin.py:46-54): out[b] = floor(keep + u[b]) / keep (fp32), where u holds the
 * n uniform samples the caller drew (torch.rand in the activation dtype, so the random stream matches the reference) and the
 * sum is rounded to that dtype before the floor, as the reference's `keep_prob + torch.rand(...)` is. */
int vsw_drop_path_scale(const void* u, float keep, float* out, int n, int dtype, void* stream);

/* vsw_ln_bwd with dgamma / dbeta written in `dparam_dtype` (the parameters' dtype) instead of fp32: the fixed-order fp32 column
 * reduction is rounded once by the finish kernel, so no cast kernel has to follow. */
int vsw_ln_bwd_ex(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                  const int32_t* map, const void* dres, void* dx, void* dgamma, void* dbeta,
                  int B, int Tin, int Tout, int C, int dtype, int dy_dtype, int dparam_dtype, void* ws, size_t ws_bytes,
                  void* stream);

/* Residual add with window-reverse scatter, for a residual stream kept wider than the branch (torch.autocast keeps
 * `shortcut + drop_path(x)` in fp32 while the Linear outputs are 16-bit, video_swin.py:256, 261):
 *   out[b, map[r], :] = x[b, map[r], :] + rowscale[b] * y[b, r, :]     (rows with map[r] < 0 skipped; map NULL = identity)
 * x / out (B, dst_rows_per_batch, C) in `dtype`, y (B, rows_per_batch, C) in `y_dtype`; C % 4 == 0. */
int vsw_residual_add(const void* x, const void* y, const int32_t* rowmap, const float* rowscale, void* out, int B,
                     int rows_per_batch, int dst_rows_per_batch, int C, int dtype, int y_dtype, void* stream);

/* PatchMerging front half (video_swin.py:276-286): y[b,r,:] = LN_{4C}(concat_g x[b, map4[r,g], :]),
 * map4 < 0 -> zero input (padding BEFORE the norm).  y is (B, Tout, 4C). */
int vsw_merge_ln_fwd(const void* x, const void* gamma, const void* beta, const int32_t* map4,
                     void* y, float* mean, float* rstd,
                     int B, int Tin, int Tout, int C, float eps, int dtype, void* stream);
int vsw_merge_ln_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                     const int32_t* map4, void* dx, float* dgamma, float* dbeta,
                     int B, int Tin, int Tout, int C, int dtype, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Linear layers  (nn.Linear convention: w is (N, K) row-major, y = x w^T + bias).
 * ---------------------------------------------------------------------------------------------- */

/* y = epilogue(x[M,K] w[N,K]^T + bias[N]).
 *   VSW_EPI_GELU:     aux_out (M,N) receives the pre-activation when non-NULL.
 *   VSW_EPI_GELU_GRAD: aux_out (M,N), required, receives gelu'(pre-activation).
 *   VSW_EPI_RESIDUAL: rows are grouped in batches of rows_per_batch; source row m = b*rows_per_batch + r
 *                     is written to destination row b*dst_rows_per_batch + (rowmap ? rowmap[r] : r),
 *                     skipped when rowmap[r] < 0;  y[dst] = res[dst] + (rowscale ? rowscale[b] : 1) * (acc+bias).
 *                     rowscale[b] is the drop-path factor floor(keep+U)/keep (video_swin.py:46-54). */
int vsw_linear_fwd(const void* x, const void* w, const void* bias, void* y,
                   int M, int N, int K, int epilogue,
                   void* aux_out, const void* res, const int32_t* rowmap, const float* rowscale,
                   int rows_per_batch, int dst_rows_per_batch,
                   int dtype, void* stream);

/* dx[M,K] = A w[N,K]   where A[m,:] = a_scale * dy[src(m), :] (*) gelu'(pre[m,:]) :
 *   a_rowmap (optional): A row m = b*rows_per_batch + r reads dy row b*src_rows_per_batch + a_rowmap[r]
 *                        (zero row if < 0)  -- the transpose of the RESIDUAL scatter epilogue;
 *   a_rowscale (optional): per-batch factor (drop-path);
 * gelu_pre (optional, (M,K)): dx = (dy w) (*) gelu'(gelu_pre)   -- fc2 backward fused with the GELU backward.
 * When a_rowmap/a_rowscale are given and a_out != NULL the gathered+scaled A is also written to a_out (M,N)
 * so that the weight-gradient GEMM can consume it. */
int vsw_linear_dgrad(const void* dy, const void* w, void* dx,
                     int M, int N, int K,
                     const int32_t* a_rowmap, const float* a_rowscale, int rows_per_batch, int src_rows_per_batch,
                     void* a_out, const void* gelu_pre,
                     int dtype, void* stream);

/* Same as vsw_linear_dgrad with a plain elementwise factor:  dx = (A w) (*) mul[M,K]   (mul = the gelu' tensor written by
 * VSW_EPI_GELU_GRAD; autograd of fc2 followed by the GELU, video_swin.py:76-79). */
int vsw_linear_dgrad_mul(const void* dy, const void* w, void* dx,
                         int M, int N, int K,
                         const int32_t* a_rowmap, const float* a_rowscale, int rows_per_batch, int src_rows_per_batch,
                         void* a_out, const void* mul,
                         int dtype, void* stream);

/* dw[N,K] = dy[M,N]^T x[M,K]; db[N] = sum_m dy[m,:] (db may be NULL).  Split over M with an fp32
 * workspace (vsw_linear_wgrad_workspace bytes) and a fixed-order second pass.  dw/db are written in
 * grad_dtype (VSW_F32 for fp32 master parameters under autocast). */
size_t vsw_linear_wgrad_workspace(int M, int N, int K);
int vsw_linear_wgrad(const void* dy, const void* x, void* dw, void* db,
                     int M, int N, int K, int dtype, int grad_dtype,
                     void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused window attention (video_swin.py:149-169): softmax(q*scale k^T + bias + mask) v, per
 * (window, head); scores never touch HBM.
 *   qkv  (B_, N, 3, nH, hd)   B_ = B*nW, windows of one clip contiguous, batch slowest
 *   out  (B_, N, nH*hd)
 *   lse  (B_, nH, N) fp32 log-sum-exp of the biased+masked scores (saved for the backward)
 *   bias_table (L, nH) in dtype; bias[h,i,j] = bias_table[rowcode[i] + colcode[j], h] where
 *       rowcode[i] = index[i,0], colcode[j] = index[0,j] - index[0,0] (exact because
 *       relative_position_index is translation invariant; includes the [:N,:N] slice quirk, :155)
 *   region (nW, N) uint8 or NULL: additive mask -100 where region[w,i] != region[w,j] (video_swin.py:303-307)
 *   dense_mask (nW,N,N) in dtype or NULL: arbitrary additive mask (used instead of region when given)
 * ---------------------------------------------------------------------------------------------- */
int vsw_window_attn_fwd(const void* qkv, const void* bias_table, const int32_t* rowcode, const int32_t* colcode,
                        const uint8_t* region, const void* dense_mask,
                        void* out, float* lse,
                        int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims,
                        int dtype, void* stream);

/* Recompute-based backward.  dqkv (B_,N,3,nH,hd); dbias_table (L,nH) fp32, OVERWRITTEN
 * with the reduction over batch and windows, via workspace partials (fixed order).
 * window_dims (both directions) = wd_eff | wh << 8 | ww << 16, any field 0 if unknown -- layout hints only, validated
 * against the codes inside the kernels:  wd_eff = depth of the EFFECTIVE window (N = wd_eff * tokens per plane; lets the
 * backward order keys plane-minor);  wh, ww = height / width of the CONFIGURED window whose relative_position_index
 * produced rowcode/colcode.  With wh and ww given (ww <= 8, N <= 448, head_dim 32) the second-generation tcgen05 kernels
 * run for VSW_BF16 and VSW_F16: they keep a per-head table T[w_i][d_i-d_j][h_i-h_j][8 key slots] on chip, so the bias of 8
 * consecutive keys is one 16-byte load; without the hint the first-generation bf16 kernels or the CUDA-core kernels run.
 * Same results either way (the kernels check the hint against rowcode / colcode; a wrong hint yields NaN, never a wrong
 * finite result). */
size_t vsw_window_attn_bwd_workspace(int B_, int N, int nH, int hd, int L);
int vsw_window_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                        const void* bias_table, const int32_t* rowcode, const int32_t* colcode,
                        const uint8_t* region, const void* dense_mask,
                        void* dqkv, float* dbias_table,
                        int B_, int nW, int N, int nH, int hd, int L, float scale, int window_dims,
                        int dtype, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PatchEmbed3D (video_swin.py:390-400): zero-pad H,W up to patch multiples, append one zero frame,
 * Conv3d(k=(pd,ph,pw), stride=(1,ph,pw)).  The convolution is run as  im2col -> vsw_linear_fwd
 * (K = Cin*pd*ph*pw = 96), so forward, weight gradient and input gradient all go through the GEMM
 * family above; the patch_norm LayerNorm is a vsw_ln_fwd on the result.
 *   x   (B,Cin,D,H,W) in x_dtype (fp32 clips are converted on load)
 *   col (B*Dout*Hp*Wp, Cin*pd*ph*pw) in dtype, column order (c,dt,dy,dx) = Conv3d weight order,
 *       Dout = D + 2 - pd, Hp = ceil(H/ph), Wp = ceil(W/pw); rows are tokens in (b,d,h',w') order.
 * ---------------------------------------------------------------------------------------------- */
int vsw_patch_im2col(const void* x, void* col, int B, int Cin, int D, int H, int W, int pd, int ph, int pw,
                     int x_dtype, int dtype, void* stream);
/* transpose of vsw_patch_im2col: dx (B,Cin,D,H,W) in x_dtype = sum of the <= pd column entries per element */
int vsw_patch_col2im(const void* dcol, void* dx, int B, int Cin, int D, int H, int W, int pd, int ph, int pw,
                     int x_dtype, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * EncVideo tail (model.py:57-76) -- the consumer of the Swin output in every VIOLET step.
 * After `fc` (a vsw_linear_fwd, model.py:42) each frame's h*w tokens get a class row in front, the position and
 * frame-length (or frame-order) embeddings are added and the rows are LayerNorm-ed (eps 1e-5):
 *   pre[b,t,p,:]      = (p == 0 ? emb_cls : f[b,t,p-1,:]) + emb_pos[p,:] + (odr && odr[b,t] != t ? emb_odr : emb_len[t,:])
 *   out[b, t*P+p, :]  = LN(pre[b,t,p,:]) * gamma + beta          P = 1 + hw          (model.py:57-70)
 *   m_img[b, t*P+p]   = vt_mask ? vt_mask[b,t,p] : 1             int64                (model.py:72-76)
 *   f (B*Tn*hw, C) in dtype; out (B*Tn*P, C) in out_dtype (dtype, or fp32 as under autocast);
 *   emb_cls (C), emb_pos (pos_rows >= P, C), emb_len (len_rows >= Tn, C), emb_odr (C), gamma, beta (C): fp32;
 *   odr (B*Tn) int32 or NULL; vt_mask (B*Tn*P) int64 or NULL; m_img (B*Tn*P) int64 or NULL;
 *   mean / rstd (B*Tn*P) fp32 or both NULL (saved for the backward).
 * ---------------------------------------------------------------------------------------------- */
int vsw_enc_video_tail_fwd(const void* f, const float* emb_cls, const float* emb_pos, const float* emb_len,
                           const float* emb_odr, const int32_t* odr, const float* gamma, const float* beta,
                           const int64_t* vt_mask, void* out, int64_t* m_img, float* mean, float* rstd,
                           int B, int Tn, int hw, int C, int pos_rows, int len_rows, float eps,
                           int dtype, int out_dtype, void* stream);

/* Backward of the above (pre is recomputed from the inputs).  df (B*Tn*hw, C) in dtype or NULL; the parameter gradients
 * are fp32 and OVERWRITTEN: demb_cls (C), demb_pos (pos_rows, C; rows >= P get 0), demb_len (len_rows, C; rows >= Tn
 * get 0), demb_odr (C; 0 when odr == NULL), dgamma, dbeta (C).  Fixed reduction orders through the fp32 workspace. */
size_t vsw_enc_video_tail_bwd_workspace(int B, int Tn, int hw, int C);
int vsw_enc_video_tail_bwd(const void* dy, const void* f, const float* emb_cls, const float* emb_pos,
                           const float* emb_len, const float* emb_odr, const int32_t* odr, const float* gamma,
                           const float* mean, const float* rstd, void* df, float* demb_cls, float* demb_pos,
                           float* demb_len, float* demb_odr, float* dgamma, float* dbeta,
                           int B, int Tn, int hw, int C, int pos_rows, int len_rows, int dtype, int dy_dtype,
                           void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MVM masking and loss on either side of the encoder (main_pretrain.py).
 * ---------------------------------------------------------------------------------------------- */

/* Patch masking of the clip fed to the student encoder (main_pretrain.py:355-362):
 *   img_out[f,c,y,x] = img[f,c,y,x] * (1 - cov[f, y/ps, x/ps]);   mvm_mask[f,c,y,x] = cov[f, y/ps, x/ps]  (fp32)
 * img / img_out (frames, Cin, H, W) in dtype, may alias (the reference masks in place); cov (frames, H/ps, W/ps) uint8 in
 * {0,1}; frames = B*T.  img_out may be NULL (only the mask is wanted) and mvm_mask may be NULL (only the clip). */
int vsw_block_mask_apply(const void* img, const uint8_t* cov, void* img_out, float* mvm_mask,
                         int frames, int Cin, int H, int W, int ps, int dtype, void* stream);

/* Masked L1 between a prediction and the teacher encoder's tokens (main_pretrain.py:520-522, 427-428, 534-535):
 *   loss = sum_{r,c} |pred[r,c] - target[r,c]| * row_weight[r] / (sum_r row_weight[r] + 1e-5) / in_c        (fp32 scalar)
 * pred (rows, C) in dtype, target (rows, C) in target_dtype (dtype or fp32), row_weight (rows) fp32 (the patch coverage:
 * max_pool2d(mvm_mask, ps).sum(1) / 3).  loss and msum (= sum_r row_weight) are DEVICE scalars; rows with weight 0 are
 * never read.  Two-stage fixed-order reduction through ws (vsw_masked_l1_workspace() bytes). */
size_t vsw_masked_l1_workspace(void);
int vsw_masked_l1_fwd(const void* pred, const void* target, const float* row_weight, float* loss, float* msum,
                      long long rows, int C, float in_c, int dtype, int target_dtype,
                      void* ws, size_t ws_bytes, void* stream);
/* dpred[r,c] = sign(pred - target) * row_weight[r] * dloss / (msum + 1e-5) / in_c ;  dloss, msum: device scalars */
int vsw_masked_l1_bwd(const void* pred, const void* target, const float* row_weight, const float* msum,
                      const float* dloss, void* dpred, long long rows, int C, float in_c,
                      int dtype, int target_dtype, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* VSW_H_ */
