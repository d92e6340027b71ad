/* cti_sm100.h -- C ABI of libcti_sm100.so, the sm_100a (B200) kernels behind the drop-in
 * FCNet / TCNet / TriAttention / BCNet / BiAttention modules.
 *
 * The reference (aioz-ai/ICCV19_VQA-CTI) is pure PyTorch: its "FFI" for this path is the set of
 * ATen calls its five hot-path files make.  Each entry point below names the reference call
 * site(s) it replaces (file:line under the reference tree).
 *
 * Conventions
 *   - Plain C: raw device pointers and sizes only; no torch / ATen / pybind types.
 *   - bf16 buffers are passed as `const void*` / `void*` (2 bytes per element, row-major).
 *   - The caller owns every buffer (inputs, outputs, saved tensors, workspace) and keeps it
 *     alive until `stream` has passed the call.  The library allocates nothing persistent.
 *   - All launches are asynchronous on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - Return value: 0 = ok; negative = argument/shape/alignment error, nothing was launched;
 *     positive = the cudaError_t reported by the launch.  `cti_last_error()` returns a
 *     thread-local message for the last non-zero return on this thread.  Nothing throws or exits.
 *   - Buffers named `*_accum` are accumulated into (fp32 atomics or TMA reduce-add) and must be zeroed by the caller.
 *   - The device is the current device of the calling thread (one process per GPU for DP).
 */
#ifndef CTI_SM100_H_
#define CTI_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library version: major * 10000 + minor * 100 + patch. */
int cti_version(void);
/* Message for the last failing call on this thread ("" if none). */
const char* cti_last_error(void);

/* ---- row cast + zero-row mask -------------------------------------------------------------
 * out[r,:] = bf16(x[r,:]);  rowmask[r] = (sum_c |x[r,c]| == 0)   (rowmask may be NULL)
 * replaces: src/attention.py:55 and :36  `(0 == v.abs().sum(2))`, plus the fp32->bf16 operand cast. */
int cti_cast_rows_mask(const float* x, void* out_bf16, uint8_t* rowmask, int64_t rows, int cols, void* stream);
/* The same mask for features that arrive as bf16 already (the loader wire format of SURVEY 8f row 5: padded
 * (images, K, v_dim) bf16 -- half the host-to-device bytes of the reference's fp32 batches, no cast on the device):
 * rowmask[r] = 1 iff row r is all zeros.  replaces: `v.abs().sum(2) == 0`, src/attention.py:55 / :36. */
int cti_rowmask_bf16(const void* x_bf16, uint8_t* rowmask, int64_t rows, int cols, void* stream);

/* ---- training-mode dropout -------------------------------------------------------------------
 * The keep mask is a pure function of (seed, offset, element index): Philox4x32-10 keyed by `seed`, counter
 * (element / 4, offset); kept values are scaled by 1 / (1 - p).  Backward regenerates the mask from the same
 * (seed, offset) instead of storing it.
 * replaces: nn.Dropout on the FCNet input (src/fc.py:25-26) and BCNet's attention dropout (src/bc.py:53). */
/* out = bf16(dropout(x)); rowmask (may be NULL) is computed from the undropped rows; cols % 8 == 0. */
int cti_cast_rows_dropout(const float* x, void* out_bf16, uint8_t* rowmask, int64_t rows, int cols, float p,
                          uint64_t seed, uint64_t offset, void* stream);
/* x *= keep / (1 - p) in place, fp32: backward of an input dropout applied by cti_cast_rows_dropout (n % 4 == 0). */
int cti_dropout_f32(float* x, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* out = dropout(x) on bf16 data (n % 4 == 0); out may alias x. */
int cti_dropout_bf16(const void* x, void* out, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);

/* Per-rank input dropout of the R per-rank nets (src/tc.py:29-31,47-49): every rank r masks the shared input with
 * its own mask keep_r; element index of (r, row, c) is (r * rows + row) * cols + c.
 *   expand: xt[row, j*cols + c] = x[row, c] * keep_{r0+j}[row, c] / (1-p),  j < rank_group      (bf16, (rows, rank_group*cols))
 *   reduce: acc[row, c] += sum_j dxt[row, j*cols + c] * keep_{r0+j}[row, c] / (1-p)              (fp32 accumulator) */
int cti_dropout_expand(const void* x, void* xt, int64_t rows, int cols, int rank_group, int r0, float p, uint64_t seed,
                       uint64_t offset, void* stream);
int cti_dropout_reduce(const void* dxt, float* acc, int64_t rows, int cols, int rank_group, int r0, float p, uint64_t seed,
                       uint64_t offset, void* stream);

/* ---- weight-norm fold ----------------------------------------------------------------------
 * A matrix of n_groups stacked nn.Linear weights, each (rows_per_group, cols), each with its own
 * scalar g:  sumsq[i] = ||V_i||_F^2,  W_eff_i = bf16(V_i * g_i / ||V_i||_F).  The norm is reduced in a fixed order
 * (no atomics): equal weights give bit-equal packs.  `sumsq` (and `dot_ws` of the backward) must hold
 * cti_wn_scratch_floats(...) floats: the first n_groups are the result, the rest per-segment partial sums that every
 * consumer block re-reduces itself (two launches per call instead of three).
 * replaces: torch.nn.utils.weight_norm(nn.Linear, dim=None) as used by src/fc.py:22,27. */
size_t cti_wn_scratch_floats(int n_groups, int rows_per_group, int cols);
/* The fold of many layers in two launches (all weight packs of a model after an optimizer step).  One entry = one
 * group: v_ptrs / g_ptrs / w_ptrs / sumsq_ptrs are DEVICE arrays of n_entries device pointers (fp32 weight_v, scalar
 * weight_g, bf16 destination, fp32 destination of ||V||^2), elems its element count (multiple of 4).  Segment table:
 * segment s (4096 elements) = (seg_entry[s], seg_index[s]), the segments of an entry are consecutive and start at
 * first_seg[entry] (n_seg[entry] of them); block table: block b (1024 elements) = (blk_entry[b], blk_index[b]).
 * partials: n_segs floats of scratch.  Bit-identical to cti_wn_pack of each entry on its own. */
int cti_wn_pack_multi(const void* v_ptrs_dev, const void* g_ptrs_dev, const void* w_ptrs_dev, const void* sumsq_ptrs_dev,
                      const int64_t* elems_dev, const int32_t* first_seg_dev, const int32_t* n_seg_dev,
                      const int32_t* seg_entry_dev, const int32_t* seg_index_dev, int n_segs, const int32_t* blk_entry_dev,
                      const int32_t* blk_index_dev, int n_blks, float* partials, void* stream);
/* The backward of the same fold for every entry in two launches: dv[e] = (g/||V||) (dW - <dW,V> V / ||V||^2),
 * dg[e] = <dW,V> / ||V||, with ||V||^2 as left in sumsq[e] by cti_wn_pack_multi.  Tables as above plus dw / dv / dg
 * pointer tables; same arithmetic and summation order as cti_wn_grad (bit-identical results).
 * replaces: autograd through torch's weight_norm re-parametrisation (src/fc.py:22,27), all layers of a model at once. */
int cti_wn_grad_multi(const void* dw_ptrs_dev, const void* v_ptrs_dev, const void* g_ptrs_dev, const void* sumsq_ptrs_dev,
                      const void* dv_ptrs_dev, const void* dg_ptrs_dev, const int64_t* elems_dev, const int32_t* first_seg_dev,
                      const int32_t* n_seg_dev, const int32_t* seg_entry_dev, const int32_t* seg_index_dev, int n_segs,
                      const int32_t* blk_entry_dev, const int32_t* blk_index_dev, int n_blks, float* partials, void* stream);
int cti_wn_pack(const float* v, const float* g, void* w_eff_bf16, float* sumsq, int n_groups, int rows_per_group,
                int cols, void* stream);
/* Backward of the fold: given dW_eff (fp32) returns dV and dg.  dot_ws: cti_wn_scratch_floats() floats of scratch;
 * sumsq: the buffer cti_wn_pack filled. */
int cti_wn_grad(const float* dw_eff, const float* v, const float* g, const float* sumsq, float* dv, float* dg,
                float* dot_ws, int n_groups, int rows_per_group, int cols, void* stream);

/* ---- dense projection GEMM (tcgen05 / TMEM / TMA) --------------------------------------------
 * C[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T)
 *   a_mn_major = 0: A stored [M][lda] (K contiguous);  1: stored [K][lda] (M contiguous)
 *   b_mn_major = 0: B stored [N][ldb] (K contiguous);  1: stored [K][ldb] (N contiguous)
 *   epilogue: + bias[N] (if non-NULL); ReLU (if relu); * (relu_aux[M,ld_aux] > 0) (if non-NULL);
 *             store bf16 and/or fp32 with pitch ldc; or, with atomic_f32, out_f32 += result
 *             (k_splits > 1 splits the K loop over CTAs; needs atomic_f32).
 * replaces: nn.Linear inside FCNet.forward (src/fc.py:33-34), nn.ReLU (src/fc.py:29), and the
 *           autograd dgrad / wgrad of the same layer. */
int cti_gemm_bf16(const void* a, int lda, int a_mn_major, const void* b, int ldb, int b_mn_major, int M, int N, int K,
                  float alpha, const float* bias, int relu, const void* relu_aux, int ld_aux, void* out_bf16,
                  float* out_f32, int ldc, int atomic_f32, int k_splits, int tile_n, void* stream);
/* Two independent problems with the same operand layouts in ONE launch (fields as the arguments of cti_gemm_bf16):
 * the question-side and answer-side projections of a module call (q_net / a_net of src/tc.py:25-27,44-49 and their
 * autograd) have the same weight shapes and different row counts; alone each leaves SMs idle and pays the launch and
 * pipeline-fill floor.  The tiles of d0 are scheduled first. */
typedef struct cti_gemm_desc {
  const void* a; int lda; int a_mn_major;
  const void* b; int ldb; int b_mn_major;
  int M, N, K; float alpha;
  const float* bias; int relu;
  const void* relu_aux; int ld_aux;
  void* out_bf16; float* out_f32; int ldc; int atomic_f32; int k_splits; int tile_n;
} cti_gemm_desc;
int cti_gemm_bf16_pair(const cti_gemm_desc* d0, const cti_gemm_desc* d1, void* stream);

/* ---- activation backward + bias gradient ----------------------------------------------------
 * dz = dy * (y > 0) (y may be NULL: no activation), written as bf16 (dz may be NULL);
 * dbias_accum[n] += sum_m dz[m,n].  dy is fp32 (dy_is_bf16 = 0) or bf16.
 * replaces: autograd of nn.ReLU + the bias term of nn.Linear (src/fc.py:27-29). */
int cti_act_bwd_bias(const void* dy, int dy_is_bf16, const void* y_bf16, void* dz_bf16, float* dbias_accum,
                     int64_t rows, int cols, void* stream);

/* ---- distillation loss, forward + gradient in one pass -----------------------------------------
 * loss = mean_b KL(softmax(teacher/T) || softmax(x/T)) * alpha T^2 + sum BCEWithLogits(x, target) / B * (1 - alpha);
 * dx = d loss / d x.  x, target, dx: (B, N) fp32; teacher: (B, N) fp16 (teacher_is_fp16 != 0) or fp32; row_loss: B floats
 * of scratch; loss: one float.  2 launches.
 * replaces: Distillation_Loss.forward, src/loss_function.py:20-25 (nn.KLDivLoss + log_softmax + softmax +
 * nn.BCEWithLogitsLoss and their autograd). */
int cti_kd_loss(const float* x, const void* teacher, int teacher_is_fp16, const float* target, float* dx,
                float* row_loss, float* loss, int B, int N, float T, float alpha, void* stream);

/* ---- masked softmax over the flattened attention domain ---------------------------------------
 * rows of `len` contiguous floats; masked positions already hold -inf.
 * replaces: torch.softmax in src/attention.py:58 (TriAttention) and :39 (BiAttention). */
int cti_masked_softmax_fwd(const float* logits, float* p, int64_t rows, int len, void* stream);
/* dlogits = p * (dp - sum(p*dp)); p, dlogits are (batch, groups, len) contiguous; dp[b,g,e] is read at
 * dp + b*dp_stride_b + g*dp_stride_g + e*dp_stride_e (element strides). */
int cti_masked_softmax_bwd(const float* p, const float* dp, int64_t dp_stride_b, int64_t dp_stride_g,
                           int64_t dp_stride_e, float* dlogits, int64_t batch, int groups, int len, void* stream);

/* out[g,e] = sum_{j<rep} x[g*rep+j, e]: x (groups*rep, row_elems) bf16 -> out (groups, row_elems) bf16, fp32 sum;
 * row_elems % 8 == 0.  Folds per-row dzv back onto the shared image (the gradient of the expand() in
 * src/MC/train.py:75). */
int cti_sum_row_groups(const void* x_bf16, void* out_bf16, int64_t groups, int rep, int64_t row_elems, void* stream);

/* ---- GRU timestep, pointwise part (the products are cti_gemm_bf16 calls) --------------------------------
 * gx (row stride gx_row_stride floats) = x_t W_ih^T + b_ih and gh (rows, 3H) = h_{t-1} W_hh^T + b_hh, gate order r,z,n.
 * fwd: r = sigmoid(gx_r+gh_r), z = sigmoid(gx_z+gh_z), n = tanh(gx_n + r*gh_n), h = (1-z) n + z h_prev
 *      -> h_out fp32 (row stride ho_row_stride), h_bf16 / r / z / n / gh_n (rows, H) bf16.  h_prev NULL = zeros.
 * bwd: on entry dh (rows,H) = gradient carried from step t+1; dout = this step's output gradient (added inside);
 *      writes dgx (row stride dgx_row_stride) and dgh (rows, 3H) bf16 and leaves dh = dh*z (the direct part of the
 *      gradient w.r.t. h_{t-1}; the caller adds dgh W_hh with an accumulating GEMM).
 * replaces: the cuDNN GRU behind nn.GRU in QuestionEmbedding (src/language_model.py:56-61,93-98). */
int cti_gru_gate_fwd(const float* gx, int64_t gx_row_stride, const float* gh, const float* h_prev, int64_t hp_row_stride,
                     float* h_out, int64_t ho_row_stride, void* h_bf16, void* r_bf16, void* z_bf16, void* n_bf16,
                     void* ghn_bf16, int64_t rows, int H, void* stream);
int cti_gru_gate_bwd(float* dh, const float* dout, int64_t do_row_stride, const float* h_prev, int64_t hp_row_stride,
                     const void* r_bf16, const void* z_bf16, const void* n_bf16, const void* ghn_bf16, void* dgx_bf16,
                     int64_t dgx_row_stride, void* dgh_bf16, int64_t rows, int H, void* stream);

/* ---- trainer tail: gradient norm, rescale + clip + Adamax, multi-tensor ---------------------------------
 * Tables in DEVICE memory: g_ptrs / p_ptrs / m_ptrs / u_ptrs = arrays of n_tensors device pointers (fp32 tensors of
 * numel[t] elements: gradient, parameter, Adamax exp_avg and exp_inf); chunk c covers elements
 * [chunk_start[c], chunk_start[c] + chunk_elems) of tensor chunk_tensor[c] (clipped to numel).
 * cti_grad_sumsq_multi: sumsq[0] = sum over all tensors of g^2 (partials: n_chunks floats of scratch; fixed order).
 * cti_adamax_multi: norm = sqrt(sumsq) * inv_denom -> norm_out (may be NULL); g' = g * inv_denom * coef with
 *   coef = clip_norm / (norm + 1e-6) if norm > clip_norm > 0 else 1; m = b1 m + (1-b1) g'; u = max(b2 u, |g'| + eps);
 *   p -= clr * m / u  (clr = lr / (1 - b1^t)).
 * replaces: Trainer._all_reduce_and_rescale + clip_grad_norm_ + torch.optim.Adamax.step
 *           (src/MC/trainer.py:208-219,252-256; src/utils.py:323-328; src/MC/train.py:32). */
int cti_grad_sumsq_multi(const void* g_ptrs_dev, const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                         const int64_t* chunk_start_dev, int n_chunks, int chunk_elems, float* partials, float* sumsq,
                         void* stream);
int cti_adamax_multi(const void* p_ptrs_dev, const void* g_ptrs_dev, const void* m_ptrs_dev, const void* u_ptrs_dev,
                     const int64_t* numel_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_start_dev,
                     int n_chunks, int chunk_elems, const float* sumsq, float inv_denom, float clip_norm, float clr,
                     float beta1, float beta2, float eps, float* norm_out, void* stream);

/* ---- trilinear logit map ----------------------------------------------------------------------
 * vc (B,K,R*16), qc (B,Q,R*16), ac (B,A,R*16) bf16: the per-rank projections, column = r*16 + i.
 * tpack (R,16,16*G*16) bf16: T_eff[r][l][(i,g,j)] (see DESIGN.md for the T_g -> T_eff permutation).
 * tpack_perm (forward only; may be NULL): the same core with the last axis in the accumulator-lane order of the
 *   tcgen05 kernel, tpack_perm[r][l][(j%4)*128 + g*64 + i*4 + j/4] = tpack[r][l][i*32 + g*16 + j]  (G = 2).  With NULL,
 *   or for shapes outside the fast path (G != 2, A > 6, K > 64), the generic tensor-core kernel reads tpack.
 * logits (B,G,K,Q,A) fp32, -inf where rowmask[b*K+k] != 0 (rowmask may be NULL).
 * v_rep >= 1 (B % v_rep == 0): rows b*v_rep .. b*v_rep+v_rep-1 share ONE image -- vc and rowmask then hold B/v_rep
 *   samples and row b reads sample b / v_rep.  This is the MC x4 candidate clone of src/MC/train.py:75-76 without the
 *   copies; v_rep = 1 is the reference layout.  The backward still writes dzv per row b, (B,K,R*16): fold it with
 *   cti_sum_row_groups before the v-side wgrad.
 * replaces: the rank loop of TCNet.forward (src/tc.py:46-52) incl. Tensor.ModeProduct
 *           (src/Tensor.py:3-19) and the masked_fill_ of src/attention.py:55-56.
 * n1_save (may be NULL; cti_trilinear_n1_bytes() bytes, 0 = nothing to save for this shape): training.  The tcgen05
 *   kernel also writes its bf16 intermediate N1 = T x_l Ac (per row and rank quad) there; handing it to
 *   cti_trilinear_logits_bwd as n1_saved selects the fast backward, which does not recompute it. */
int cti_trilinear_logits_fwd(const void* vc, const void* qc, const void* ac, const void* tpack, const void* tpack_perm,
                             const uint8_t* rowmask, float* logits, void* n1_save, int B, int K, int Q, int A, int G,
                             int R, int v_rep, void* stream);
size_t cti_trilinear_n1_bytes(int B, int K, int Q, int A, int G, int R);
size_t cti_trilinear_logits_bwd_workspace(int B, int K, int Q, int A, int G, int R);
/* dz* are the PRE-activation gradients of the per-rank projections (ReLU masks applied);
 * db*_accum (R*16 each) and dtpack_accum (same shape as tpack, fp32) are accumulated into. */
int cti_trilinear_logits_bwd(const void* vc, const void* qc, const void* ac, const void* tpack, const float* dlogits,
                             const void* n1_saved, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum, float* dba_accum,
                             float* dtpack_accum, void* workspace, size_t workspace_bytes, int B, int K, int Q, int A,
                             int G, int R, int v_rep, void* stream);

/* ---- attention-weighted pooling ---------------------------------------------------------------
 * out[b,c] = sum_{k,q,a} V[b,k,c] w[b,k,q,a] Qp[b,q,c] Ap[b,a,c];  A = 0 drops the Ap factor.
 * v (B/v_rep,K,C), q (B,Q,C), a (B,A,C) bf16; w: (K,Q,A) contiguous per sample, samples w_stride_b floats apart.
 * v_rep as for cti_trilinear_logits_fwd (row b pools image b / v_rep); dzv of the backward is per row b, (B,K,C).
 * replaces: the einsum of TCNet.forward_with_weights (src/tc.py:59) and the two matmuls of
 *           BCNet.forward_with_weights (src/bc.py:73). */
int cti_tri_pool_fwd(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b, float* out,
                     int B, int K, int Q, int A, int C, int v_rep, void* stream);
/* dz* pre-activation gradients (ReLU masks of v, q, a applied); db*_accum (C each) accumulated into;
 * dw (B,K,Q,A) contiguous fp32. */
int cti_tri_pool_bwd(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b,
                     const float* dout, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum,
                     float* dba_accum, float* dw, int B, int K, int Q, int A, int C, int v_rep, void* stream);

/* Same, with the attention gradient written at a row stride: dw[b * dw_stride_b + (k*Q + q)*A + a].  The fused glimpse
 * loop hands in the glimpse-g slice of ONE (B, G, K*Q*A) buffer -- the layout cti_masked_softmax_bwd reads -- so the
 * backward of `att[:, :, :, :, g]` (reference src/MC/base_model.py:146: a zero fill, a strided copy and an add per
 * glimpse under autograd) costs nothing.  dw_stride_b = 0: contiguous. */
int cti_tri_pool_bwd_strided(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b,
                             const float* dout, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum,
                             float* dba_accum, float* dw, int64_t dw_stride_b, int B, int K, int Q, int A, int C, int v_rep,
                             void* stream);

/* ---- the R per-rank projections of a modality with PER-RANK input dropout (training mode) -----------------------------
 * replaces: `self.v_net[r](v_tucker)` for r < rank (src/tc.py:29-31,47-49), each FCNet = Dropout(p) -> weight_norm(Linear
 * (512, 16)) -> ReLU (src/fc.py:25-29), i.e. R independent Bernoulli masks on the same (M, 512) input.
 *   fwd  : out[m, r*16+j] = relu(s * sum_k keep_r[m,k] y[m,k] W[r*16+j, k] + bias[r*16+j])          (bf16)
 *   dgrad: dzt[m, k]      = (y[m,k] > 0) * s * sum_r keep_r[m,k] * sum_j dz[m, r*16+j] W[r*16+j, k]  (bf16; y's ReLU mask applied)
 *   wgrad: dw_accum[r*16+j, k] += s * sum_m dz[m, r*16+j] keep_r[m,k] y[m,k]                          (fp32, accumulates)
 * H == 512, R == 16 or 32.  One call takes 1-4 problems (the modalities of a TCNet); problems with the same mask kind
 * share a launch, so the question and answer sides fill one wave together.  The masks are never stored: each kernel
 * regenerates them in registers from (seed, site) with Philox4x32-7 -- one random bit per decision when p == 0.5, else 8 bits
 * per decision (drop rate round(256 p) / 256; s = cti_rank_proj_dropout_scale(p) = 256 / (256 - round(256 p))).
 * cti_rank_proj_dropout_mask writes keep[r, m, k] (uint8, R*M*H) for tests.  Each pass reads only the fields it needs. */
typedef struct cti_rank_proj_problem {
  const void* y;        /* (M, H) bf16: input of the per-rank nets (the post-ReLU tucker output) */
  const void* w_eff;    /* (R*16, H) bf16 weight-norm fold                       fwd, dgrad */
  const float* bias;    /* (R*16)                                               fwd */
  void* out;            /* (M, R*16) bf16                                       fwd */
  const void* dz;       /* (M, R*16) bf16 pre-activation gradient               dgrad, wgrad */
  void* dzt;            /* (M, H) bf16                                          dgrad */
  float* dw_accum;      /* (R*16, H) fp32                                       wgrad */
  int64_t M;
  float p;
  uint64_t seed, site;
} cti_rank_proj_problem;
float cti_rank_proj_dropout_scale(float p);
int cti_rank_proj_dropout_fwd(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream);
int cti_rank_proj_dropout_dgrad(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream);
int cti_rank_proj_dropout_wgrad(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream);
int cti_rank_proj_dropout_mask(uint8_t* keep, int64_t M, int H, int R, float p, uint64_t seed, uint64_t site, void* stream);

/* ---- caller glue of the glimpse loop (opt-in fused call) ------------------------------------------
 * replaces the torch ops of src/MC/base_model.py:147-150 / src/FFOE/base_model.py:127-130:
 *     q_emb = q_prj[g](b_emb[g].unsqueeze(1)) + q_emb ; ans_emb = a_prj[g](...) + ans_emb ; q_emb.sum(1) + ans_emb.sum(1)
 * x* (B, T*, D) fp32 or bf16 token tensors; res_* = HOST arrays of n_res (<= 4) device pointers to (B, D) fp32 projection
 * outputs, added in order ((x + r0) + r1) + ... exactly as the caller's loop does; D % 8 == 0; xa may be NULL.
 * residual_cast: out_*[b,t,:] = bf16(x*[b,t,:] + sum_i res_*[i][b,:]) -- the only form of the updated q_emb / ans_emb the next
 *     glimpse consumes (operand of its q_tucker / a_tucker GEMM).
 * token_sum:     out[b,:] = sum_t (xq[b,t,:] + ...) + sum_t (xa[b,t,:] + ...)  (fp32 and / or bf16 output).
 * bcast_rows:    out_*[b,t,:] = x[b,:]: the joint gradient as the initial value of d q_emb / d ans_emb, which the dgrad GEMMs
 *     of every glimpse then accumulate into. */
int cti_glimpse_residual_cast(const void* xq, int q_is_bf16, const float* const* res_q, int Tq, void* out_q_bf16,
                              const void* xa, int a_is_bf16, const float* const* res_a, int Ta, void* out_a_bf16,
                              int n_res, int64_t B, int D, void* stream);
int cti_glimpse_token_sum(const void* xq, int q_is_bf16, const float* const* res_q, int Tq, const void* xa, int a_is_bf16,
                          const float* const* res_a, int Ta, int n_res, float* out, void* out_bf16, int64_t B, int D,
                          void* stream);
int cti_glimpse_bcast_rows(const float* x, float* out_q, int Tq, float* out_a, int Ta, int64_t B, int D, void* stream);

/* ---- bilinear attention logits (BAN) ------------------------------------------------------------
 * logits[b,g,k,q] = sum_c Vb[b,k,c] h[g,c] Qb[b,q,c] + hbias[g]; -inf where rowmask[b*K+k] != 0.
 * vb (B,K,C), qb (B,Q,C) bf16; hmat (G,C), hbias (G) fp32; logits (B,G,K,Q) fp32.
 * replaces: BCNet.forward, h_out <= 32 branch (src/bc.py:52-58) and masked_fill_ of src/attention.py:36-37. */
int cti_bilinear_logits_fwd(const void* vb, const void* qb, const float* hmat, const float* hbias,
                            const uint8_t* rowmask, float* logits, int B, int K, int Q, int G, int C, void* stream);
int cti_bilinear_logits_bwd(const void* vb, const void* qb, const float* hmat, const float* dlogits, void* dzv,
                            void* dzq, float* dbv_accum, float* dbq_accum, float* dhmat_accum, float* dhbias_accum,
                            int B, int K, int Q, int G, int C, void* stream);

/* ---- gradient sum over NVLink peer memory (data-parallel training) -----------------------------------
 * The one collective of the path: the sum of the parameter gradients over the ranks of one node, which the reference's
 * Trainer._all_reduce_and_rescale names but never issues (src/MC/trainer.py:208-219).  Bytes move on the copy engines
 * (peer writes over NVLink / NVSwitch), not in an SM-resident transfer kernel, so the transfer overlaps backward without
 * displacing its persistent kernels; see csrc/peer.cu for the protocol and dp.PeerAllReducer for the host side.
 * Every rank allocates ONE region with cti_peer_alloc (zero-filled: flag block first, CTI_PEER_FLAG_BYTES), exports it
 * (64-byte CUDA IPC handle, exchanged by the host) and imports its peers' regions.
 * cti_peer_barrier: one-CTA kernel on `stream`; flag_blocks = HOST array of `world` device pointers (entry `rank` = the
 *   local region).  Orders everything enqueued before it on every rank's stream before everything after it on this rank's.
 *   A peer that does not arrive within timeout_s is recorded (cti_peer_error returns 1 + its rank), the kernel leaves.
 * cti_peer_barrier_memops: the same barrier as stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32): no
 *   resident kernel, so a rank waiting for a slower one holds no SM resources (a spinning CTA keeps one SM from taking a
 *   CTA of the path's one-CTA-per-SM kernels).  Consecutive barriers must use different slots (0..7); no timeout.
 * cti_peer_flag_op: one stream memory operation on word `index` (< 1024; words 512.. are free for the host protocol) of a
 *   flag block, local or a peer's: wait = 0 writes `value`, wait = 1 waits until the word is >= `value`.
 * cti_peer_flag_ops: up to 16 such operations on one flag block as ONE batch (one graph node), executed in order.
 * cti_peer_allreduce_fused: the whole exchange of one range as ONE kernel (one CTA per SM; 16-byte stores over NVLink into
 *   the peers' staging, flag round, owner-side sum in rank order stored into every rank's slab, flag round): for the last
 *   bucket of a step, where nothing is left to overlap with and latency is what counts.  slab_ranges[r] = start of the
 *   range in rank r's slab, stagings[r] = rank r's staging buffer for this call (>= n + 4 * world floats, not shared with
 *   the copy-engine exchanges); n % 4 == 0.  Issue it where no other kernel competes for the SMs: every CTA spins.
 * cti_peer_copy: cudaMemcpyAsync between local and peer-mapped memory (copy engine).
 * cti_sum_staged: dst[i] = sum over ranks, in rank order, of the local copy (dst, at position `rank`) and n_staged staged
 *   copies (staged + s * stride floats); n % 4 == 0, 16-byte aligned. */
#define CTI_PEER_FLAG_BYTES 4096
int cti_peer_alloc(size_t bytes, void** ptr);
int cti_peer_free(void* ptr);
int cti_peer_export(void* ptr, void* handle64);
int cti_peer_import(const void* handle64, void** ptr);
int cti_peer_close(void* ptr);
int cti_peer_barrier(void* const* flag_blocks, int rank, int world, int slot, double timeout_s, void* stream);
int cti_peer_barrier_memops(void* const* flag_blocks, int rank, int world, int slot, void* stream);
int cti_peer_flag_ops(void* flag_block, const int* index, const uint32_t* value, const int* wait, int count, void* stream);
int cti_peer_flag_op(void* flag_block, int index, uint32_t value, int wait, void* stream);
int cti_peer_error(const void* flag_block, int* out);
/* debug: *dst = %globaltimer (ns) when the stream reaches this point (timeline of the overlapped transfers) */
int cti_peer_stamp(uint64_t* dst, void* stream);
int cti_peer_allreduce_fused(void* const* flag_blocks, void* const* slab_ranges, void* const* stagings, int rank,
                             int world, int64_t n, double timeout_s, void* stream);
int cti_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
int cti_sum_staged(float* dst, const float* staged, int n_staged, int rank, int64_t n, int64_t stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTI_SM100_H_ */
