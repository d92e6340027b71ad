// extern "C" boundary of libcti_sm100.so (declared in include/cti_sm100.h).
// Thin argument marshalling + thread-local error state; all work is in the kernel files.
#include "../../include/cti_sm100.h"

#include "cti_common.cuh"
#include "cti_kernels.h"

#include <cstdarg>
#include <cstdio>

namespace cti {

namespace {
thread_local char g_err[512] = {0};
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace cti

extern "C" {

int cti_version(void) { return 102; }   // 0.1.1: v_rep (rows sharing one v sample), cti_sum_row_groups

const char* cti_last_error(void) { return cti::g_err; }

int cti_cast_rows_mask(const float* x, void* out_bf16, uint8_t* rowmask, int64_t rows, int cols, void* stream) {
  return cti::cast_rows_mask(x, static_cast<__nv_bfloat16*>(out_bf16), rowmask, rows, cols,
                             static_cast<cudaStream_t>(stream));
}

int cti_cast_rows_dropout(const float* x, void* out_bf16, uint8_t* rowmask, int64_t rows, int cols, float p,
                          uint64_t seed, uint64_t offset, void* stream) {
  return cti::cast_rows_dropout(x, static_cast<__nv_bfloat16*>(out_bf16), rowmask, rows, cols, p, seed, offset,
                                static_cast<cudaStream_t>(stream));
}

int cti_dropout_f32(float* x, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  return cti::dropout_f32(x, n, p, seed, offset, static_cast<cudaStream_t>(stream));
}

int cti_dropout_bf16(const void* x, void* out, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  return cti::dropout_bf16(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), n, p, seed, offset,
                           static_cast<cudaStream_t>(stream));
}

int cti_dropout_expand(const void* x, void* xt, int64_t rows, int cols, int rank_group, int r0, float p, uint64_t seed,
                       uint64_t offset, void* stream) {
  return cti::dropout_expand(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(xt), rows, cols, rank_group,
                             r0, p, seed, offset, static_cast<cudaStream_t>(stream));
}

int cti_dropout_reduce(const void* dxt, float* acc, int64_t rows, int cols, int rank_group, int r0, float p, uint64_t seed,
                       uint64_t offset, void* stream) {
  return cti::dropout_reduce(static_cast<const __nv_bfloat16*>(dxt), acc, rows, cols, rank_group, r0, p, seed, offset,
                             static_cast<cudaStream_t>(stream));
}

int cti_wn_pack_multi(const void* v_ptrs_dev, const void* g_ptrs_dev, const void* w_ptrs_dev, const void* sumsq_ptrs_dev,
                      const int64_t* elems_dev, const int32_t* first_seg_dev, const int32_t* n_seg_dev,
                      const int32_t* seg_entry_dev, const int32_t* seg_index_dev, int n_segs, const int32_t* blk_entry_dev,
                      const int32_t* blk_index_dev, int n_blks, float* partials, void* stream) {
  return cti::wn_pack_multi(v_ptrs_dev, g_ptrs_dev, w_ptrs_dev, sumsq_ptrs_dev, reinterpret_cast<const long*>(elems_dev),
                            first_seg_dev, n_seg_dev, seg_entry_dev, seg_index_dev, n_segs, blk_entry_dev, blk_index_dev,
                            n_blks, partials, static_cast<cudaStream_t>(stream));
}

int cti_rowmask_bf16(const void* x_bf16, uint8_t* rowmask, int64_t rows, int cols, void* stream) {
  return cti::rowmask_bf16(static_cast<const __nv_bfloat16*>(x_bf16), rowmask, rows, cols, static_cast<cudaStream_t>(stream));
}

int cti_wn_grad_multi(const void* dw_ptrs_dev, const void* v_ptrs_dev, const void* g_ptrs_dev, const void* sumsq_ptrs_dev,
                      const void* dv_ptrs_dev, const void* dg_ptrs_dev, const int64_t* elems_dev, const int32_t* first_seg_dev,
                      const int32_t* n_seg_dev, const int32_t* seg_entry_dev, const int32_t* seg_index_dev, int n_segs,
                      const int32_t* blk_entry_dev, const int32_t* blk_index_dev, int n_blks, float* partials, void* stream) {
  return cti::wn_grad_multi(dw_ptrs_dev, v_ptrs_dev, g_ptrs_dev, sumsq_ptrs_dev, dv_ptrs_dev, dg_ptrs_dev,
                            reinterpret_cast<const long*>(elems_dev), first_seg_dev, n_seg_dev, seg_entry_dev, seg_index_dev,
                            n_segs, blk_entry_dev, blk_index_dev, n_blks, partials, static_cast<cudaStream_t>(stream));
}

size_t cti_wn_scratch_floats(int n_groups, int rows_per_group, int cols) {
  return cti::wn_scratch_floats(n_groups, rows_per_group, cols);
}

int cti_wn_pack(const float* v, const float* g, void* w_eff_bf16, float* sumsq, int n_groups, int rows_per_group,
                int cols, void* stream) {
  return cti::wn_pack(v, g, static_cast<__nv_bfloat16*>(w_eff_bf16), sumsq, n_groups, rows_per_group, cols,
                      static_cast<cudaStream_t>(stream));
}

int cti_wn_grad(const float* dw_eff, const float* v, const float* g, const float* sumsq, float* dv, float* dg,
                float* dot_ws, int n_groups, int rows_per_group, int cols, void* stream) {
  return cti::wn_grad(dw_eff, v, g, sumsq, dv, dg, dot_ws, n_groups, rows_per_group, cols,
                      static_cast<cudaStream_t>(stream));
}

int cti_gemm_bf16(const void* a, int lda, int a_mn_major, const void* b, int ldb, int b_mn_major, int M, int N, int K,
                  float alpha, const float* bias, int relu, const void* relu_aux, int ld_aux, void* out_bf16,
                  float* out_f32, int ldc, int atomic_f32, int k_splits, int tile_n, void* stream) {
  cti::GemmArgs g;
  g.a = static_cast<const __nv_bfloat16*>(a);
  g.b = static_cast<const __nv_bfloat16*>(b);
  g.M = M; g.N = N; g.K = K;
  g.lda = lda; g.ldb = ldb;
  g.a_mn_major = a_mn_major != 0;
  g.b_mn_major = b_mn_major != 0;
  g.bias = bias;
  g.relu_aux = static_cast<const __nv_bfloat16*>(relu_aux);
  g.ld_aux = ld_aux;
  g.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  g.out_f32 = out_f32;
  g.ldc = ldc;
  g.relu = relu;
  g.atomic_f32 = atomic_f32;
  g.k_splits = k_splits;
  g.alpha = alpha;
  g.tile_n = tile_n;
  return cti::gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

static cti::GemmArgs gemm_args_of(const cti_gemm_desc& d) {
  cti::GemmArgs g;
  g.a = static_cast<const __nv_bfloat16*>(d.a);
  g.b = static_cast<const __nv_bfloat16*>(d.b);
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.lda = d.lda; g.ldb = d.ldb;
  g.a_mn_major = d.a_mn_major != 0;
  g.b_mn_major = d.b_mn_major != 0;
  g.bias = d.bias;
  g.relu_aux = static_cast<const __nv_bfloat16*>(d.relu_aux);
  g.ld_aux = d.ld_aux;
  g.out_bf16 = static_cast<__nv_bfloat16*>(d.out_bf16);
  g.out_f32 = d.out_f32;
  g.ldc = d.ldc;
  g.relu = d.relu;
  g.atomic_f32 = d.atomic_f32;
  g.k_splits = d.k_splits;
  g.alpha = d.alpha;
  g.tile_n = d.tile_n;
  return g;
}

int cti_gemm_bf16_pair(const cti_gemm_desc* d0, const cti_gemm_desc* d1, void* stream) {
  if (d0 == nullptr || d1 == nullptr) {
    cti::set_error("cti_gemm_bf16_pair: null descriptor");
    return -1;
  }
  return cti::gemm_bf16_pair(gemm_args_of(*d0), gemm_args_of(*d1), static_cast<cudaStream_t>(stream));
}

int cti_act_bwd_bias(const void* dy, int dy_is_bf16, const void* y_bf16, void* dz_bf16, float* dbias_accum,
                     int64_t rows, int cols, void* stream) {
  return cti::act_bwd_bias(dy, dy_is_bf16, static_cast<const __nv_bfloat16*>(y_bf16),
                           static_cast<__nv_bfloat16*>(dz_bf16), dbias_accum, rows, cols,
                           static_cast<cudaStream_t>(stream));
}

int cti_sum_row_groups(const void* x_bf16, void* out_bf16, int64_t groups, int rep, int64_t row_elems, void* stream) {
  return cti::sum_row_groups(static_cast<const __nv_bfloat16*>(x_bf16), static_cast<__nv_bfloat16*>(out_bf16), groups, rep,
                             row_elems, static_cast<cudaStream_t>(stream));
}

int cti_grad_sumsq_multi(const void* g_ptrs_dev, const int64_t* numel_dev, const int32_t* chunk_tensor_dev,
                         const int64_t* chunk_start_dev, int n_chunks, int chunk_elems, float* partials, float* sumsq,
                         void* stream) {
  return cti::grad_sumsq_multi(static_cast<const float* const*>(g_ptrs_dev), reinterpret_cast<const long*>(numel_dev),
                               chunk_tensor_dev, reinterpret_cast<const long*>(chunk_start_dev), n_chunks, chunk_elems,
                               partials, sumsq, static_cast<cudaStream_t>(stream));
}

int cti_adamax_multi(const void* p_ptrs_dev, const void* g_ptrs_dev, const void* m_ptrs_dev, const void* u_ptrs_dev,
                     const int64_t* numel_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_start_dev,
                     int n_chunks, int chunk_elems, const float* sumsq, float inv_denom, float clip_norm, float clr,
                     float beta1, float beta2, float eps, float* norm_out, void* stream) {
  return cti::adamax_multi(static_cast<float* const*>(const_cast<void*>(p_ptrs_dev)),
                           static_cast<const float* const*>(g_ptrs_dev),
                           static_cast<float* const*>(const_cast<void*>(m_ptrs_dev)),
                           static_cast<float* const*>(const_cast<void*>(u_ptrs_dev)),
                           reinterpret_cast<const long*>(numel_dev), chunk_tensor_dev,
                           reinterpret_cast<const long*>(chunk_start_dev), n_chunks, chunk_elems, sumsq, inv_denom,
                           clip_norm, clr, beta1, beta2, eps, norm_out, static_cast<cudaStream_t>(stream));
}

int cti_gru_gate_fwd(const float* gx, int64_t gx_row_stride, const float* gh, const float* h_prev, int64_t hp_row_stride,
                     float* h_out, int64_t ho_row_stride, void* h_bf16, void* r_bf16, void* z_bf16, void* n_bf16,
                     void* ghn_bf16, int64_t rows, int H, void* stream) {
  using b16 = __nv_bfloat16;
  return cti::gru_gate_fwd(gx, gx_row_stride, gh, h_prev, hp_row_stride, h_out, ho_row_stride, static_cast<b16*>(h_bf16),
                           static_cast<b16*>(r_bf16), static_cast<b16*>(z_bf16), static_cast<b16*>(n_bf16),
                           static_cast<b16*>(ghn_bf16), rows, H, static_cast<cudaStream_t>(stream));
}

int cti_gru_gate_bwd(float* dh, const float* dout, int64_t do_row_stride, const float* h_prev, int64_t hp_row_stride,
                     const void* r_bf16, const void* z_bf16, const void* n_bf16, const void* ghn_bf16, void* dgx_bf16,
                     int64_t dgx_row_stride, void* dgh_bf16, int64_t rows, int H, void* stream) {
  using b16 = __nv_bfloat16;
  return cti::gru_gate_bwd(dh, dout, do_row_stride, h_prev, hp_row_stride, static_cast<const b16*>(r_bf16),
                           static_cast<const b16*>(z_bf16), static_cast<const b16*>(n_bf16),
                           static_cast<const b16*>(ghn_bf16), static_cast<b16*>(dgx_bf16), dgx_row_stride,
                           static_cast<b16*>(dgh_bf16), rows, H, static_cast<cudaStream_t>(stream));
}

int cti_kd_loss(const float* x, const void* teacher, int teacher_is_fp16, const float* target, float* dx, float* row_loss,
                float* loss, int B, int N, float T, float alpha, void* stream) {
  return cti::kd_loss(x, teacher, teacher_is_fp16, target, dx, row_loss, loss, B, N, T, alpha,
                      static_cast<cudaStream_t>(stream));
}

int cti_masked_softmax_fwd(const float* logits, float* p, int64_t rows, int len, void* stream) {
  return cti::masked_softmax_fwd(logits, p, rows, len, static_cast<cudaStream_t>(stream));
}

int cti_masked_softmax_bwd(const float* p, const float* dp, int64_t dp_stride_b, int64_t dp_stride_g,
                           int64_t dp_stride_e, float* dlogits, int64_t batch, int groups, int len, void* stream) {
  return cti::masked_softmax_bwd(p, dp, dp_stride_b, dp_stride_g, dp_stride_e, dlogits, batch, groups, len,
                                 static_cast<cudaStream_t>(stream));
}

int cti_trilinear_logits_fwd(const void* vc, const void* qc, const void* ac, const void* tpack, const void* tpack_perm,
                             const uint8_t* rowmask, float* logits, void* n1_save, int B, int K, int Q, int A, int G, int R,
                             int v_rep, void* stream) {
  cti::TriDims d{B, K, Q, A, G, R, v_rep};
  return cti::trilinear_fwd(static_cast<const __nv_bfloat16*>(vc), static_cast<const __nv_bfloat16*>(qc),
                            static_cast<const __nv_bfloat16*>(ac), static_cast<const __nv_bfloat16*>(tpack),
                            static_cast<const __nv_bfloat16*>(tpack_perm), rowmask, logits, n1_save, d,
                            static_cast<cudaStream_t>(stream));
}

size_t cti_trilinear_n1_bytes(int B, int K, int Q, int A, int G, int R) {
  cti::TriDims d{B, K, Q, A, G, R};
  return cti::trilinear_n1_bytes(d);
}

int cti_debug_prof_read(unsigned long long* host_dst, int n) { return cti::debug_prof_read(host_dst, n); }
int cti_debug_prof_read_bwd1(unsigned long long* host_dst, int n) { return cti::debug_prof_read_bwd1(host_dst, n); }

size_t cti_trilinear_logits_bwd_workspace(int B, int K, int Q, int A, int G, int R) {
  cti::TriDims d{B, K, Q, A, G, R};
  return cti::trilinear_bwd_workspace(d);
}

int cti_trilinear_logits_bwd(const void* vc, const void* qc, const void* ac, const void* tpack, const float* dlogits,
                             const void* n1_saved, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum, float* dba_accum,
                             float* dtpack_accum, void* workspace, size_t workspace_bytes, int B, int K, int Q, int A,
                             int G, int R, int v_rep, void* stream) {
  cti::TriDims d{B, K, Q, A, G, R, v_rep};
  return cti::trilinear_bwd(static_cast<const __nv_bfloat16*>(vc), static_cast<const __nv_bfloat16*>(qc),
                            static_cast<const __nv_bfloat16*>(ac), static_cast<const __nv_bfloat16*>(tpack), dlogits, n1_saved,
                            static_cast<__nv_bfloat16*>(dzv), static_cast<__nv_bfloat16*>(dzq),
                            static_cast<__nv_bfloat16*>(dza), dbv_accum, dbq_accum, dba_accum, dtpack_accum, workspace,
                            workspace_bytes, d, static_cast<cudaStream_t>(stream));
}

int cti_tri_pool_fwd(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b, float* out,
                     int B, int K, int Q, int A, int C, int v_rep, void* stream) {
  cti::PoolDims d{B, K, Q, A, C, v_rep};
  return cti::tri_pool_fwd(static_cast<const __nv_bfloat16*>(v), static_cast<const __nv_bfloat16*>(q),
                           static_cast<const __nv_bfloat16*>(a), w, w_stride_b, out, d,
                           static_cast<cudaStream_t>(stream));
}

int cti_tri_pool_bwd(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b,
                     const float* dout, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum,
                     float* dba_accum, float* dw, int B, int K, int Q, int A, int C, int v_rep, void* stream) {
  cti::PoolDims d{B, K, Q, A, C, v_rep};
  return cti::tri_pool_bwd(static_cast<const __nv_bfloat16*>(v), static_cast<const __nv_bfloat16*>(q),
                           static_cast<const __nv_bfloat16*>(a), w, w_stride_b, dout,
                           static_cast<__nv_bfloat16*>(dzv), static_cast<__nv_bfloat16*>(dzq),
                           static_cast<__nv_bfloat16*>(dza), dbv_accum, dbq_accum, dba_accum, dw, 0, d,
                           static_cast<cudaStream_t>(stream));
}

int cti_tri_pool_bwd_strided(const void* v, const void* q, const void* a, const float* w, int64_t w_stride_b,
                             const float* dout, void* dzv, void* dzq, void* dza, float* dbv_accum, float* dbq_accum,
                             float* dba_accum, float* dw, int64_t dw_stride_b, int B, int K, int Q, int A, int C, int v_rep,
                             void* stream) {
  cti::PoolDims d{B, K, Q, A, C, v_rep};
  return cti::tri_pool_bwd(static_cast<const __nv_bfloat16*>(v), static_cast<const __nv_bfloat16*>(q),
                           static_cast<const __nv_bfloat16*>(a), w, w_stride_b, dout,
                           static_cast<__nv_bfloat16*>(dzv), static_cast<__nv_bfloat16*>(dzq),
                           static_cast<__nv_bfloat16*>(dza), dbv_accum, dbq_accum, dba_accum, dw, (long)dw_stride_b, d,
                           static_cast<cudaStream_t>(stream));
}

float cti_rank_proj_dropout_scale(float p) { return cti::rank_proj_scale(p); }

static int rank_proj_convert(const cti_rank_proj_problem* probs, int n, cti::RankProjProblem (&out)[4]) {
  if (probs == nullptr || n < 1 || n > 4) {
    cti::set_error("cti_rank_proj_dropout: 1 to 4 problems per call (got %d)", n);
    return -1;
  }
  for (int i = 0; i < n; ++i) {
    const cti_rank_proj_problem& q = probs[i];
    cti::RankProjProblem& d = out[i];
    d.y = static_cast<const __nv_bfloat16*>(q.y);
    d.w_eff = static_cast<const __nv_bfloat16*>(q.w_eff);
    d.bias = q.bias;
    d.out = static_cast<__nv_bfloat16*>(q.out);
    d.dz = static_cast<const __nv_bfloat16*>(q.dz);
    d.dzt = static_cast<__nv_bfloat16*>(q.dzt);
    d.dw_accum = q.dw_accum;
    d.M = (long)q.M;
    d.p = q.p;
    d.seed = q.seed;
    d.site = q.site;
  }
  return 0;
}

int cti_rank_proj_dropout_fwd(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream) {
  cti::RankProjProblem ps[4];
  if (int rc = rank_proj_convert(probs, n, ps)) return rc;
  return cti::rank_proj_dropout_fwd(ps, n, H, R, static_cast<cudaStream_t>(stream));
}

int cti_rank_proj_dropout_dgrad(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream) {
  cti::RankProjProblem ps[4];
  if (int rc = rank_proj_convert(probs, n, ps)) return rc;
  return cti::rank_proj_dropout_dgrad(ps, n, H, R, static_cast<cudaStream_t>(stream));
}

int cti_rank_proj_dropout_wgrad(const cti_rank_proj_problem* probs, int n, int H, int R, void* stream) {
  cti::RankProjProblem ps[4];
  if (int rc = rank_proj_convert(probs, n, ps)) return rc;
  return cti::rank_proj_dropout_wgrad(ps, n, H, R, static_cast<cudaStream_t>(stream));
}

int cti_rank_proj_dropout_mask(uint8_t* keep, int64_t M, int H, int R, float p, uint64_t seed, uint64_t site, void* stream) {
  return cti::rank_proj_dropout_mask(keep, (long)M, H, R, p, seed, site, static_cast<cudaStream_t>(stream));
}

int cti_glimpse_residual_cast(const void* xq, int q_is_bf16, const float* const* res_q, int Tq, void* out_q_bf16,
                              const void* xa, int a_is_bf16, const float* const* res_a, int Ta, void* out_a_bf16,
                              int n_res, int64_t B, int D, void* stream) {
  return cti::glimpse_residual_cast(xq, q_is_bf16, res_q, Tq, static_cast<__nv_bfloat16*>(out_q_bf16), xa, a_is_bf16, res_a,
                                    Ta, static_cast<__nv_bfloat16*>(out_a_bf16), n_res, (long)B, D,
                                    static_cast<cudaStream_t>(stream));
}

int cti_glimpse_token_sum(const void* xq, int q_is_bf16, const float* const* res_q, int Tq, const void* xa, int a_is_bf16,
                          const float* const* res_a, int Ta, int n_res, float* out, void* out_bf16, int64_t B, int D,
                          void* stream) {
  return cti::glimpse_token_sum(xq, q_is_bf16, res_q, Tq, xa, a_is_bf16, res_a, Ta, n_res, out,
                                static_cast<__nv_bfloat16*>(out_bf16), (long)B, D, static_cast<cudaStream_t>(stream));
}

int cti_glimpse_bcast_rows(const float* x, float* out_q, int Tq, float* out_a, int Ta, int64_t B, int D, void* stream) {
  return cti::glimpse_bcast_rows(x, out_q, Tq, out_a, Ta, (long)B, D, static_cast<cudaStream_t>(stream));
}

int cti_bilinear_logits_fwd(const void* vb, const void* qb, const float* hmat, const float* hbias,
                            const uint8_t* rowmask, float* logits, int B, int K, int Q, int G, int C, void* stream) {
  cti::BiDims d{B, K, Q, G, C};
  return cti::bilinear_fwd(static_cast<const __nv_bfloat16*>(vb), static_cast<const __nv_bfloat16*>(qb), hmat, hbias,
                           rowmask, logits, d, static_cast<cudaStream_t>(stream));
}

int cti_bilinear_logits_bwd(const void* vb, const void* qb, const float* hmat, const float* dlogits, void* dzv,
                            void* dzq, float* dbv_accum, float* dbq_accum, float* dhmat_accum, float* dhbias_accum,
                            int B, int K, int Q, int G, int C, void* stream) {
  cti::BiDims d{B, K, Q, G, C};
  return cti::bilinear_bwd(static_cast<const __nv_bfloat16*>(vb), static_cast<const __nv_bfloat16*>(qb), hmat, dlogits,
                           static_cast<__nv_bfloat16*>(dzv), static_cast<__nv_bfloat16*>(dzq), dbv_accum, dbq_accum,
                           dhmat_accum, dhbias_accum, d, static_cast<cudaStream_t>(stream));
}

int cti_peer_alloc(size_t bytes, void** ptr) { return cti::peer_alloc(bytes, ptr); }
int cti_peer_free(void* ptr) { return cti::peer_free(ptr); }
int cti_peer_export(void* ptr, void* handle64) { return cti::peer_export(ptr, handle64); }
int cti_peer_import(const void* handle64, void** ptr) { return cti::peer_import(handle64, ptr); }
int cti_peer_close(void* ptr) { return cti::peer_close(ptr); }
int cti_peer_barrier(void* const* flag_blocks, int rank, int world, int slot, double timeout_s, void* stream) {
  return cti::peer_barrier(flag_blocks, rank, world, slot, timeout_s, static_cast<cudaStream_t>(stream));
}
int cti_peer_barrier_memops(void* const* flag_blocks, int rank, int world, int slot, void* stream) {
  return cti::peer_barrier_memops(flag_blocks, rank, world, slot, static_cast<cudaStream_t>(stream));
}
int cti_peer_flag_ops(void* flag_block, const int* index, const uint32_t* value, const int* wait, int count, void* stream) {
  return cti::peer_flag_ops(flag_block, index, value, wait, count, static_cast<cudaStream_t>(stream));
}
int cti_peer_flag_op(void* flag_block, int index, uint32_t value, int wait, void* stream) {
  return cti::peer_flag_op(flag_block, index, value, wait, static_cast<cudaStream_t>(stream));
}
int cti_peer_stamp(uint64_t* dst, void* stream) {
  return cti::peer_stamp(reinterpret_cast<unsigned long long*>(dst), static_cast<cudaStream_t>(stream));
}
int cti_peer_error(const void* flag_block, int* out) { return cti::peer_error(flag_block, out); }
int cti_peer_allreduce_fused(void* const* flag_blocks, void* const* slab_ranges, void* const* stagings, int rank,
                             int world, int64_t n, double timeout_s, void* stream) {
  return cti::peer_allreduce_fused(flag_blocks, slab_ranges, stagings, rank, world, n, timeout_s,
                                   static_cast<cudaStream_t>(stream));
}
int cti_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  return cti::peer_copy(dst, src, bytes, static_cast<cudaStream_t>(stream));
}
int cti_sum_staged(float* dst, const float* staged, int n_staged, int rank, int64_t n, int64_t stride, void* stream) {
  return cti::sum_staged(dst, staged, n_staged, rank, n, stride, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
