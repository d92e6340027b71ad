// Internal C++ declarations shared by the kernel translation units and the C-ABI
// wrapper (cti_capi.cu).  Not part of the public interface: see include/cti_sm100.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cti {

struct GemmArgs {
  const __nv_bfloat16* a = nullptr;   // A: logical (M x K). K-major: [M][lda]; MN-major: [K][lda]
  const __nv_bfloat16* b = nullptr;   // B: logical (N x K). K-major: [N][ldb]; MN-major: [K][ldb]
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0;
  bool a_mn_major = false, b_mn_major = false;
  const float* bias = nullptr;
  const __nv_bfloat16* relu_aux = nullptr;
  int ld_aux = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  float* out_f32 = nullptr;
  int ldc = 0;
  int relu = 0;
  int atomic_f32 = 0;
  int k_splits = 1;
  float alpha = 1.f;
  int tile_n = 0;      // 0 = heuristic, 128 or 256 = forced
  int max_ctas = 0;    // 0 = one per SM
};
int gemm_bf16(const GemmArgs& g, cudaStream_t stream);
int gemm_bf16_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t stream);

// elementwise.cu
int cast_rows_mask(const float* x, __nv_bfloat16* out, uint8_t* rowmask, long rows, int cols, cudaStream_t s);
int cast_rows_dropout(const float* x, __nv_bfloat16* out, uint8_t* rowmask, long rows, int cols, float p, uint64_t seed,
                      uint64_t offset, cudaStream_t s);
int dropout_f32(float* x, long n, float p, uint64_t seed, uint64_t offset, cudaStream_t s);
int dropout_bf16(const __nv_bfloat16* x, __nv_bfloat16* out, long n, float p, uint64_t seed, uint64_t offset,
                 cudaStream_t s);
int dropout_expand(const __nv_bfloat16* x, __nv_bfloat16* xt, long rows, int cols, int RG, int r0, float p, uint64_t seed,
                   uint64_t offset, cudaStream_t s);
int dropout_reduce(const __nv_bfloat16* dxt, float* acc, long rows, int cols, int RG, int r0, float p, uint64_t seed,
                   uint64_t offset, cudaStream_t s);
int sum_row_groups(const __nv_bfloat16* x, __nv_bfloat16* out, long groups, int rep, long row_elems, cudaStream_t s);
int rowmask_bf16(const __nv_bfloat16* x, uint8_t* rowmask, long rows, int cols, cudaStream_t s);
int wn_grad_multi(const void* dw_ptrs, const void* v_ptrs, const void* g_ptrs, const void* sumsq_ptrs, const void* dv_ptrs,
                  const void* dg_ptrs, const long* elems, const int* first_seg, const int* n_seg, const int* seg_entry,
                  const int* seg_index, int n_segs, const int* blk_entry, const int* blk_index, int n_blks, float* partials,
                  cudaStream_t s);
int wn_pack_multi(const void* v_ptrs, const void* g_ptrs, const void* w_ptrs, const void* sumsq_ptrs, const long* elems,
                  const int* first_seg, const int* n_seg, const int* seg_entry, const int* seg_index, int n_segs,
                  const int* blk_entry, const int* blk_index, int n_blks, float* partials, cudaStream_t s);
size_t wn_scratch_floats(int n_groups, int rows_per_group, int cols);
int wn_pack(const float* v, const float* g, __nv_bfloat16* w, float* sumsq, int n_groups, int rows_per_group, int cols,
            cudaStream_t s);
int wn_grad(const float* dw, const float* v, const float* g, const float* sumsq, float* dv, float* dg, float* dot_ws,
            int n_groups, int rows_per_group, int cols, cudaStream_t s);
int act_bwd_bias(const void* dy, int dy_is_bf16, const __nv_bfloat16* y, __nv_bfloat16* dz, float* dbias, long rows,
                 int cols, cudaStream_t s);

// glimpse.cu  (caller glue of the glimpse loop for the fused call: residual add + cast, token sums, gradient broadcast)
int glimpse_residual_cast(const void* xq, int q_bf16, const float* const* pq, int Tq, __nv_bfloat16* oq, const void* xa,
                          int a_bf16, const float* const* pa, int Ta, __nv_bfloat16* oa, int n_res, long B, int D,
                          cudaStream_t s);
int glimpse_token_sum(const void* xq, int q_bf16, const float* const* pq, int Tq, const void* xa, int a_bf16,
                      const float* const* pa, int Ta, int n_res, float* out, __nv_bfloat16* out_bf16, long B, int D,
                      cudaStream_t s);
int glimpse_bcast_rows(const float* x, float* oq, int Tq, float* oa, int Ta, long B, int D, cudaStream_t s);

// rank_proj.cu  (the R per-rank projections of a modality with per-rank input dropout, masks applied in registers)
struct RankProjProblem {
  const __nv_bfloat16* y = nullptr;        // (M, H) input of the per-rank nets (post-ReLU tucker output)
  const __nv_bfloat16* w_eff = nullptr;    // (R*16, H) weight-norm fold             (fwd, dgrad)
  const float* bias = nullptr;             // (R*16)                                 (fwd)
  __nv_bfloat16* out = nullptr;            // (M, R*16)                              (fwd)
  const __nv_bfloat16* dz = nullptr;       // (M, R*16) pre-activation gradient      (dgrad, wgrad)
  __nv_bfloat16* dzt = nullptr;            // (M, H)                                 (dgrad)
  float* dw_accum = nullptr;               // (R*16, H)                              (wgrad)
  long M = 0;
  float p = 0.f;
  uint64_t seed = 0, site = 0;
};
float rank_proj_scale(float p);
int rank_proj_dropout_fwd(const RankProjProblem* probs, int n, int H, int R, cudaStream_t s);
int rank_proj_dropout_dgrad(const RankProjProblem* probs, int n, int H, int R, cudaStream_t s);
int rank_proj_dropout_wgrad(const RankProjProblem* probs, int n, int H, int R, cudaStream_t s);
int rank_proj_dropout_mask(uint8_t* keep, long M, int H, int R, float p, uint64_t seed, uint64_t site, cudaStream_t s);

// optim.cu  (multi-tensor trainer tail; pointer / chunk tables live in device memory)
int grad_sumsq_multi(const float* const* g_ptrs, const long* numel, const int* chunk_tensor, const long* chunk_start,
                     int n_chunks, int chunk_elems, float* partials, float* sumsq, cudaStream_t s);
int adamax_multi(float* const* p_ptrs, const float* const* g_ptrs, float* const* m_ptrs, float* const* u_ptrs,
                 const long* numel, const int* chunk_tensor, const long* chunk_start, int n_chunks, int chunk_elems,
                 const float* sumsq, float inv_denom, float clip_norm, float clr, float beta1, float beta2, float eps,
                 float* norm_out, cudaStream_t s);

// gru.cu  (pointwise stages of a GRU timestep; the products run on gemm_bf16)
int gru_gate_fwd(const float* gx, long gx_row_stride, const float* gh, const float* h_prev, long hp_row_stride, float* h_out,
                 long ho_row_stride, __nv_bfloat16* h_bf16, __nv_bfloat16* r_s, __nv_bfloat16* z_s, __nv_bfloat16* n_s,
                 __nv_bfloat16* ghn_s, long rows, int H, cudaStream_t s);
int gru_gate_bwd(float* dh, const float* dout, long do_row_stride, const float* h_prev, long hp_row_stride,
                 const __nv_bfloat16* r_s, const __nv_bfloat16* z_s, const __nv_bfloat16* n_s, const __nv_bfloat16* ghn_s,
                 __nv_bfloat16* dgx, long dgx_row_stride, __nv_bfloat16* dgh, long rows, int H, cudaStream_t s);

// loss.cu
int kd_loss(const float* x, const void* teacher, int teacher_is_fp16, const float* target, float* dx, float* row_loss,
            float* loss, int B, int N, float T, float alpha, cudaStream_t s);

// softmax.cu
int masked_softmax_fwd(const float* logits, float* p, long rows, int len, cudaStream_t s);
int masked_softmax_bwd(const float* p, const float* dp, long dp_row_stride_b, long dp_row_stride_g, long dp_elem_stride,
                       float* dlogits, long batch, int groups, int len, cudaStream_t s);

// trilinear.cu
struct TriDims {
  int B, K, Q, A, G, R;   // d is fixed at 16
  int VR = 1;             // v_rep: rows b of q / a / logits use row b / VR of vc and rowmask (vc has B / VR samples)
};
int trilinear_fwd(const __nv_bfloat16* vc, const __nv_bfloat16* qc, const __nv_bfloat16* ac, const __nv_bfloat16* tpack,
                  const __nv_bfloat16* tpack_perm, const uint8_t* rowmask, float* logits, void* n1_save, TriDims d,
                  cudaStream_t s);
size_t trilinear_n1_bytes(TriDims d);      // size of n1_save (0: shape outside the tcgen05 path)
size_t trilinear_bwd_workspace(TriDims d);
int debug_prof_read(unsigned long long* host_dst, int n);   // CTI_PROF builds only (returns -1 otherwise)
int debug_prof_read_bwd1(unsigned long long* host_dst, int n);
int trilinear_bwd(const __nv_bfloat16* vc, const __nv_bfloat16* qc, const __nv_bfloat16* ac, const __nv_bfloat16* tpack,
                  const float* dlogits, const void* n1_saved, __nv_bfloat16* dzv, __nv_bfloat16* dzq, __nv_bfloat16* dza,
                  float* dbv, float* dbq, float* dba, float* dtpack, void* workspace, size_t workspace_bytes, TriDims d,
                  cudaStream_t s);

// pool.cu  (A == 0 selects the bilinear pooling of BCNet.forward_with_weights)
struct PoolDims {
  int B, K, Q, A, C;
  int VR = 1;             // v_rep: row b uses sample b / VR of v (v has B / VR samples); dzv stays per row b
};
int tri_pool_fwd(const __nv_bfloat16* v, const __nv_bfloat16* q, const __nv_bfloat16* a, const float* w, long w_stride_b,
                 float* out, PoolDims d, cudaStream_t s);
int tri_pool_bwd(const __nv_bfloat16* v, const __nv_bfloat16* q, const __nv_bfloat16* a, const float* w, long w_stride_b,
                 const float* dout, __nv_bfloat16* dzv, __nv_bfloat16* dzq, __nv_bfloat16* dza, float* dbv, float* dbq,
                 float* dba, float* dw, long dw_stride_b, PoolDims d, cudaStream_t s);   // dw_stride_b 0 = contiguous

// bilinear.cu
struct BiDims {
  int B, K, Q, G, C;
};
int bilinear_fwd(const __nv_bfloat16* vb, const __nv_bfloat16* qb, const float* hmat, const float* hbias,
                 const uint8_t* rowmask, float* logits, BiDims d, cudaStream_t s);
int bilinear_bwd(const __nv_bfloat16* vb, const __nv_bfloat16* qb, const float* hmat, const float* dlogits,
                 __nv_bfloat16* dzv, __nv_bfloat16* dzq, float* dbv, float* dbq, float* dhmat, float* dhbias, BiDims d,
                 cudaStream_t s);

// peer.cu
int peer_alloc(size_t bytes, void** ptr);
int peer_free(void* ptr);
int peer_export(void* ptr, void* handle64);
int peer_import(const void* handle64, void** ptr);
int peer_close(void* ptr);
int peer_barrier(void* const* flag_blocks, int rank, int world, int slot, double timeout_s, cudaStream_t s);
int peer_barrier_memops(void* const* flag_blocks, int rank, int world, int slot, cudaStream_t s);
int peer_flag_ops(void* flag_block, const int* index, const uint32_t* value, const int* wait, int count, cudaStream_t s);
int peer_flag_op(void* flag_block, int index, uint32_t value, int wait, cudaStream_t s);
int peer_error(const void* flag_block, int* out);
int peer_stamp(unsigned long long* dst, cudaStream_t s);
int peer_allreduce_fused(void* const* flag_blocks, void* const* slab_ranges, void* const* stagings, int rank, int world,
                         long n, double timeout_s, cudaStream_t s);
int peer_copy(void* dst, const void* src, size_t bytes, cudaStream_t s);
int sum_staged(float* dst, const float* staged, int n_staged, int rank, long n, long stride, cudaStream_t s);

}  // namespace cti
