// Trainer tail as two multi-tensor passes over ALL parameters:
//   1. global gradient norm            (reference: torch.norm of the flat gradient copy, src/utils.py:324)
//   2. rescale + clip + Adamax update  (reference: flat.div_(grad_denom), clip_grad_norm_, copy back, src/MC/trainer.py:208-219,
//                                       then torch.optim.Adamax.step, src/MC/train.py:32)
// The reference flattens 346 gradients into one buffer, clips, copies them back and runs a per-tensor optimizer
// (~10 elementwise passes); here every gradient is read twice and every parameter / moment once: 32 B per element.
// A "chunk" is up to chunk_elems consecutive elements of one tensor; blockIdx.x = chunk.  Scalar, fully coalesced
// accesses (tensors may start at any 4-byte offset inside the flat buckets).  No atomics: results are reproducible.
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {

namespace {

__global__ void __launch_bounds__(256)
grad_sumsq_multi_kernel(const float* const* __restrict__ g_ptrs, const long* __restrict__ numel,
                        const int* __restrict__ chunk_tensor, const long* __restrict__ chunk_start, int chunk_elems,
                        float* __restrict__ partials) {
  pdl_prologue_done();
  const int t = chunk_tensor[blockIdx.x];
  const long lo = chunk_start[blockIdx.x];
  const long hi = min(lo + chunk_elems, numel[t]);
  const float* g = g_ptrs[t];
  float s = 0.f;
  for (long base = lo + threadIdx.x; base < hi; base += 4 * 256) {      // four independent loads in flight per thread
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = (base + k * 256 < hi) ? g[base + k * 256] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s = fmaf(x[k], x[k], s);
  }
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

// sumsq[0] = sum of the partials in a fixed order (one block)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n, float* __restrict__ sumsq) {
  pdl_prologue_done();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) sumsq[0] = v;
  }
}

__global__ void __launch_bounds__(256)
adamax_multi_kernel(float* const* __restrict__ p_ptrs, const float* const* __restrict__ g_ptrs, float* const* __restrict__ m_ptrs,
                    float* const* __restrict__ u_ptrs, const long* __restrict__ numel, const int* __restrict__ chunk_tensor,
                    const long* __restrict__ chunk_start, int chunk_elems, const float* __restrict__ sumsq, float inv_denom,
                    float clip_norm, float clr, float beta1, float beta2, float eps, float* __restrict__ norm_out) {
  pdl_prologue_done();
  const int t = chunk_tensor[blockIdx.x];
  const long lo = chunk_start[blockIdx.x];
  const long hi = min(lo + chunk_elems, numel[t]);
  // norm of (g / denom); clip_grad_norm_: scale by clip / (norm + 1e-6) only if norm > clip > 0 (src/utils.py:325-327)
  const float norm = sqrtf(sumsq[0]) * inv_denom;
  const float coef = (clip_norm > 0.f && norm > clip_norm) ? clip_norm / (norm + 1e-6f) : 1.f;
  const float scale = inv_denom * coef;
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
  float* p = p_ptrs[t];
  const float* g = g_ptrs[t];
  float* m = m_ptrs[t];
  float* u = u_ptrs[t];
  const float omb1 = 1.f - beta1;
  for (long base = lo + threadIdx.x; base < hi; base += 4 * 256) {      // 16 independent loads in flight per thread
    float gv[4], mv[4], uv[4], pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long i = base + k * 256;
      const bool ok = i < hi;
      gv[k] = ok ? g[i] : 0.f;
      mv[k] = ok ? m[i] : 0.f;
      uv[k] = ok ? u[i] : 1.f;
      pv[k] = ok ? p[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long i = base + k * 256;
      if (i < hi) {
        const float gg = gv[k] * scale;
        const float mi = fmaf(omb1, gg - mv[k], mv[k]);            // exp_avg.lerp_(grad, 1 - beta1)
        const float ui = fmaxf(uv[k] * beta2, fabsf(gg) + eps);    // exp_inf = max(beta2 * exp_inf, |grad| + eps)
        m[i] = mi;
        u[i] = ui;
        p[i] = pv[k] - clr * (mi / ui);                            // param.addcdiv_(exp_avg, exp_inf, value=-clr)
      }
    }
  }
}

}  // namespace

int grad_sumsq_multi(const float* const* g_ptrs, const long* numel, const int* chunk_tensor, const long* chunk_start,
                     int n_chunks, int chunk_elems, float* partials, float* sumsq, cudaStream_t s) {
  CTI_REQUIRE(n_chunks > 0 && chunk_elems > 0, "grad_sumsq_multi: empty chunk table");
  launch_pdl(grad_sumsq_multi_kernel, dim3(n_chunks), dim3(256), 0, s, g_ptrs, numel, chunk_tensor, chunk_start, chunk_elems, partials);
  int rc = check_launch("grad_sumsq_multi_kernel");
  if (rc) return rc;
  launch_pdl(reduce_partials_kernel, dim3(1), dim3(256), 0, s, partials, n_chunks, sumsq);
  return check_launch("reduce_partials_kernel");
}

int adamax_multi(float* const* p_ptrs, const float* const* g_ptrs, float* const* m_ptrs, float* const* u_ptrs,
                 const long* numel, const int* chunk_tensor, const long* chunk_start, int n_chunks, int chunk_elems,
                 const float* sumsq, float inv_denom, float clip_norm, float clr, float beta1, float beta2, float eps,
                 float* norm_out, cudaStream_t s) {
  CTI_REQUIRE(n_chunks > 0 && chunk_elems > 0, "adamax_multi: empty chunk table");
  CTI_REQUIRE(sumsq != nullptr, "adamax_multi: the squared gradient norm (cti_grad_sumsq_multi) is required");
  launch_pdl(adamax_multi_kernel, dim3(n_chunks), dim3(256), 0, s, p_ptrs, g_ptrs, m_ptrs, u_ptrs, numel, chunk_tensor, chunk_start, chunk_elems,
                                               sumsq, inv_denom, clip_norm, clr, beta1, beta2, eps, norm_out);
  return check_launch("adamax_multi_kernel");
}

}  // namespace cti
