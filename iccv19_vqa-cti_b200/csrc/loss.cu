// Distillation loss of the BAN student / CTI teacher (reference src/loss_function.py:12-25), forward and gradient in
// one pass over the class logits (HBM-bound: x, teacher and target are read once from DRAM, dx written once):
//
//   loss = mean_b( sum_c pt (log pt - log_softmax(x / T)) ) * alpha T^2
//        + sum_{b,c} bce_with_logits(x, y) / B * (1 - alpha),            pt = softmax(teacher / T)
//   dx   = alpha T / B * (softmax(x / T) - pt) + (1 - alpha) / B * (sigmoid(x) - y)
//
// One CTA per row (3129 classes for VQA 2.0, 1484 TDIUC-shaped); the teacher logits arrive as fp16 (the wire format
// of the reference's teacher-logit files, src/FFOE/test.py:129) or fp32.  Row losses go to a buffer that a second
// single-CTA kernel sums in a fixed order (deterministic loss).
#include <cuda_fp16.h>

#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kThreads / 32) ? sh[l] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max(t) : warp_sum(t);
}

template <bool kHalfTeacher>
__global__ void __launch_bounds__(kThreads) kd_loss_kernel(const float* __restrict__ x, const void* __restrict__ teacher,
                                                           const float* __restrict__ target, float* __restrict__ dx,
                                                           float* __restrict__ row_loss, int N, float inv_T, float alpha,
                                                           float inv_B) {
  pdl_prologue_done();
  __shared__ float sh[8];
  const size_t off = static_cast<size_t>(blockIdx.x) * N;
  const float* xr = x + off;
  const float* yr = target + off;
  auto tch = [&](int i) -> float {
    if (kHalfTeacher) return __half2float(reinterpret_cast<const __half*>(teacher)[off + i]);
    return reinterpret_cast<const float*>(teacher)[off + i];
  };
  float mx = -INFINITY, mt = -INFINITY;
  for (int i = threadIdx.x; i < N; i += kThreads) {
    mx = fmaxf(mx, xr[i] * inv_T);
    mt = fmaxf(mt, tch(i) * inv_T);
  }
  mx = block_reduce(mx, true, sh);
  mt = block_reduce(mt, true, sh);
  float sx = 0.f, st = 0.f;
  for (int i = threadIdx.x; i < N; i += kThreads) {
    sx += __expf(xr[i] * inv_T - mx);
    st += __expf(tch(i) * inv_T - mt);
  }
  sx = block_reduce(sx, false, sh);
  st = block_reduce(st, false, sh);
  const float lsx = __logf(sx), lst = __logf(st), isx = 1.f / sx, ist = 1.f / st;
  const float T = 1.f / inv_T;
  const float c_kd = alpha * T * inv_B, c_bce = (1.f - alpha) * inv_B;
  float kl = 0.f, bce = 0.f;
  for (int i = threadIdx.x; i < N; i += kThreads) {
    const float xi = xr[i], zi = xi * inv_T - mx, ti = tch(i) * inv_T - mt, yi = yr[i];
    const float pt = __expf(ti) * ist, px = __expf(zi) * isx;
    if (pt > 0.f) kl += pt * ((ti - lst) - (zi - lsx));          // KLDivLoss(reduction='none'): 0 where the target is 0
    bce += fmaxf(xi, 0.f) - xi * yi + log1pf(__expf(-fabsf(xi)));
    const float sg = 1.f / (1.f + __expf(-xi));
    dx[off + i] = c_kd * (px - pt) + c_bce * (sg - yi);
  }
  kl = block_reduce(kl, false, sh);
  bce = block_reduce(bce, false, sh);
  if (threadIdx.x == 0) row_loss[blockIdx.x] = kl * (alpha * T * T) * inv_B + bce * c_bce;
}

__global__ void __launch_bounds__(kThreads) sum_rows_kernel(const float* __restrict__ row_loss, float* __restrict__ out, int B) {
  pdl_prologue_done();
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += kThreads) s += row_loss[i];
  s = block_reduce(s, false, sh);
  if (threadIdx.x == 0) out[0] = s;
}

}  // namespace

int kd_loss(const float* x, const void* teacher, int teacher_is_fp16, const float* target, float* dx, float* row_loss,
            float* loss, int B, int N, float T, float alpha, cudaStream_t s) {
  CTI_REQUIRE(B > 0 && N > 0 && T > 0.f, "kd_loss: bad shape / temperature (B=%d N=%d T=%f)", B, N, T);
  if (teacher_is_fp16)
    launch_pdl(kd_loss_kernel<true>, dim3(B), dim3(kThreads), 0, s, x, teacher, target, dx, row_loss, N, 1.f / T, alpha, 1.f / B);
  else
    launch_pdl(kd_loss_kernel<false>, dim3(B), dim3(kThreads), 0, s, x, teacher, target, dx, row_loss, N, 1.f / T, alpha, 1.f / B);
  if (int rc = check_launch("kd_loss_kernel")) return rc;
  launch_pdl(sum_rows_kernel, dim3(1), dim3(kThreads), 0, s, (const float*)row_loss, loss, B);
  return check_launch("kd_sum_rows_kernel");
}

}  // namespace cti
