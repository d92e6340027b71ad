// Masked softmax over the flattened attention domain (sm_100a, HBM-bound).
//   forward : p = softmax(row)            rows already hold -inf at masked regions
//             (reference src/attention.py:55-58 TriAttention, :36-39 BiAttention)
//   backward: dlogits = p * (dp - sum(p * dp))     
// One CTA per row, the row lives in registers (128-bit loads), max / sum by
// warp-shuffle + one shared-memory hop.  Row length 3600 (MC), 1800 (FFOE), 600 (BAN).
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {

namespace {

constexpr int kThreads = 256;
constexpr int kVec = 4;            // float4 per thread held in registers -> rows up to 4096 floats
constexpr int kMaxRegLen = kThreads * kVec * 4;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();   // protect sh from the previous reduction
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < kThreads / 32) ? sh[l] : (is_max ? -INFINITY : 0.f);
  t = is_max ? warp_max(t) : warp_sum(t);
  return t;
}

// Vector path: len % 4 == 0, len <= kMaxRegLen.
__global__ void __launch_bounds__(kThreads) softmax_fwd_vec_kernel(const float* __restrict__ logits,
                                                                   float* __restrict__ p, int len) {
  pdl_prologue_done();
  __shared__ float sh[8];
  const float4* in = reinterpret_cast<const float4*>(logits + static_cast<size_t>(blockIdx.x) * len);
  float4* out = reinterpret_cast<float4*>(p + static_cast<size_t>(blockIdx.x) * len);
  const int n4 = len >> 2;
  float4 r[kVec];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    if (idx < n4) {
      r[i] = __ldcs(in + idx);
      mx = fmaxf(mx, fmaxf(fmaxf(r[i].x, r[i].y), fmaxf(r[i].z, r[i].w)));
    }
  }
  mx = block_reduce(mx, true, sh);
  constexpr float kLog2e = 1.4426950408889634f;
  const float mb = mx * kLog2e;   // a fully masked row gives (-inf) - (-inf) = NaN, like the reference
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    if (idx < n4) {
      r[i].x = exp2f(r[i].x * kLog2e - mb);
      r[i].y = exp2f(r[i].y * kLog2e - mb);
      r[i].z = exp2f(r[i].z * kLog2e - mb);
      r[i].w = exp2f(r[i].w * kLog2e - mb);
      sum += (r[i].x + r[i].y) + (r[i].z + r[i].w);
    }
  }
  sum = block_reduce(sum, false, sh);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    if (idx < n4) out[idx] = make_float4(r[i].x * inv, r[i].y * inv, r[i].z * inv, r[i].w * inv);
  }
}

// Generic path: any length, three passes over global/L2.
__global__ void __launch_bounds__(kThreads) softmax_fwd_gen_kernel(const float* __restrict__ logits,
                                                                   float* __restrict__ p, int len) {
  pdl_prologue_done();
  __shared__ float sh[8];
  const float* in = logits + static_cast<size_t>(blockIdx.x) * len;
  float* out = p + static_cast<size_t>(blockIdx.x) * len;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < len; i += kThreads) mx = fmaxf(mx, in[i]);
  mx = block_reduce(mx, true, sh);
  float sum = 0.f;
  for (int i = threadIdx.x; i < len; i += kThreads) sum += expf(in[i] - mx);
  sum = block_reduce(sum, false, sh);
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < len; i += kThreads) out[i] = expf(in[i] - mx) * inv;
}

// Backward. p and dlogits are contiguous rows (b, g); dp is addressed through strides so the
// gradient can arrive either in the kernel's own (B,G,L) order or in the logical (B,L,G) order
// autograd produces for slices of the returned attention.
__global__ void __launch_bounds__(kThreads) softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                               long sb, long sg, long se, float* __restrict__ dlogits,
                                                               int groups, int len, int vec_ok) {
  pdl_prologue_done();
  __shared__ float sh[8];
  const long b = blockIdx.x / groups;
  const int g = blockIdx.x - static_cast<int>(b) * groups;
  const float* pr = p + static_cast<size_t>(blockIdx.x) * len;
  const float* dr = dp + b * sb + g * sg;
  float* out = dlogits + static_cast<size_t>(blockIdx.x) * len;
  float dot = 0.f;
  if (vec_ok && se == 1 && (len & 3) == 0 && ((sb | sg) & 3) == 0 && len <= kMaxRegLen) {
    const float4* p4 = reinterpret_cast<const float4*>(pr);
    const float4* d4 = reinterpret_cast<const float4*>(dr);
    float4* o4 = reinterpret_cast<float4*>(out);
    const int n4 = len >> 2;
    float4 rp[kVec], rd[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      if (idx < n4) {
        rp[i] = __ldcs(p4 + idx);
        rd[i] = __ldcs(d4 + idx);
        dot += rp[i].x * rd[i].x + rp[i].y * rd[i].y + rp[i].z * rd[i].z + rp[i].w * rd[i].w;
      }
    }
    dot = block_reduce(dot, false, sh);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      if (idx < n4)
        o4[idx] = make_float4(rp[i].x * (rd[i].x - dot), rp[i].y * (rd[i].y - dot), rp[i].z * (rd[i].z - dot),
                              rp[i].w * (rd[i].w - dot));
    }
    return;
  }
  for (int i = threadIdx.x; i < len; i += kThreads) dot += pr[i] * dr[static_cast<long>(i) * se];
  dot = block_reduce(dot, false, sh);
  for (int i = threadIdx.x; i < len; i += kThreads) out[i] = pr[i] * (dr[static_cast<long>(i) * se] - dot);
}

}  // namespace

int masked_softmax_fwd(const float* logits, float* p, long rows, int len, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && len > 0, "masked_softmax_fwd: bad shape rows=%ld len=%d", rows, len);
  if (rows == 0) return 0;
  CTI_REQUIRE(rows < (1l << 31), "masked_softmax_fwd: too many rows");
  const bool vec = (len % 4 == 0) && len <= kMaxRegLen && (((uintptr_t)logits | (uintptr_t)p) & 15) == 0;
  if (vec) launch_pdl(softmax_fwd_vec_kernel, dim3((unsigned)rows), dim3(kThreads), 0, s, logits, p, len);
  else     launch_pdl(softmax_fwd_gen_kernel, dim3((unsigned)rows), dim3(kThreads), 0, s, logits, p, len);
  return check_launch("masked_softmax_fwd");
}

int masked_softmax_bwd(const float* p, const float* dp, long sb, long sg, long se, float* dlogits, long batch,
                       int groups, int len, cudaStream_t s) {
  CTI_REQUIRE(batch >= 0 && groups > 0 && len > 0, "masked_softmax_bwd: bad shape");
  if (batch == 0) return 0;
  CTI_REQUIRE(batch * groups < (1l << 31), "masked_softmax_bwd: too many rows");
  const int aligned = ((((uintptr_t)p | (uintptr_t)dp | (uintptr_t)dlogits) & 15) == 0) ? 1 : 0;
  launch_pdl(softmax_bwd_kernel, dim3((unsigned)(batch * groups)), dim3(kThreads), 0, s, p, dp, sb, sg, se, dlogits, groups, len, aligned);
  return check_launch("masked_softmax_bwd");
}

}  // namespace cti
