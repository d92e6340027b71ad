// Caller glue of the glimpse loop (reference src/MC/base_model.py:145-150, src/FFOE/base_model.py:125-130) as three
// streaming kernels, for the opt-in fused call (glimpse.py):
//
//   q_emb   = q_prj[g](b_emb[g].unsqueeze(1)) + q_emb      -> only its bf16 copy is ever consumed (next glimpse's
//   ans_emb = a_prj[g](b_emb[g].unsqueeze(1)) + ans_emb       q_tucker / a_tucker GEMM): glimpse_residual_cast
//   joint   = q_emb.sum(1) + ans_emb.sum(1)                -> glimpse_token_sum (the residual terms are re-added in
//                                                             registers, the (B, T, D) running sums never exist)
//   backward: d q_emb starts as the joint gradient broadcast over the tokens (glimpse_bcast_rows); the dgrad GEMMs of
//   every glimpse accumulate into it (TMA reduce-add), glimpse_token_sum of it is d(q_prj output).
//
// All HBM-bound: 16-byte accesses, one pass over each (B, T, D) tensor.
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {
namespace {

struct ResidualPtrs {
  const float* p[4];
  int n;
};

__device__ __forceinline__ void load8(const void* x, int is_bf16, long i8, float (&v)[8]) {
  if (is_bf16) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + i8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      v[2 * t] = f.x;
      v[2 * t + 1] = f.y;
    }
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i8);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i8 + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

// residual terms of row b, columns 8 * c8 .. + 7, in the caller's order of addition: ((x + p0) + p1) + ...
__device__ __forceinline__ void load_res(const ResidualPtrs& r, long b, int c8, int D8, float (&pv)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < r.n) load8(r.p[i], 0, b * D8 + c8, pv[i]);
}

// out[b, t, :] = bf16(x[b, t, :] + p0[b, :] + p1[b, :] + ...), two tensors (question / answer tokens) per launch
__global__ void __launch_bounds__(256)
glimpse_residual_cast_kernel(const void* __restrict__ xq, int q_bf16, ResidualPtrs rq, uint4* __restrict__ oq, long nq8, int Tq,
                             const void* __restrict__ xa, int a_bf16, ResidualPtrs ra, uint4* __restrict__ oa, long na8, int Ta,
                             int D8) {
  pdl_prologue_done();
  long i = blockIdx.x * 256l + threadIdx.x;
  const bool second = i >= nq8;                 // the answer-token tensor follows the question-token tensor
  if (second) {
    i -= nq8;
    if (i >= na8) return;
  }
  const void* x = second ? xa : xq;
  const int bf = second ? a_bf16 : q_bf16, T = second ? Ta : Tq, n = second ? ra.n : rq.n;
  uint4* o = second ? oa : oq;
  const long row = i / D8;
  const int c8 = (int)(i - row * D8);
  const long b = row / T;
  float v[8];
  load8(x, bf, i, v);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n) {
      float pv[8];
      load8(second ? ra.p[k] : rq.p[k], 0, b * D8 + c8, pv);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += pv[e];
    }
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  o[i] = u;
}

// out[b, :] = sum_t (xq[b, t, :] + rq...) + sum_t (xa[b, t, :] + ra...); xa may be null.
// One thread per (b, 8 columns): B * D / 8 threads, each token row a coalesced 16 / 32-byte access.
__global__ void __launch_bounds__(256)
glimpse_token_sum_kernel(const void* __restrict__ xq, int q_bf16, ResidualPtrs rq, int Tq, const void* __restrict__ xa,
                         int a_bf16, ResidualPtrs ra, int Ta, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16,
                         long B, int D8) {
  pdl_prologue_done();
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= B * D8) return;
  const long b = i / D8;
  const int c8 = (int)(i - b * D8);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const void* x = side ? xa : xq;
    if (x == nullptr) continue;
    const int bf = side ? a_bf16 : q_bf16, T = side ? Ta : Tq;
    const ResidualPtrs& r = side ? ra : rq;
    float pv[4][8];
    load_res(r, b, c8, D8, pv);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < T; ++t) {
      float v[8];
      load8(x, bf, (b * T + t) * D8 + c8, v);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < r.n) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += pv[k][e];
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += s[e];
  }
  if (out != nullptr) {
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  if (out_bf16 != nullptr) {
    uint4 u;
    u.x = pack_bf16x2(acc[0], acc[1]);
    u.y = pack_bf16x2(acc[2], acc[3]);
    u.z = pack_bf16x2(acc[4], acc[5]);
    u.w = pack_bf16x2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out_bf16)[i] = u;
  }
}

// out[b, t, :] = x[b, :] for t < T (fp32): the joint gradient as the initial value of d q_emb / d ans_emb
__global__ void __launch_bounds__(256)
glimpse_bcast_rows_kernel(const float4* __restrict__ x, float4* __restrict__ oq, int Tq, float4* __restrict__ oa, int Ta,
                          long B, int D4) {
  pdl_prologue_done();
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= B * D4) return;
  const long b = i / D4;
  const int c = (int)(i - b * D4);
  const float4 v = __ldg(x + i);
  for (int t = 0; t < Tq; ++t) oq[(b * Tq + t) * D4 + c] = v;
  if (oa != nullptr)
    for (int t = 0; t < Ta; ++t) oa[(b * Ta + t) * D4 + c] = v;
}

int make_res(const float* const* p, int n, ResidualPtrs& r, const char* who) {
  CTI_REQUIRE(n >= 0 && n <= 4, "%s: at most 4 residual terms (got %d)", who, n);
  r.n = n;
  for (int i = 0; i < 4; ++i) {
    r.p[i] = i < n ? p[i] : nullptr;
    CTI_REQUIRE(i >= n || (p[i] != nullptr && ((uintptr_t)p[i] & 15) == 0), "%s: residual term %d must be a 16-byte aligned pointer", who, i);
  }
  return 0;
}

}  // namespace

int glimpse_residual_cast(const void* xq, int q_bf16, const float* const* pq, int Tq, __nv_bfloat16* oq, const void* xa,
                          int a_bf16, const float* const* pa, int Ta, __nv_bfloat16* oa, int n_res, long B, int D,
                          cudaStream_t s) {
  CTI_REQUIRE(B >= 0 && D > 0 && (D & 7) == 0 && Tq > 0 && (xa == nullptr || Ta > 0), "glimpse_residual_cast: bad shape (D=%d must be a multiple of 8)", D);
  if (B == 0) return 0;
  ResidualPtrs rq, ra;
  if (int rc = make_res(pq, n_res, rq, "glimpse_residual_cast")) return rc;
  if (int rc = make_res(pa, xa ? n_res : 0, ra, "glimpse_residual_cast")) return rc;
  CTI_REQUIRE(((uintptr_t)xq & 15) == 0 && ((uintptr_t)oq & 15) == 0 && ((uintptr_t)xa & 15) == 0 && ((uintptr_t)oa & 15) == 0,
              "glimpse_residual_cast: buffers must be 16-byte aligned");
  const long nq8 = B * Tq * (D / 8), na8 = xa ? B * Ta * (D / 8) : 0;
  const long blocks = (nq8 + na8 + 255) / 256;
  CTI_REQUIRE(blocks < (1l << 31), "glimpse_residual_cast: too many elements");
  launch_pdl(glimpse_residual_cast_kernel, dim3((unsigned)blocks), dim3(256), 0, s, xq, q_bf16, rq, reinterpret_cast<uint4*>(oq),
             nq8, Tq, xa, a_bf16, ra, reinterpret_cast<uint4*>(oa), na8, Ta, D / 8);
  return check_launch("glimpse_residual_cast_kernel");
}

int glimpse_token_sum(const void* xq, int q_bf16, const float* const* pq, int Tq, const void* xa, int a_bf16,
                      const float* const* pa, int Ta, int n_res, float* out, __nv_bfloat16* out_bf16, long B, int D,
                      cudaStream_t s) {
  CTI_REQUIRE(B >= 0 && D > 0 && (D & 7) == 0 && Tq > 0 && (xa == nullptr || Ta > 0), "glimpse_token_sum: bad shape (D=%d must be a multiple of 8)", D);
  CTI_REQUIRE(out != nullptr || out_bf16 != nullptr, "glimpse_token_sum: no output");
  if (B == 0) return 0;
  ResidualPtrs rq, ra;
  if (int rc = make_res(pq, n_res, rq, "glimpse_token_sum")) return rc;
  if (int rc = make_res(pa, xa ? n_res : 0, ra, "glimpse_token_sum")) return rc;
  CTI_REQUIRE(((uintptr_t)xq & 15) == 0 && ((uintptr_t)xa & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0,
              "glimpse_token_sum: buffers must be 16-byte aligned");
  const long n = B * (D / 8);
  launch_pdl(glimpse_token_sum_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, xq, q_bf16, rq, Tq, xa, a_bf16, ra,
             Ta, out, out_bf16, B, D / 8);
  return check_launch("glimpse_token_sum_kernel");
}

int glimpse_bcast_rows(const float* x, float* oq, int Tq, float* oa, int Ta, long B, int D, cudaStream_t s) {
  CTI_REQUIRE(B >= 0 && D > 0 && (D & 3) == 0 && Tq > 0 && (oa == nullptr || Ta > 0), "glimpse_bcast_rows: bad shape");
  if (B == 0) return 0;
  CTI_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)oq & 15) == 0 && ((uintptr_t)oa & 15) == 0,
              "glimpse_bcast_rows: buffers must be 16-byte aligned");
  const long n = B * (D / 4);
  launch_pdl(glimpse_bcast_rows_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, reinterpret_cast<const float4*>(x),
             reinterpret_cast<float4*>(oq), Tq, reinterpret_cast<float4*>(oa), Ta, B, D / 4);
  return check_launch("glimpse_bcast_rows_kernel");
}

}  // namespace cti
