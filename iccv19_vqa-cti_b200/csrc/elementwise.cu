// HBM-bound helper kernels of the CTI hot path (sm_100a):
//   cast_rows_mask : fp32 -> bf16 row cast fused with the zero-row mask of
//                    reference src/attention.py:55 / :36  (v.abs().sum(2) == 0)
//   wn_pack        : weight-norm fold  W_eff = V * g/||V||_F -> bf16  (src/fc.py:22,27, dim=None)
//   wn_grad        : weight-norm backward  dV, dg from dW_eff          
//   act_bwd_bias   : dz = dy * [y > 0]  (ReLU backward), bf16 out, fused bias gradient
// All are single-pass, 128-bit vectorised and coalesced; grids are sized from the data.
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {

namespace {

// ------------------------------------------------------------------------- //
__global__ void __launch_bounds__(256) cast_rows_mask_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                             uint8_t* __restrict__ rowmask, long rows, int cols) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * cols;
  __nv_bfloat16* orow = out + row * cols;
  bool nz = false;
  if ((cols & 7) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xr);
    uint4* o4 = reinterpret_cast<uint4*>(orow);
    const int n8 = cols >> 3;
#pragma unroll 4
    for (int i = lane; i < n8; i += 32) {
      const float4 a = __ldcs(x4 + 2 * i);
      const float4 b = __ldcs(x4 + 2 * i + 1);
      nz |= (a.x != 0.f) | (a.y != 0.f) | (a.z != 0.f) | (a.w != 0.f) | (b.x != 0.f) | (b.y != 0.f) | (b.z != 0.f) |
            (b.w != 0.f);
      uint4 u;
      u.x = pack_bf16x2(a.x, a.y);
      u.y = pack_bf16x2(a.z, a.w);
      u.z = pack_bf16x2(b.x, b.y);
      u.w = pack_bf16x2(b.z, b.w);
      o4[i] = u;
    }
  } else {
    for (int i = lane; i < cols; i += 32) {
      const float f = xr[i];
      nz |= (f != 0.f);
      orow[i] = __float2bfloat16(f);
    }
  }
  if (rowmask != nullptr) {
    const unsigned any = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) rowmask[row] = (any == 0u) ? 1 : 0;
  }
}

// Zero-row mask of features that already ARE bf16 (the loader's wire format, loader.py): rowmask[r] = 1 iff every
// element of row r is +-0.  One warp per row, 16-byte loads, read-only.
__global__ void __launch_bounds__(256) rowmask_bf16_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ rowmask,
                                                           long rows, int cols) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* x4 = reinterpret_cast<const uint4*>(x + row * cols);
  uint32_t acc = 0;
#pragma unroll 4
  for (int i = lane; i < (cols >> 3); i += 32) {
    const uint4 u = __ldcs(x4 + i);
    acc |= (u.x | u.y | u.z | u.w) & 0x7FFF7FFFu;        // drop the sign bits: -0 counts as zero
  }
  const unsigned any = __ballot_sync(0xffffffffu, acc != 0u);
  if (lane == 0) rowmask[row] = (any == 0u) ? 1 : 0;
}

// ------------------------------------------------------------------------- //
// Dropout (training mode): keep mask = Philox4x32-10(seed, (element / 4, offset)) word (element % 4) >= p * 2^32,
// kept values scaled by 1 / (1 - p).  The mask is a pure function of (seed, offset, element index), so the
// backward kernels regenerate it instead of storing it (reference: nn.Dropout on the FCNet input, src/fc.py:25-26).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

struct DropoutRng {
  uint2 key;
  uint32_t offset, threshold;
  float scale;
};
__host__ inline DropoutRng make_rng(float p, uint64_t seed, uint64_t offset) {
  DropoutRng r;
  r.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  r.offset = (uint32_t)offset;
  const double t = (double)p * 4294967296.0;
  r.threshold = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  r.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return r;
}
// keep flags (times scale) of the 4 consecutive elements starting at element index 4 * quad
__device__ __forceinline__ float4 dropout_scale4(const DropoutRng& r, uint64_t quad) {
  const uint4 w = philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), r.offset, 0u), r.key);
  return make_float4(w.x >= r.threshold ? r.scale : 0.f, w.y >= r.threshold ? r.scale : 0.f,
                     w.z >= r.threshold ? r.scale : 0.f, w.w >= r.threshold ? r.scale : 0.f);
}

// out = bf16(dropout(x)), rowmask from the UNDROPPED row (cols % 8 == 0)
__global__ void __launch_bounds__(256) cast_rows_dropout_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                                uint8_t* __restrict__ rowmask, long rows, int cols,
                                                                const DropoutRng rng) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + row * cols);
  uint4* o4 = reinterpret_cast<uint4*>(out + row * cols);
  const uint64_t quad0 = static_cast<uint64_t>(row) * (cols >> 2);
  const int n8 = cols >> 3;
  bool nz = false;
  for (int i = lane; i < n8; i += 32) {
    const float4 a = __ldcs(x4 + 2 * i), b = __ldcs(x4 + 2 * i + 1);
    nz |= (a.x != 0.f) | (a.y != 0.f) | (a.z != 0.f) | (a.w != 0.f) | (b.x != 0.f) | (b.y != 0.f) | (b.z != 0.f) | (b.w != 0.f);
    const float4 ka = dropout_scale4(rng, quad0 + 2 * i), kb = dropout_scale4(rng, quad0 + 2 * i + 1);
    uint4 u;
    u.x = pack_bf16x2(a.x * ka.x, a.y * ka.y);
    u.y = pack_bf16x2(a.z * ka.z, a.w * ka.w);
    u.z = pack_bf16x2(b.x * kb.x, b.y * kb.y);
    u.w = pack_bf16x2(b.z * kb.z, b.w * kb.w);
    o4[i] = u;
  }
  if (rowmask != nullptr) {
    const unsigned any = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) rowmask[row] = (any == 0u) ? 1 : 0;
  }
}

// x[e] *= keep(e) / (1 - p), fp32 in place: backward of the input dropout (n % 4 == 0)
__global__ void __launch_bounds__(256) dropout_f32_kernel(float* __restrict__ x, long n4, const DropoutRng rng) {
  pdl_prologue_done();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<float4*>(x)[i];
  const float4 k = dropout_scale4(rng, (uint64_t)i);
  v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
  reinterpret_cast<float4*>(x)[i] = v;
}

// out = dropout(x), bf16 (n % 4 == 0); out may alias x
__global__ void __launch_bounds__(256) dropout_bf16_kernel(const __nv_bfloat16* x, __nv_bfloat16* out, long n4,
                                                           const DropoutRng rng) {
  pdl_prologue_done();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uint2 u = reinterpret_cast<const uint2*>(x)[i];
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  const float4 k = dropout_scale4(rng, (uint64_t)i);
  uint2 o;
  o.x = pack_bf16x2(a.x * k.x, a.y * k.y);
  o.y = pack_bf16x2(b.x * k.z, b.y * k.w);
  reinterpret_cast<uint2*>(out)[i] = o;
}

// Per-rank input dropout of the R per-rank nets (reference src/tc.py:29-31, 47-49): rank r sees x * keep_r with its
// OWN mask.  Mask of (rank r, row, column c): dropout element index (r * rows + row) * cols + c.
//   expand: xt[row, j * cols + c] = x[row, c] * keep_{r0 + j}[row, c] / (1 - p)       j < RG   (bf16)
//   reduce: acc[row, c] += sum_j dxt[row, j * cols + c] * keep_{r0 + j}[row, c] / (1 - p)       (fp32 accumulate)
__global__ void __launch_bounds__(256) dropout_expand_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xt,
                                                             long rows, int cols4, int RG, int r0, const DropoutRng rng) {
  pdl_prologue_done();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;      // (row, quad of 4 columns)
  if (i >= rows * cols4) return;
  const long row = i / cols4;
  const int c4 = static_cast<int>(i - row * cols4);
  const uint2 u = reinterpret_cast<const uint2*>(x)[i];
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  for (int j = 0; j < RG; ++j) {
    const float4 k = dropout_scale4(rng, (static_cast<uint64_t>(r0 + j) * rows + row) * cols4 + c4);
    uint2 o;
    o.x = pack_bf16x2(a.x * k.x, a.y * k.y);
    o.y = pack_bf16x2(b.x * k.z, b.y * k.w);
    reinterpret_cast<uint2*>(xt)[(row * RG + j) * cols4 + c4] = o;
  }
}

__global__ void __launch_bounds__(256) dropout_reduce_kernel(const __nv_bfloat16* __restrict__ dxt, float* __restrict__ acc,
                                                             long rows, int cols4, int RG, int r0, const DropoutRng rng) {
  pdl_prologue_done();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * cols4) return;
  const long row = i / cols4;
  const int c4 = static_cast<int>(i - row * cols4);
  float4 s = reinterpret_cast<float4*>(acc)[i];
  for (int j = 0; j < RG; ++j) {
    const uint2 u = reinterpret_cast<const uint2*>(dxt)[(row * RG + j) * cols4 + c4];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    const float4 k = dropout_scale4(rng, (static_cast<uint64_t>(r0 + j) * rows + row) * cols4 + c4);
    s.x += a.x * k.x; s.y += a.y * k.y; s.z += b.x * k.z; s.w += b.y * k.w;
  }
  reinterpret_cast<float4*>(acc)[i] = s;
}

// ------------------------------------------------------------------------- //
// weight norm.  A "group" is rows_per_group consecutive rows of the (n_groups*rows_per_group, cols)
// matrix; each group has its own scalar g and Frobenius norm (one group = one nn.Linear).
constexpr int kSeg = 4096;   // elements reduced by one block

// sum_{i in [lo, hi)} a[i] * b[i] of one thread of a 256-thread block, 16-byte loads (lo, hi and the bases are multiples
// of 4 elements: every group has a multiple of 4 elements).  One fixed order for the per-layer and the multi-tensor
// kernels: the same weights give the same sums bit for bit.
__device__ __forceinline__ float seg_dot(const float* __restrict__ a, const float* __restrict__ b, long lo, long hi) {
  float s = 0.f;
  for (long i = lo + 4 * threadIdx.x; i < hi; i += 4 * 256) {
    const float4 x = *reinterpret_cast<const float4*>(a + i);
    const float4 y = *reinterpret_cast<const float4*>(b + i);
    s = __fmaf_rn(x.x, y.x, s);          // explicit fma chain: the same rounding in every instantiation
    s = __fmaf_rn(x.y, y.y, s);
    s = __fmaf_rn(x.z, y.z, s);
    s = __fmaf_rn(x.w, y.w, s);
  }
  return s;
}

__device__ __forceinline__ void seg_reduce_store(float s, float* __restrict__ partial) {
  s = warp_sum(s);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ v, const float* __restrict__ w2,
                                                    float* __restrict__ partial, long group_elems, int segs_per_group) {
  pdl_prologue_done();
  // partial[group * segs + seg] = sum over the segment of v*w2 (w2 == v for the squared norm, == dW for <dW,V>).
  // No atomics: the same weights always give the same norm, hence the same bf16 pack (runs are reproducible).
  const int group = blockIdx.x / segs_per_group;
  const int seg = blockIdx.x - group * segs_per_group;
  const long base = static_cast<long>(group) * group_elems;
  const long lo = static_cast<long>(seg) * kSeg;
  const long hi = min(lo + kSeg, group_elems);
  const float s = seg_dot(v + base, w2 + base, lo, hi);
  seg_reduce_store(s, partial);
}

// Total of one group's segment partials, computed redundantly (and identically) by every consumer block instead of
// by a third launch: the block reduces the partials of the group its first element belongs to; a thread whose element
// lies in a later group (only when a group is smaller than, or not aligned to, the block's 1024 elements) sums its
// group's partials itself.  Fixed summation order: the same weights always give the same norm.
__device__ __forceinline__ float group_total(const float* __restrict__ partial, int segs, int block_group, int my_group,
                                             float* sh /* 9 floats */) {
  const float* p = partial + static_cast<long>(block_group) * segs;
  float s = 0.f;
  for (int i = threadIdx.x; i < segs; i += 256) s += p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    sh[8] = t;
  }
  __syncthreads();
  if (my_group == block_group) return sh[8];
  const float* q = partial + static_cast<long>(my_group) * segs;
  float t = 0.f;
  for (int i = 0; i < segs; ++i) t += q[i];
  return t;
}

__global__ void __launch_bounds__(256) wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                       float* __restrict__ sumsq, const float* __restrict__ partial, int segs,
                                                       __nv_bfloat16* __restrict__ w, long group_elems, long total) {
  pdl_prologue_done();
  __shared__ float sh[9];
  const long i0 = static_cast<long>(blockIdx.x) * 1024;
  long i = i0 + threadIdx.x * 4;
  const bool live = i < total;
  if (!live) i = total - 4;
  const int group = static_cast<int>(i / group_elems);   // group_elems % 4 == 0 is required by the caller
  const float n2 = group_total(partial, segs, static_cast<int>(i0 / group_elems), group, sh);
  if (!live) return;
  if (i % group_elems == 0) sumsq[group] = n2;            // kept for the backward of the fold
  const float s = g[group] * rsqrtf(n2);
  const float4 f = *reinterpret_cast<const float4*>(v + i);
  uint2 u;
  u.x = pack_bf16x2(f.x * s, f.y * s);
  u.y = pack_bf16x2(f.z * s, f.w * s);
  *reinterpret_cast<uint2*>(w + i) = u;
}

__global__ void __launch_bounds__(256) wn_grad_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                      const float* __restrict__ g, const float* __restrict__ sumsq,
                                                      const float* __restrict__ partial, int segs, float* __restrict__ dv,
                                                      float* __restrict__ dg, long group_elems, long total) {
  pdl_prologue_done();
  __shared__ float sh[9];
  const long i0 = static_cast<long>(blockIdx.x) * 1024;
  long i = i0 + threadIdx.x * 4;
  const bool live = i < total;
  if (!live) i = total - 4;
  const int group = static_cast<int>(i / group_elems);
  const float dt = group_total(partial, segs, static_cast<int>(i0 / group_elems), group, sh);   // <dW_eff, V>
  if (!live) return;
  const float n2 = sumsq[group];
  const float rn = rsqrtf(n2);
  const float s = g[group] * rn;          // g / ||V||
  const float c = dt / n2;
  const float4 a = *reinterpret_cast<const float4*>(dw + i);
  const float4 b = *reinterpret_cast<const float4*>(v + i);
  float4 o;
  o.x = s * (a.x - c * b.x);
  o.y = s * (a.y - c * b.y);
  o.z = s * (a.z - c * b.z);
  o.w = s * (a.w - c * b.w);
  *reinterpret_cast<float4*>(dv + i) = o;
  if (i % group_elems == 0) dg[group] = dt * rn;
}

// ------------------------------------------------------------------------- //
// dz[m,n] = dy[m,n] * (y[m,n] > 0);  dbias[n] += sum_m dz[m,n].
// Round 1 gave every thread two columns (4-byte loads): 0.06 of the copy bandwidth.  Now a thread owns EIGHT consecutive
// columns (16-byte bf16 / 2 x 16-byte fp32 loads), a block covers 8 x kAbbLanes columns x rows_per_block rows with
// 256 / kAbbLanes row lanes, the row lanes' partial column sums meet in shared memory and each block issues one atomic
// per column (the host sizes rows_per_block for a few waves of 148 SMs).
constexpr int kAbbLanes = 32;                 // column lanes per block: 32 x 8 = 256 columns
constexpr int kAbbRows = 256 / kAbbLanes;     // row lanes per block

template <bool DY_BF16>
__global__ void __launch_bounds__(256) act_bwd_bias_kernel(const void* __restrict__ dy_, const __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ dz, float* __restrict__ dbias,
                                                           long rows, int cols, int rows_per_block) {
  pdl_prologue_done();
  __shared__ float red[kAbbRows][kAbbLanes * 8 + 8];
  const int cl = threadIdx.x % kAbbLanes, rl = threadIdx.x / kAbbLanes;
  const int c = (blockIdx.x * kAbbLanes + cl) * 8;
  const long r0 = static_cast<long>(blockIdx.y) * rows_per_block;
  const long r1 = min(r0 + rows_per_block, rows);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {                                     // cols % 8 == 0 (checked by the caller)
#pragma unroll 4
    for (long r = r0 + rl; r < r1; r += kAbbRows) {
      const long off = r * cols + c;
      float d[8];
      if (DY_BF16) {
        const uint4 u = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(dy_) + off));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16x2(w[t]);
          d[2 * t] = f.x;
          d[2 * t + 1] = f.y;
        }
      } else {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + off));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + off) + 1);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      }
      if (y != nullptr) {
        const uint4 u = *reinterpret_cast<const uint4*>(y + off);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = unpack_bf16x2(w[t]);
          if (!(f.x > 0.f)) d[2 * t] = 0.f;
          if (!(f.y > 0.f)) d[2 * t + 1] = 0.f;
        }
      }
      if (dz != nullptr) {
        uint4 o;
        o.x = pack_bf16x2(d[0], d[1]); o.y = pack_bf16x2(d[2], d[3]); o.z = pack_bf16x2(d[4], d[5]); o.w = pack_bf16x2(d[6], d[7]);
        *reinterpret_cast<uint4*>(dz + off) = o;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) s[t] += d[t];
    }
  }
  if (dbias == nullptr) return;
#pragma unroll
  for (int t = 0; t < 8; ++t) red[rl][cl * 8 + t] = s[t];
  __syncthreads();
  // 256 threads, 256 columns: thread x sums column x over the row lanes
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < kAbbRows; ++j) tot += red[j][threadIdx.x];
  const int col = blockIdx.x * kAbbLanes * 8 + threadIdx.x;
  if (col < cols) atomicAdd(dbias + col, tot);
}

// ------------------------------------------------------------------------- //
// The fold of MANY layers in two launches (one entry = one group with its own scalar g; the R per-rank nets of a
// TCNet modality are R entries that write into one stacked pack).  Same arithmetic, in the same order, as
// sumsq_kernel + wn_scale_kernel: a layer packed here is bit-identical to the same layer packed on its own.
struct WnMultiTable {
  const float* const* v;        // [entries] weight_v of the entry
  const float* const* g;        // [entries] its scalar weight_g
  __nv_bfloat16* const* w;      // [entries] where its bf16 W_eff goes
  float* const* sumsq;          // [entries] where ||V||_F^2 goes (kept for the backward)
  const long* elems;            // [entries] elements of the entry (multiple of 4)
  const int* first_seg;         // [entries] index of its first partial
  const int* n_seg;             // [entries] number of kSeg-element segments
};

__global__ void __launch_bounds__(256) wn_multi_sumsq_kernel(const WnMultiTable t, const int* __restrict__ seg_entry,
                                                             const int* __restrict__ seg_index, float* __restrict__ partial) {
  pdl_prologue_done();
  const int e = seg_entry[blockIdx.x];
  const float* v = t.v[e];
  const long lo = static_cast<long>(seg_index[blockIdx.x]) * kSeg;
  const long hi = min(lo + kSeg, t.elems[e]);
  seg_reduce_store(seg_dot(v, v, lo, hi), partial);
}

__global__ void __launch_bounds__(256) wn_multi_scale_kernel(const WnMultiTable t, const int* __restrict__ blk_entry,
                                                             const int* __restrict__ blk_index,
                                                             const float* __restrict__ partial) {
  pdl_prologue_done();
  __shared__ float sh[9];
  const int e = blk_entry[blockIdx.x];
  const float n2 = group_total(partial + t.first_seg[e], t.n_seg[e], 0, 0, sh);     // the entry's partials, fixed order
  const long i = static_cast<long>(blk_index[blockIdx.x]) * 1024 + threadIdx.x * 4;
  if (i >= t.elems[e]) return;
  if (i == 0) *t.sumsq[e] = n2;
  const float s = *t.g[e] * rsqrtf(n2);
  const float4 f = *reinterpret_cast<const float4*>(t.v[e] + i);
  uint2 u;
  u.x = pack_bf16x2(f.x * s, f.y * s);
  u.y = pack_bf16x2(f.z * s, f.w * s);
  *reinterpret_cast<uint2*>(t.w[e] + i) = u;
}

// Backward of the fold for MANY layers in two launches (mirror of the two kernels above; same arithmetic and summation
// order as sumsq_kernel(v, dw) + wn_grad_kernel, so the results equal the per-layer backward bit for bit).
struct WnGradMultiTable {
  const float* const* dw;       // [entries] dW_eff of the entry (contiguous, elems floats)
  const float* const* v;        // [entries] weight_v
  const float* const* g;        // [entries] scalar weight_g
  const float* const* sumsq;    // [entries] ||V||_F^2 left by the forward fold
  float* const* dv;             // [entries] out: gradient of weight_v
  float* const* dg;             // [entries] out: gradient of weight_g
  const long* elems;
  const int* first_seg;
  const int* n_seg;
};

__global__ void __launch_bounds__(256) wn_multi_dot_kernel(const WnGradMultiTable t, const int* __restrict__ seg_entry,
                                                           const int* __restrict__ seg_index, float* __restrict__ partial) {
  pdl_prologue_done();
  const int e = seg_entry[blockIdx.x];
  const float* v = t.v[e];
  const float* dw = t.dw[e];
  const long lo = static_cast<long>(seg_index[blockIdx.x]) * kSeg;
  const long hi = min(lo + kSeg, t.elems[e]);
  seg_reduce_store(seg_dot(dw, v, lo, hi), partial);
}

__global__ void __launch_bounds__(256) wn_multi_grad_kernel(const WnGradMultiTable t, const int* __restrict__ blk_entry,
                                                            const int* __restrict__ blk_index,
                                                            const float* __restrict__ partial) {
  pdl_prologue_done();
  __shared__ float sh[9];
  const int e = blk_entry[blockIdx.x];
  const float dt = group_total(partial + t.first_seg[e], t.n_seg[e], 0, 0, sh);     // <dW_eff, V>, fixed order
  const long i = static_cast<long>(blk_index[blockIdx.x]) * 1024 + threadIdx.x * 4;
  if (i >= t.elems[e]) return;
  const float n2 = *t.sumsq[e];
  const float rn = rsqrtf(n2);
  const float s = *t.g[e] * rn;          // g / ||V||
  const float c = dt / n2;
  const float4 a = *reinterpret_cast<const float4*>(t.dw[e] + i);
  const float4 b = *reinterpret_cast<const float4*>(t.v[e] + i);
  float4 o;
  o.x = s * (a.x - c * b.x);
  o.y = s * (a.y - c * b.y);
  o.z = s * (a.z - c * b.z);
  o.w = s * (a.w - c * b.w);
  *reinterpret_cast<float4*>(t.dv[e] + i) = o;
  if (i == 0) *t.dg[e] = dt * rn;
}

// out[g, e] = sum_{j < rep} x[g * rep + j, e]  (bf16 in / out, fp32 sum): folds the per-row gradients of rows that share
// one v sample (v_rep of the contraction / pooling kernels) back onto that sample.  8 elements (16 bytes) per thread.
__global__ void __launch_bounds__(256)
sum_row_groups_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long groups, int rep, long row_vec) {
  pdl_prologue_done();
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= groups * row_vec) return;
  const long g = i / row_vec, e = i - g * row_vec;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < rep; ++j) {
    const uint4 u = __ldg(x + (g * rep + j) * row_vec + e);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16x2(w[t]);
      acc[2 * t] += f.x;
      acc[2 * t + 1] += f.y;
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  out[i] = o;
}

}  // namespace

int cast_rows_mask(const float* x, __nv_bfloat16* out, uint8_t* rowmask, long rows, int cols, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && cols > 0, "cast_rows_mask: bad shape rows=%ld cols=%d", rows, cols);
  if (rows == 0) return 0;
  CTI_REQUIRE(((cols & 7) != 0) || (((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0),
              "cast_rows_mask: buffers must be 16-byte aligned");
  const int warps = 8;
  const long blocks = (rows + warps - 1) / warps;
  CTI_REQUIRE(blocks < (1l << 31), "cast_rows_mask: too many rows");
  launch_pdl(cast_rows_mask_kernel, dim3((unsigned)blocks), dim3(warps * 32), 0, s, x, out, rowmask, rows, cols);
  return check_launch("cast_rows_mask_kernel");
}

int rowmask_bf16(const __nv_bfloat16* x, uint8_t* rowmask, long rows, int cols, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && cols > 0 && (cols & 7) == 0, "rowmask_bf16: cols=%d must be a multiple of 8", cols);
  if (rows == 0) return 0;
  CTI_REQUIRE(((uintptr_t)x & 15) == 0, "rowmask_bf16: features must be 16-byte aligned");
  launch_pdl(rowmask_bf16_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, s, x, rowmask, rows, cols);
  return check_launch("rowmask_bf16_kernel");
}

int cast_rows_dropout(const float* x, __nv_bfloat16* out, uint8_t* rowmask, long rows, int cols, float p, uint64_t seed,
                      uint64_t offset, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && cols > 0 && (cols & 7) == 0, "cast_rows_dropout: cols=%d must be a multiple of 8", cols);
  CTI_REQUIRE(p >= 0.f && p < 1.f, "cast_rows_dropout: p=%f outside [0, 1)", p);
  if (rows == 0) return 0;
  CTI_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "cast_rows_dropout: buffers must be 16-byte aligned");
  const long blocks = (rows + 7) / 8;
  CTI_REQUIRE(blocks < (1l << 31), "cast_rows_dropout: too many rows");
  launch_pdl(cast_rows_dropout_kernel, dim3((unsigned)blocks), dim3(256), 0, s, x, out, rowmask, rows, cols, make_rng(p, seed, offset));
  return check_launch("cast_rows_dropout_kernel");
}

int sum_row_groups(const __nv_bfloat16* x, __nv_bfloat16* out, long groups, int rep, long row_elems, cudaStream_t s) {
  CTI_REQUIRE(groups >= 0 && rep >= 1 && row_elems > 0 && (row_elems & 7) == 0, "sum_row_groups: bad arguments (row_elems=%ld)",
              row_elems);
  if (groups == 0) return 0;
  CTI_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "sum_row_groups: buffers must be 16-byte aligned");
  const long n = groups * (row_elems / 8);
  CTI_REQUIRE((n + 255) / 256 < (1l << 31), "sum_row_groups: too many elements");
  launch_pdl(sum_row_groups_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, reinterpret_cast<const uint4*>(x),
                                                                   reinterpret_cast<uint4*>(out), groups, rep, row_elems / 8);
  return check_launch("sum_row_groups_kernel");
}

int dropout_f32(float* x, long n, float p, uint64_t seed, uint64_t offset, cudaStream_t s) {
  CTI_REQUIRE(n >= 0 && (n & 3) == 0 && p >= 0.f && p < 1.f, "dropout_f32: bad arguments (n=%ld, p=%f)", n, p);
  if (n == 0) return 0;
  launch_pdl(dropout_f32_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, s, x, n / 4, make_rng(p, seed, offset));
  return check_launch("dropout_f32_kernel");
}

int dropout_bf16(const __nv_bfloat16* x, __nv_bfloat16* out, long n, float p, uint64_t seed, uint64_t offset,
                 cudaStream_t s) {
  CTI_REQUIRE(n >= 0 && (n & 3) == 0 && p >= 0.f && p < 1.f, "dropout_bf16: bad arguments (n=%ld, p=%f)", n, p);
  if (n == 0) return 0;
  launch_pdl(dropout_bf16_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, s, x, out, n / 4, make_rng(p, seed, offset));
  return check_launch("dropout_bf16_kernel");
}

int dropout_expand(const __nv_bfloat16* x, __nv_bfloat16* xt, long rows, int cols, int RG, int r0, float p, uint64_t seed,
                   uint64_t offset, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && cols > 0 && (cols & 3) == 0 && RG > 0 && r0 >= 0 && p >= 0.f && p < 1.f, "dropout_expand: bad arguments");
  if (rows == 0) return 0;
  const long n = rows * (cols / 4);
  launch_pdl(dropout_expand_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, x, xt, rows, cols / 4, RG, r0, make_rng(p, seed, offset));
  return check_launch("dropout_expand_kernel");
}

int dropout_reduce(const __nv_bfloat16* dxt, float* acc, long rows, int cols, int RG, int r0, float p, uint64_t seed,
                   uint64_t offset, cudaStream_t s) {
  CTI_REQUIRE(rows >= 0 && cols > 0 && (cols & 3) == 0 && RG > 0 && r0 >= 0 && p >= 0.f && p < 1.f, "dropout_reduce: bad arguments");
  if (rows == 0) return 0;
  const long n = rows * (cols / 4);
  launch_pdl(dropout_reduce_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dxt, acc, rows, cols / 4, RG, r0, make_rng(p, seed, offset));
  return check_launch("dropout_reduce_kernel");
}

int wn_pack_multi(const void* v_ptrs, const void* g_ptrs, const void* w_ptrs, const void* sumsq_ptrs, const long* elems,
                  const int* first_seg, const int* n_seg, const int* seg_entry, const int* seg_index, int n_segs,
                  const int* blk_entry, const int* blk_index, int n_blks, float* partials, cudaStream_t s) {
  CTI_REQUIRE(n_segs > 0 && n_blks > 0, "wn_pack_multi: empty table");
  WnMultiTable t;
  t.v = static_cast<const float* const*>(v_ptrs);
  t.g = static_cast<const float* const*>(g_ptrs);
  t.w = static_cast<__nv_bfloat16* const*>(w_ptrs);
  t.sumsq = static_cast<float* const*>(sumsq_ptrs);
  t.elems = elems; t.first_seg = first_seg; t.n_seg = n_seg;
  launch_pdl(wn_multi_sumsq_kernel, dim3(n_segs), dim3(256), 0, s, t, seg_entry, seg_index, partials);
  int rc = check_launch("wn_multi_sumsq_kernel");
  if (rc) return rc;
  launch_pdl(wn_multi_scale_kernel, dim3(n_blks), dim3(256), 0, s, t, blk_entry, blk_index, static_cast<const float*>(partials));
  return check_launch("wn_multi_scale_kernel");
}

int wn_grad_multi(const void* dw_ptrs, const void* v_ptrs, const void* g_ptrs, const void* sumsq_ptrs, const void* dv_ptrs,
                  const void* dg_ptrs, const long* elems, const int* first_seg, const int* n_seg, const int* seg_entry,
                  const int* seg_index, int n_segs, const int* blk_entry, const int* blk_index, int n_blks, float* partials,
                  cudaStream_t s) {
  CTI_REQUIRE(n_segs > 0 && n_blks > 0, "wn_grad_multi: empty table");
  WnGradMultiTable t;
  t.dw = static_cast<const float* const*>(dw_ptrs);
  t.v = static_cast<const float* const*>(v_ptrs);
  t.g = static_cast<const float* const*>(g_ptrs);
  t.sumsq = static_cast<const float* const*>(sumsq_ptrs);
  t.dv = static_cast<float* const*>(dv_ptrs);
  t.dg = static_cast<float* const*>(dg_ptrs);
  t.elems = elems; t.first_seg = first_seg; t.n_seg = n_seg;
  launch_pdl(wn_multi_dot_kernel, dim3(n_segs), dim3(256), 0, s, t, seg_entry, seg_index, partials);
  int rc = check_launch("wn_multi_dot_kernel");
  if (rc) return rc;
  launch_pdl(wn_multi_grad_kernel, dim3(n_blks), dim3(256), 0, s, t, blk_entry, blk_index, static_cast<const float*>(partials));
  return check_launch("wn_multi_grad_kernel");
}

size_t wn_scratch_floats(int n_groups, int rows_per_group, int cols) {
  const long ge = static_cast<long>(rows_per_group) * cols;
  return static_cast<size_t>(n_groups) * (1 + (ge + kSeg - 1) / kSeg);
}

// sumsq: wn_scratch_floats() floats -- [0, n_groups) the squared norms (output), the rest per-segment partial sums.
int wn_pack(const float* v, const float* g, __nv_bfloat16* w, float* sumsq, int n_groups, int rows_per_group, int cols,
            cudaStream_t s) {
  const long ge = static_cast<long>(rows_per_group) * cols;
  const long total = ge * n_groups;
  CTI_REQUIRE(n_groups > 0 && ge > 0, "wn_pack: empty weight");
  CTI_REQUIRE(ge % 4 == 0, "wn_pack: group size %ld must be a multiple of 4", ge);
  const int segs = (int)((ge + kSeg - 1) / kSeg);
  float* partial = sumsq + n_groups;
  launch_pdl(sumsq_kernel, dim3(n_groups * segs), dim3(256), 0, s, v, v, partial, ge, segs);
  int rc = check_launch("sumsq_kernel");
  if (rc) return rc;
  launch_pdl(wn_scale_kernel, dim3((unsigned)((total + 1023) / 1024)), dim3(256), 0, s, v, g, sumsq, partial, segs, w, ge, total);
  return check_launch("wn_scale_kernel");
}

// dot_ws: wn_scratch_floats() floats of scratch.
int wn_grad(const float* dw, const float* v, const float* g, const float* sumsq, float* dv, float* dg, float* dot_ws,
            int n_groups, int rows_per_group, int cols, cudaStream_t s) {
  const long ge = static_cast<long>(rows_per_group) * cols;
  const long total = ge * n_groups;
  CTI_REQUIRE(n_groups > 0 && ge > 0 && ge % 4 == 0, "wn_grad: bad group size %ld", ge);
  const int segs = (int)((ge + kSeg - 1) / kSeg);
  float* partial = dot_ws + n_groups;
  launch_pdl(sumsq_kernel, dim3(n_groups * segs), dim3(256), 0, s, dw, v, partial, ge, segs);
  int rc = check_launch("wn_dot_kernel");
  if (rc) return rc;
  launch_pdl(wn_grad_kernel, dim3((unsigned)((total + 1023) / 1024)), dim3(256), 0, s, dw, v, g, sumsq, partial, segs, dv, dg, ge, total);
  return check_launch("wn_grad_kernel");
}

int act_bwd_bias(const void* dy, int dy_is_bf16, const __nv_bfloat16* y, __nv_bfloat16* dz, float* dbias, long rows,
                 int cols, cudaStream_t s) {
  CTI_REQUIRE(rows > 0 && cols > 0 && (cols % 8) == 0, "act_bwd_bias: bad shape rows=%ld cols=%d (cols must be a multiple of 8)",
              rows, cols);
  CTI_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)dz & 15) == 0,
              "act_bwd_bias: buffers must be 16-byte aligned");
  const int gx = (cols + kAbbLanes * 8 - 1) / (kAbbLanes * 8);
  long rpb = (rows * gx + 4 * kNumSMsB200 - 1) / (4 * kNumSMsB200);       // ~4 waves of blocks
  rpb = rpb < 32 ? 32 : (rpb > 1024 ? 1024 : rpb);
  dim3 grid(gx, (unsigned)((rows + rpb - 1) / rpb));
  if (dy_is_bf16) launch_pdl(act_bwd_bias_kernel<true>, dim3(grid), dim3(256), 0, s, dy, y, dz, dbias, rows, cols, (int)rpb);
  else            launch_pdl(act_bwd_bias_kernel<false>, dim3(grid), dim3(256), 0, s, dy, y, dz, dbias, rows, cols, (int)rpb);
  return check_launch("act_bwd_bias_kernel");
}

}  // namespace cti
