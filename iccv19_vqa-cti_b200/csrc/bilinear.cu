// Bilinear attention logits of BCNet.forward, `h_out <= 32` branch (reference src/bc.py:52-58,
// wrapped by BiAttention, src/attention.py:19-20,33):
//
//   logits[b,g,k,q] = sum_c Vb[b,k,c] h[g,c] Qb[b,q,c] + h_bias[g]
//
// The reference materialises h_ = v_ * h_mat as a (B,G,K,C) fp32 tensor (315 MB at B=256) and
// then runs a batched matmul.  Here h[g,:] is folded into the small operand on the fly:
//   HQ[(g,q), c] = h[g,c] Qb[q,c]        (built per 128-channel chunk in shared memory)
//   logits[k,(g,q)] = Vb (K x C) . HQ^T  (tensor cores, fp32 accumulate over all chunks)
// Backward (appendix B of SURVEY.md):
//   dVb[k,c] = sum_{g,q} dL[k,(g,q)] HQ[(g,q),c]
//   P[(g,q),c] = sum_k dL[k,(g,q)] Vb[k,c];  dQb[q,c] = sum_g h[g,c] P;  dh[g,c] += sum_q Qb[q,c] P
//   dbias[g] = sum_{k,q} dL
// Outputs are pre-activation gradients (ReLU masks of Vb / Qb applied) plus bias gradients.
#include "cti_common.cuh"
#include "cti_kernels.h"

#include "wmma_tiles.cuh"

namespace cti {

namespace {

using namespace tiles;

constexpr int kCC = 128;
constexpr int kLdC = kCC + 8;

struct BiShape {
  int B, K, Q, G, C;
  int MT;     // ceil(K / 16)
  int NCH;    // C / kCC
  int LDG;    // G * 16 + 8
};
__host__ __device__ inline BiShape make_bi_shape(BiDims d) {
  BiShape s;
  s.B = d.B; s.K = d.K; s.Q = d.Q; s.G = d.G; s.C = d.C;
  s.MT = (d.K + 15) / 16;
  s.NCH = d.C / kCC;
  s.LDG = d.G * 16 + 8;
  return s;
}

struct BiSmem {
  size_t off_v, off_hq, off_q, off_dl, off_scr, off_dh, off_db, total;
};
__host__ __device__ inline BiSmem bi_smem(const BiShape& s, bool bwd) {
  BiSmem m;
  size_t o = 0;
  m.off_v = o; o = align_up(o + (size_t)2 * s.MT * 16 * kLdC * 2, 128);
  m.off_hq = o; o = align_up(o + (size_t)2 * s.G * 16 * kLdC * 2, 128);
  m.off_q = o; if (bwd) o = align_up(o + (size_t)16 * kLdC * 2, 128);
  m.off_dl = o; if (bwd) o = align_up(o + (size_t)s.MT * 16 * s.LDG * 2, 128);
  m.off_scr = o; o = align_up(o + (size_t)kWarps * kScrFloats * 4, 128);
  m.off_dh = o; if (bwd) o = align_up(o + (size_t)s.G * s.C * 4, 128);
  m.off_db = o; if (bwd) o = align_up(o + (size_t)2 * s.C * 4 + 64, 128);
  m.total = o;
  return m;
}

__device__ __forceinline__ void prefetch_rows(bf16* dst, const bf16* src_sample, int rows, int C, int c0) {
  for (int c = threadIdx.x; c < rows * (kCC / 8); c += kThreads) {
    const int row = c / (kCC / 8), col = (c - row * (kCC / 8)) * 8;
    cp_async16(smem_u32(dst + (size_t)row * kLdC + col), src_sample + (size_t)row * C + c0 + col);
  }
}

// HQ[(g*16+q)][c_local] = h[g][c] * Qb[q][c] for one chunk; rows q >= Q stay zero.
__device__ __forceinline__ void build_hq(bf16* sHQ, const bf16* qb_sample, const float* hmat, const BiShape& s, int c0) {
  for (int e = threadIdx.x; e < s.G * s.Q * kCC; e += kThreads) {
    const int cl = e % kCC;
    const int gq = e / kCC;
    const int g = gq / s.Q, qi = gq - g * s.Q;
    const float val = __ldg(hmat + (size_t)g * s.C + c0 + cl) * __bfloat162float(qb_sample[(size_t)qi * s.C + c0 + cl]);
    sHQ[(size_t)(g * 16 + qi) * kLdC + cl] = __float2bfloat16(val);
  }
}

// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(kThreads)
bilinear_fwd_kernel(const bf16* __restrict__ vb, const bf16* __restrict__ qb, const float* __restrict__ hmat,
                    const float* __restrict__ hbias, const uint8_t* __restrict__ rowmask, float* __restrict__ logits,
                    const BiDims dims) {
  pdl_prologue_done();
  extern __shared__ __align__(128) uint8_t smem[];
  const BiShape s = make_bi_shape(dims);
  const BiSmem lay = bi_smem(s, false);
  bf16* sV = reinterpret_cast<bf16*>(smem + lay.off_v);
  bf16* sHQ = reinterpret_cast<bf16*>(smem + lay.off_hq);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;
  float* sOut = reinterpret_cast<float*>(smem + lay.off_v);          // epilogue staging aliases sV
  const int KP = s.MT * 16;
  const int tiles = s.MT * s.G;
  const int out_ld = s.G * 16 + 4;

  for (size_t i = threadIdx.x; i < lay.off_scr / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();

  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    const bf16* vs = vb + (size_t)b * s.K * s.C;
    const bf16* qs = qb + (size_t)b * s.Q * s.C;
    __syncthreads();                                     // previous epilogue done with sOut
    // pad rows of both V buffers were clobbered by the staging area: re-zero them
    for (int e = threadIdx.x; e < 2 * (KP - s.K) * kLdC; e += kThreads) {
      const int buf = e / ((KP - s.K) * kLdC);
      const int rem = e - buf * (KP - s.K) * kLdC;
      sV[((size_t)buf * KP + s.K) * kLdC + rem] = __float2bfloat16(0.f);
    }
    prefetch_rows(sV, vs, s.K, s.C, 0);
    cp_async_commit();
    build_hq(sHQ, qs, hmat, s, 0);
    FragC acc[kMaxAcc];
#pragma unroll
    for (int u = 0; u < kMaxAcc; ++u) wmma::fill_fragment(acc[u], 0.f);

    for (int ch = 0; ch < s.NCH; ++ch) {
      const int buf = ch & 1;
      cp_async_wait_all();
      __syncthreads();                                   // chunk ch staged (V via cp.async, HQ via st.shared)
      if (ch + 1 < s.NCH) {
        prefetch_rows(sV + (size_t)(buf ^ 1) * KP * kLdC, vs, s.K, s.C, (ch + 1) * kCC);
        cp_async_commit();
        build_hq(sHQ + (size_t)(buf ^ 1) * s.G * 16 * kLdC, qs, hmat, s, (ch + 1) * kCC);
      }
      const bf16* vbuf = sV + (size_t)buf * KP * kLdC;
      const bf16* hq = sHQ + (size_t)buf * s.G * 16 * kLdC;
#pragma unroll
      for (int u = 0; u < kMaxAcc; ++u) {
        const int t = warp + u * kWarps;
        if (t < tiles) {
          const int mt = t % s.MT, g = t / s.MT;
          for (int ks = 0; ks < kCC / 16; ++ks) {
            FragAR fa;                                   // (m = k, k = c)
            FragBC fb;                                   // (k = c, n = q) at hq[g*16 + q][c]
            wmma::load_matrix_sync(fa, vbuf + (size_t)mt * 16 * kLdC + ks * 16, kLdC);
            wmma::load_matrix_sync(fb, hq + (size_t)g * 16 * kLdC + ks * 16, kLdC);
            wmma::mma_sync(acc[u], fa, fb, acc[u]);
          }
        }
      }
    }
    __syncthreads();                                     // all warps done reading sV before it becomes sOut
#pragma unroll
    for (int u = 0; u < kMaxAcc; ++u) {
      const int t = warp + u * kWarps;
      if (t < tiles) {
        const int mt = t % s.MT, g = t / s.MT;
        wmma::store_matrix_sync(sOut + (size_t)mt * 16 * out_ld + g * 16, acc[u], out_ld, wmma::mem_row_major);
      }
    }
    __syncthreads();
    const int per_g = s.K * s.Q;
    float* dst = logits + (size_t)b * s.G * per_g;
    for (int e = threadIdx.x; e < s.G * per_g; e += kThreads) {
      const int g = e / per_g;
      const int rem = e - g * per_g;
      const int k = rem / s.Q, qi = rem - k * s.Q;
      float val = sOut[(size_t)k * out_ld + g * 16 + qi] + __ldg(hbias + g);
      if (rowmask != nullptr && rowmask[(size_t)b * s.K + k]) val = -INFINITY;
      dst[e] = val;
    }
  }
  (void)scr;
  cp_async_wait_all();
}

// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(kThreads, 1)
bilinear_bwd_kernel(const bf16* __restrict__ vb, const bf16* __restrict__ qb, const float* __restrict__ hmat,
                    const float* __restrict__ dlogits, bf16* __restrict__ dzv, bf16* __restrict__ dzq,
                    float* __restrict__ dbv, float* __restrict__ dbq, float* __restrict__ dhmat,
                    float* __restrict__ dhbias, const BiDims dims) {
  pdl_prologue_done();
  extern __shared__ __align__(128) uint8_t smem[];
  const BiShape s = make_bi_shape(dims);
  const BiSmem lay = bi_smem(s, true);
  bf16* sV = reinterpret_cast<bf16*>(smem + lay.off_v);
  bf16* sHQ = reinterpret_cast<bf16*>(smem + lay.off_hq);
  bf16* sQ = reinterpret_cast<bf16*>(smem + lay.off_q);         // [16][kLdC] Qb chunk
  bf16* sDL = reinterpret_cast<bf16*>(smem + lay.off_dl);       // [KP][LDG]  dL[k][(g,q16)]
  float* sDh = reinterpret_cast<float*>(smem + lay.off_dh);     // [G][C]
  float* sDb = reinterpret_cast<float*>(smem + lay.off_db);     // [2][C] + G bias slots
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;
  const int KP = s.MT * 16;

  for (size_t i = threadIdx.x; i < lay.total / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();

  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    const bf16* vs = vb + (size_t)b * s.K * s.C;
    const bf16* qs = qb + (size_t)b * s.Q * s.C;
    __syncthreads();
    // dL[k][(g, q16)] as a bf16 matrix; bias gradient = plain sums of dL.
    {
      const float* dl = dlogits + (size_t)b * s.G * s.K * s.Q;
      for (int g = 0; g < s.G; ++g) {
        float part = 0.f;
        for (int e = threadIdx.x; e < s.K * s.Q; e += kThreads) {
          const float val = __ldg(dl + (size_t)g * s.K * s.Q + e);
          const int k = e / s.Q, qi = e - k * s.Q;
          sDL[(size_t)k * s.LDG + g * 16 + qi] = __float2bfloat16(val);
          part += val;
        }
        part = warp_sum(part);
        if (lane == 0) atomicAdd(sDb + 2 * s.C + g, part);
      }
    }
    prefetch_rows(sV, vs, s.K, s.C, 0);
    cp_async_commit();

    for (int ch = 0; ch < s.NCH; ++ch) {
      const int buf = ch & 1;
      const int c0 = ch * kCC;
      __syncthreads();                                   // previous chunk's readers of sHQ / sQ are done
      build_hq(sHQ, qs, hmat, s, c0);
      for (int c = threadIdx.x; c < s.Q * (kCC / 8); c += kThreads) {
        const int row = c / (kCC / 8), col = (c - row * (kCC / 8)) * 8;
        *reinterpret_cast<uint4*>(sQ + (size_t)row * kLdC + col) =
            __ldg(reinterpret_cast<const uint4*>(qs + (size_t)row * s.C + c0 + col));
      }
      cp_async_wait_all();
      __syncthreads();
      if (ch + 1 < s.NCH) {
        prefetch_rows(sV + (size_t)(buf ^ 1) * KP * kLdC, vs, s.K, s.C, c0 + kCC);
        cp_async_commit();
      }
      const bf16* vbuf = sV + (size_t)buf * KP * kLdC;
      const int cl = warp * 16 + (lane & 15);
      const int c = c0 + cl;

      // ---- dVb[k, c] = dL[k,(g,q)] . HQ[(g,q), c] ----
      {
        float colsum = 0.f;
        for (int mt = 0; mt < s.MT; ++mt) {
          FragC cV;
          wmma::fill_fragment(cV, 0.f);
          for (int g = 0; g < s.G; ++g) {
            FragAR fa;                                   // (m = k, k = q) at sDL[k][g*16 + q]
            FragBR fb;                                   // (k = q, n = c) at sHQ[g*16 + q][c]
            wmma::load_matrix_sync(fa, sDL + (size_t)mt * 16 * s.LDG + g * 16, s.LDG);
            wmma::load_matrix_sync(fb, sHQ + (size_t)g * 16 * kLdC + warp * 16, kLdC);
            wmma::mma_sync(cV, fa, fb, cV);
          }
          wmma::store_matrix_sync(scr, cV, kScrLd, wmma::mem_row_major);     // scr[k_local][c_local]
          __syncwarp();
          for (int row = lane >> 4; row < 16; row += 2) {
            const int k = mt * 16 + row;
            if (k < s.K) {
              float gval = scr[row * kScrLd + (lane & 15)];
              if (!(__bfloat162float(vbuf[(size_t)k * kLdC + cl]) > 0.f)) gval = 0.f;
              dzv[((size_t)b * s.K + k) * s.C + c] = __float2bfloat16(gval);
              colsum += gval;
            }
          }
          __syncwarp();
        }
        colsum += __shfl_xor_sync(0xffffffffu, colsum, 16);
        if (lane < 16) sDb[c] += colsum;
      }
      // ---- P[(g,q), c] = dL^T . Vb ; dQb, dh ----
      {
        float dq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dq[i] = 0.f;
        for (int g = 0; g < s.G; ++g) {
          FragC cP;
          wmma::fill_fragment(cP, 0.f);
          for (int ks = 0; ks < s.MT; ++ks) {
            FragAC fa;                                   // (m = q, k = k) at sDL[k][g*16 + q]
            FragBR fb;                                   // (k = k, n = c)
            wmma::load_matrix_sync(fa, sDL + (size_t)ks * 16 * s.LDG + g * 16, s.LDG);
            wmma::load_matrix_sync(fb, vbuf + (size_t)ks * 16 * kLdC + warp * 16, kLdC);
            wmma::mma_sync(cP, fa, fb, cP);
          }
          wmma::store_matrix_sync(scr, cP, kScrLd, wmma::mem_row_major);     // scr[q][c_local]
          __syncwarp();
          const float hg = __ldg(hmat + (size_t)g * s.C + c);
          float dh = 0.f;
          for (int qi = lane >> 4; qi < s.Q; qi += 2) {
            const float pv = scr[qi * kScrLd + (lane & 15)];
            dh += __bfloat162float(sQ[(size_t)qi * kLdC + cl]) * pv;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i == qi) dq[i] += hg * pv;
          }
          dh += __shfl_xor_sync(0xffffffffu, dh, 16);
          if (lane < 16) sDh[(size_t)g * s.C + c] += dh;
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) dq[i] += __shfl_xor_sync(0xffffffffu, dq[i], 16);
        if (lane < 16) {
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < s.Q) {
              float gval = dq[i];
              if (!(__bfloat162float(sQ[(size_t)i * kLdC + cl]) > 0.f)) gval = 0.f;
              dzq[((size_t)b * s.Q + i) * s.C + c] = __float2bfloat16(gval);
              sq += gval;
            }
          }
          sDb[s.C + c] += sq;
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  for (int c = threadIdx.x; c < s.C; c += kThreads) {
    atomicAdd(dbv + c, sDb[c]);
    atomicAdd(dbq + c, sDb[s.C + c]);
    for (int g = 0; g < s.G; ++g) atomicAdd(dhmat + (size_t)g * s.C + c, sDh[(size_t)g * s.C + c]);
  }
  if (threadIdx.x < s.G) atomicAdd(dhbias + threadIdx.x, sDb[2 * s.C + threadIdx.x]);
}

int check_bi(const BiDims& d, const char* who) {
  CTI_REQUIRE(d.B >= 0 && d.K > 0 && d.Q > 0 && d.G > 0 && d.C > 0, "%s: bad dims", who);
  CTI_REQUIRE(d.C % kCC == 0, "%s: channel count %d must be a multiple of %d", who, d.C, kCC);
  CTI_REQUIRE(d.Q <= 16, "%s: at most 16 question tokens (Q=%d)", who, d.Q);
  CTI_REQUIRE(((d.K + 15) / 16) * d.G <= kWarps * kMaxAcc, "%s: K*G too large (K=%d G=%d)", who, d.K, d.G);
  CTI_REQUIRE(d.G <= 14, "%s: at most 14 glimpses (G=%d)", who, d.G);
  return 0;
}

}  // namespace

int bilinear_fwd_tc(const bf16* vb, const bf16* qb, const float* hmat, const float* hbias, const uint8_t* rowmask,
                    float* logits, BiDims d, cudaStream_t stream);                         // bilinear_tc.cu
int bilinear_bwd_tc(const bf16* vb, const bf16* qb, const float* hmat, const float* dlogits, bf16* dzv, bf16* dzq,
                    float* dbv, float* dbq, float* dhmat, float* dhbias, BiDims d, cudaStream_t stream);

int bilinear_fwd(const bf16* vb, const bf16* qb, const float* hmat, const float* hbias, const uint8_t* rowmask,
                 float* logits, BiDims d, cudaStream_t stream) {
  if (int rc = check_bi(d, "bilinear_fwd")) return rc;
  if (d.B == 0) return 0;
  {   // tcgen05 fast path (G <= 4, K <= 64, C <= 3072); other shapes use the generic tensor-core kernel below
    const int rc = bilinear_fwd_tc(vb, qb, hmat, hbias, rowmask, logits, d, stream);
    if (rc != -100) return rc;
  }
  const BiShape s = make_bi_shape(d);
  const BiSmem lay = bi_smem(s, false);
  const size_t out_bytes = (size_t)s.MT * 16 * (s.G * 16 + 4) * 4;
  CTI_REQUIRE(out_bytes <= lay.off_hq, "bilinear_fwd: staging area does not fit (G=%d)", d.G);
  CTI_REQUIRE(lay.total <= 227 * 1024, "bilinear_fwd: needs %zu bytes of shared memory", lay.total);
  cudaError_t e = cudaFuncSetAttribute(bilinear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("bilinear_fwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int per_sm = (lay.total <= 110 * 1024) ? 2 : 1;
  const int cap = kNumSMsB200 * per_sm;
  const int grid = d.B < cap ? d.B : cap;
  launch_pdl(bilinear_fwd_kernel, dim3(grid), dim3(kThreads), lay.total, stream, vb, qb, hmat, hbias, rowmask, logits, d);
  return check_launch("bilinear_fwd_kernel");
}

int bilinear_bwd(const bf16* vb, const bf16* qb, const float* hmat, const float* dlogits, bf16* dzv, bf16* dzq,
                 float* dbv, float* dbq, float* dhmat, float* dhbias, BiDims d, cudaStream_t stream) {
  if (int rc = check_bi(d, "bilinear_bwd")) return rc;
  if (d.B == 0) return 0;
  {
    const int rc = bilinear_bwd_tc(vb, qb, hmat, dlogits, dzv, dzq, dbv, dbq, dhmat, dhbias, d, stream);
    if (rc != -100) return rc;
  }
  const BiShape s = make_bi_shape(d);
  const BiSmem lay = bi_smem(s, true);
  CTI_REQUIRE(lay.total <= 227 * 1024, "bilinear_bwd: needs %zu bytes of shared memory", lay.total);
  cudaError_t e = cudaFuncSetAttribute(bilinear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("bilinear_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(bilinear_bwd_kernel, dim3(grid), dim3(kThreads), lay.total, stream, vb, qb, hmat, dlogits, dzv, dzq, dbv, dbq, dhmat, dhbias, d);
  return check_launch("bilinear_bwd_kernel");
}

}  // namespace cti
