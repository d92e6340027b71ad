// Attention-weighted trilinear / bilinear pooling to the joint embedding (sm_100a):
//   TCNet.forward_with_weights  (reference src/tc.py:54-61)
//       out[b,c] = sum_{k,q,a} V[b,k,c] w[b,k,q,a] Qp[b,q,c] Ap[b,a,c]
//   BCNet.forward_with_weights  (reference src/bc.py:70-74)  -- the A == 0 case, Ap == 1
//       out[b,c] = sum_{k,q}   V[b,k,c] w[b,k,q]   Qp[b,q,c]
//
// Forward, per sample and per 128-channel chunk, on tensor cores:
//   Z[(q,a), c] = sum_k w[k,(q,a)] V[k,c]          (QA x K) . (K x 128)
//   out[c]      = sum_{(q,a)} Qp[q,c] Ap[a,c] Z[(q,a),c]       (Khatri-Rao dot, in the epilogue)
// so the (B,C,K,Q,A) broadcast product of the reference's einsum never exists.
//
// Backward (appendix B of SURVEY.md), same tiling:
//   dQp[q,c] = do[c] sum_a Ap[a,c] Z[(q,a),c]     dAp[a,c] = do[c] sum_q Qp[q,c] Z[(q,a),c]
//   dV[k,c]  = sum_{(q,a)} w[k,(q,a)] KRd[(q,a),c]                  KRd = do * Qp * Ap
//   dw[k,(q,a)] = sum_c V[k,c] KRd[(q,a),c]
// The ReLU masks of the producing projections are applied here (V, Qp, Ap are post-ReLU), so
// the outputs are pre-activation gradients dz*, plus their bias gradients.
#include "cti_common.cuh"
#include "cti_kernels.h"

#include "wmma_tiles.cuh"

namespace cti {

namespace {

using namespace tiles;

constexpr int kCC = 128;            // channels per chunk (one 16-wide tile per warp)
constexpr int kLdC = kCC + 8;

struct PoolShape {
  int B, K, Q, A, C;
  int An;      // max(A, 1)
  int QA;      // Q * An
  int QAT;     // ceil(QA / 16)
  int MT;      // ceil(K / 16)
  int LDW;     // QAT * 16 + 8
  int NCH;     // C / kCC
};

__host__ __device__ inline PoolShape make_pool_shape(PoolDims d) {
  PoolShape s;
  s.B = d.B; s.K = d.K; s.Q = d.Q; s.A = d.A; s.C = d.C;
  s.An = d.A > 0 ? d.A : 1;
  s.QA = d.Q * s.An;
  s.QAT = (s.QA + 15) / 16;
  s.MT = (d.K + 15) / 16;
  s.LDW = s.QAT * 16 + 8;
  s.NCH = d.C / kCC;
  return s;
}

struct PoolSmem {
  size_t off_w, off_v, off_q, off_a, off_kr, off_scr, off_do, off_db, total;
};

__host__ __device__ inline PoolSmem pool_smem(const PoolShape& s, bool bwd) {
  PoolSmem m;
  size_t o = 0;
  m.off_w = o; o = align_up(o + (size_t)s.MT * 16 * s.LDW * 2, 128);
  m.off_v = o; o = align_up(o + (size_t)2 * s.MT * 16 * kLdC * 2, 128);
  m.off_q = o; o = align_up(o + (size_t)s.Q * s.C * 2, 128);
  m.off_a = o; o = align_up(o + (size_t)s.An * s.C * 2, 128);
  m.off_kr = o; if (bwd) o = align_up(o + (size_t)s.QAT * 16 * kLdC * 2, 128);
  m.off_scr = o; o = align_up(o + (size_t)kWarps * kScrFloats * 4, 128);
  m.off_do = o; if (bwd) o = align_up(o + (size_t)s.C * 4, 128);
  m.off_db = o; if (bwd) o = align_up(o + (size_t)3 * s.C * 4, 128);
  m.total = o;
  return m;
}

// Stage one 128-channel chunk of V (K rows) with cp.async; pad rows were zeroed once.
__device__ __forceinline__ void prefetch_v_chunk(bf16* dst, const bf16* v_sample, int K, int C, int c0) {
  for (int c = threadIdx.x; c < K * (kCC / 8); c += kThreads) {
    const int row = c / (kCC / 8), col = (c - row * (kCC / 8)) * 8;
    cp_async16(smem_u32(dst + (size_t)row * kLdC + col), v_sample + (size_t)row * C + c0 + col);
  }
  cp_async_commit();
}

// Load w (K x QA fp32) as the bf16 matrix sW[k][qa] (zero padded), Qp and Ap rows.
__device__ __forceinline__ void stage_sample(const PoolShape& s, bf16* sW, bf16* sQ, bf16* sA, const float* w_sample,
                                             const bf16* q_sample, const bf16* a_sample) {
  for (int e = threadIdx.x; e < s.K * s.QA; e += kThreads) {
    const int k = e / s.QA, qa = e - k * s.QA;
    sW[(size_t)k * s.LDW + qa] = __float2bfloat16(__ldg(w_sample + e));
  }
  const int rc = s.C / 8;
  for (int c = threadIdx.x; c < s.Q * rc; c += kThreads)
    reinterpret_cast<uint4*>(sQ)[c] = __ldg(reinterpret_cast<const uint4*>(q_sample) + c);
  if (s.A > 0) {
    for (int c = threadIdx.x; c < s.A * rc; c += kThreads)
      reinterpret_cast<uint4*>(sA)[c] = __ldg(reinterpret_cast<const uint4*>(a_sample) + c);
  }
}

// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(kThreads)
tri_pool_fwd_kernel(const bf16* __restrict__ v, const bf16* __restrict__ q, const bf16* __restrict__ a,
                    const float* __restrict__ w, long w_stride_b, float* __restrict__ out, const PoolDims dims) {
  extern __shared__ __align__(128) uint8_t smem[];
  const PoolShape s = make_pool_shape(dims);
  const PoolSmem lay = pool_smem(s, false);
  bf16* sW = reinterpret_cast<bf16*>(smem + lay.off_w);
  bf16* sV = reinterpret_cast<bf16*>(smem + lay.off_v);
  bf16* sQ = reinterpret_cast<bf16*>(smem + lay.off_q);
  bf16* sA = reinterpret_cast<bf16*>(smem + lay.off_a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;
  const int KP = s.MT * 16;

  for (size_t i = threadIdx.x; i < lay.off_q / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;   // sW, sV pads
  __syncthreads();

  int it = 0;
  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    const bf16* vb = v + (size_t)b * s.K * s.C;
    __syncthreads();                                   // previous sample done with sW / sQ / sA / sV
    stage_sample(s, sW, sQ, sA, w + (size_t)b * w_stride_b, q + (size_t)b * s.Q * s.C,
                 s.A > 0 ? a + (size_t)b * s.A * s.C : nullptr);
    prefetch_v_chunk(sV + (size_t)(it & 1) * KP * kLdC, vb, s.K, s.C, 0);
    for (int ch = 0; ch < s.NCH; ++ch, ++it) {
      cp_async_wait_all();
      __syncthreads();
      if (ch + 1 < s.NCH) prefetch_v_chunk(sV + (size_t)((it + 1) & 1) * KP * kLdC, vb, s.K, s.C, (ch + 1) * kCC);
      const bf16* vch = sV + (size_t)(it & 1) * KP * kLdC + warp * 16;
      FragC acc[kMaxAcc];
#pragma unroll
      for (int u = 0; u < kMaxAcc; ++u) wmma::fill_fragment(acc[u], 0.f);
      for (int ks = 0; ks < s.MT; ++ks) {
        FragBR fb;                                     // (k = k, n = c)
        wmma::load_matrix_sync(fb, vch + (size_t)ks * 16 * kLdC, kLdC);
#pragma unroll
        for (int u = 0; u < kMaxAcc; ++u) {
          if (u < s.QAT) {
            FragAC fa;                                 // (m = qa, k = k) at sW[k][qa]
            wmma::load_matrix_sync(fa, sW + (size_t)ks * 16 * s.LDW + u * 16, s.LDW);
            wmma::mma_sync(acc[u], fa, fb, acc[u]);
          }
        }
      }
      const int c = ch * kCC + warp * 16 + (lane & 15);
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < kMaxAcc; ++u) {
        if (u < s.QAT) {
          wmma::store_matrix_sync(scr, acc[u], kScrLd, wmma::mem_row_major);   // scr[qa_local][c_local]
          __syncwarp();
          for (int row = lane >> 4; row < 16; row += 2) {
            const int qa = u * 16 + row;
            if (qa < s.QA) {
              const int qi = qa / s.An, ai = qa - qi * s.An;
              float kr = __bfloat162float(sQ[(size_t)qi * s.C + c]);
              if (s.A > 0) kr *= __bfloat162float(sA[(size_t)ai * s.C + c]);
              part += kr * scr[row * kScrLd + (lane & 15)];
            }
          }
          __syncwarp();
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 16);
      if (lane < 16) out[(size_t)b * s.C + c] = part;
    }
  }
  cp_async_wait_all();
}

// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(kThreads, 1)
tri_pool_bwd_kernel(const bf16* __restrict__ v, const bf16* __restrict__ q, const bf16* __restrict__ a,
                    const float* __restrict__ w, long w_stride_b, const float* __restrict__ dout,
                    bf16* __restrict__ dzv, bf16* __restrict__ dzq, bf16* __restrict__ dza, float* __restrict__ dbv,
                    float* __restrict__ dbq, float* __restrict__ dba, float* __restrict__ dw, const PoolDims dims) {
  extern __shared__ __align__(128) uint8_t smem[];
  const PoolShape s = make_pool_shape(dims);
  const PoolSmem lay = pool_smem(s, true);
  bf16* sW = reinterpret_cast<bf16*>(smem + lay.off_w);
  bf16* sV = reinterpret_cast<bf16*>(smem + lay.off_v);
  bf16* sQ = reinterpret_cast<bf16*>(smem + lay.off_q);
  bf16* sA = reinterpret_cast<bf16*>(smem + lay.off_a);
  bf16* sKR = reinterpret_cast<bf16*>(smem + lay.off_kr);
  float* sDo = reinterpret_cast<float*>(smem + lay.off_do);
  float* sDb = reinterpret_cast<float*>(smem + lay.off_db);     // [3][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;
  const int KP = s.MT * 16;
  const int dw_tiles = s.MT * s.QAT;

  for (size_t i = threadIdx.x; i < lay.total / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();

  int it = 0;
  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    const bf16* vb = v + (size_t)b * s.K * s.C;
    __syncthreads();
    stage_sample(s, sW, sQ, sA, w + (size_t)b * w_stride_b, q + (size_t)b * s.Q * s.C,
                 s.A > 0 ? a + (size_t)b * s.A * s.C : nullptr);
    for (int c = threadIdx.x; c < s.C; c += kThreads) sDo[c] = __ldg(dout + (size_t)b * s.C + c);
    prefetch_v_chunk(sV + (size_t)(it & 1) * KP * kLdC, vb, s.K, s.C, 0);
    FragC accW[3];                                      // dw tiles owned by this warp (<= 3: MT*QAT <= 24)
#pragma unroll
    for (int u = 0; u < 3; ++u) wmma::fill_fragment(accW[u], 0.f);

    for (int ch = 0; ch < s.NCH; ++ch, ++it) {
      const int c0 = ch * kCC;
      __syncthreads();                                  // previous chunk's readers of sKR are done; sDo/sQ/sA visible
      // KRd[(q,a)][c] = do[c] Qp[q,c] Ap[a,c] for this chunk (pad rows stay zero)
      for (int e = threadIdx.x; e < s.QA * kCC; e += kThreads) {
        const int qa = e / kCC, cl = e - qa * kCC;
        const int qi = qa / s.An, ai = qa - qi * s.An;
        float kr = sDo[c0 + cl] * __bfloat162float(sQ[(size_t)qi * s.C + c0 + cl]);
        if (s.A > 0) kr *= __bfloat162float(sA[(size_t)ai * s.C + c0 + cl]);
        sKR[(size_t)qa * kLdC + cl] = __float2bfloat16(kr);
      }
      cp_async_wait_all();
      __syncthreads();
      if (ch + 1 < s.NCH) prefetch_v_chunk(sV + (size_t)((it + 1) & 1) * KP * kLdC, vb, s.K, s.C, c0 + kCC);
      const bf16* vbuf = sV + (size_t)(it & 1) * KP * kLdC;
      const int cl = warp * 16 + (lane & 15);
      const int c = c0 + cl;
      const float doc = sDo[c];

      // ---- (1) Z tiles -> dQp, dAp ----
      {
        FragC acc[kMaxAcc];
#pragma unroll
        for (int u = 0; u < kMaxAcc; ++u) wmma::fill_fragment(acc[u], 0.f);
        for (int ks = 0; ks < s.MT; ++ks) {
          FragBR fb;
          wmma::load_matrix_sync(fb, vbuf + (size_t)ks * 16 * kLdC + warp * 16, kLdC);
#pragma unroll
          for (int u = 0; u < kMaxAcc; ++u) {
            if (u < s.QAT) {
              FragAC fa;
              wmma::load_matrix_sync(fa, sW + (size_t)ks * 16 * s.LDW + u * 16, s.LDW);
              wmma::mma_sync(acc[u], fa, fb, acc[u]);
            }
          }
        }
        float dq[16], da[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { dq[i] = 0.f; da[i] = 0.f; }
#pragma unroll
        for (int u = 0; u < kMaxAcc; ++u) {
          if (u < s.QAT) {
            wmma::store_matrix_sync(scr, acc[u], kScrLd, wmma::mem_row_major);
            __syncwarp();
            for (int row = lane >> 4; row < 16; row += 2) {
              const int qa = u * 16 + row;
              if (qa < s.QA) {
                const int qi = qa / s.An, ai = qa - qi * s.An;
                const float z = scr[row * kScrLd + (lane & 15)];
                const float qv = __bfloat162float(sQ[(size_t)qi * s.C + c]);
                const float av = s.A > 0 ? __bfloat162float(sA[(size_t)ai * s.C + c]) : 1.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i == qi) dq[i] += av * z;
                  if (i == ai) da[i] += qv * z;
                }
              }
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          dq[i] += __shfl_xor_sync(0xffffffffu, dq[i], 16);
          da[i] += __shfl_xor_sync(0xffffffffu, da[i], 16);
        }
        if (lane < 16) {
          float sq = 0.f, sa = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < s.Q) {
              float g = doc * dq[i];
              if (!(__bfloat162float(sQ[(size_t)i * s.C + c]) > 0.f)) g = 0.f;
              dzq[((size_t)b * s.Q + i) * s.C + c] = __float2bfloat16(g);
              sq += g;
            }
            if (i < s.A) {
              float g = doc * da[i];
              if (!(__bfloat162float(sA[(size_t)i * s.C + c]) > 0.f)) g = 0.f;
              dza[((size_t)b * s.A + i) * s.C + c] = __float2bfloat16(g);
              sa += g;
            }
          }
          sDb[s.C + c] += sq;          // column c is owned by exactly one lane of one warp
          if (s.A > 0) sDb[2 * s.C + c] += sa;
        }
      }
      // ---- (2) dV[k, c] = sum_qa w[k,qa] KRd[qa,c]  (already scaled by do) ----
      {
        float colsum = 0.f;
        for (int mt = 0; mt < s.MT; ++mt) {
          FragC cU;
          wmma::fill_fragment(cU, 0.f);
          for (int ks = 0; ks < s.QAT; ++ks) {
            FragAR fa;                                  // (m = k, k = qa) at sW[k][qa]
            FragBR fb;                                  // (k = qa, n = c)
            wmma::load_matrix_sync(fa, sW + (size_t)mt * 16 * s.LDW + ks * 16, s.LDW);
            wmma::load_matrix_sync(fb, sKR + (size_t)ks * 16 * kLdC + warp * 16, kLdC);
            wmma::mma_sync(cU, fa, fb, cU);
          }
          wmma::store_matrix_sync(scr, cU, kScrLd, wmma::mem_row_major);     // scr[k_local][c_local]
          __syncwarp();
          for (int row = lane >> 4; row < 16; row += 2) {
            const int k = mt * 16 + row;
            if (k < s.K) {
              float g = scr[row * kScrLd + (lane & 15)];
              if (!(__bfloat162float(vbuf[(size_t)k * kLdC + cl]) > 0.f)) g = 0.f;
              dzv[((size_t)b * s.K + k) * s.C + c] = __float2bfloat16(g);
              colsum += g;
            }
          }
          __syncwarp();
        }
        colsum += __shfl_xor_sync(0xffffffffu, colsum, 16);
        if (lane < 16) sDb[c] += colsum;
      }
      // ---- (3) dw[k, qa] += sum_{c in chunk} V[k,c] KRd[qa,c] ----
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int t = warp + u * kWarps;
        if (t < dw_tiles) {
          const int mt = t % s.MT, nt = t / s.MT;
          for (int ks = 0; ks < kCC / 16; ++ks) {
            FragAR fa;                                  // (m = k, k = c)
            FragBC fb;                                  // (k = c, n = qa) at sKR[qa][c]
            wmma::load_matrix_sync(fa, vbuf + (size_t)mt * 16 * kLdC + ks * 16, kLdC);
            wmma::load_matrix_sync(fb, sKR + (size_t)nt * 16 * kLdC + ks * 16, kLdC);
            wmma::mma_sync(accW[u], fa, fb, accW[u]);
          }
        }
      }
    }
    // ---- write dw for this sample ----
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int t = warp + u * kWarps;
      if (t < dw_tiles) {
        const int mt = t % s.MT, nt = t / s.MT;
        wmma::store_matrix_sync(scr, accW[u], kScrLd, wmma::mem_row_major);   // scr[k_local][qa_local]
        __syncwarp();
        for (int e = lane; e < 256; e += 32) {
          const int k = mt * 16 + (e >> 4), qa = nt * 16 + (e & 15);
          if (k < s.K && qa < s.QA) dw[((size_t)b * s.K + k) * s.QA + qa] = scr[(e >> 4) * kScrLd + (e & 15)];
        }
        __syncwarp();
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  for (int c = threadIdx.x; c < s.C; c += kThreads) {
    atomicAdd(dbv + c, sDb[c]);
    atomicAdd(dbq + c, sDb[s.C + c]);
    if (s.A > 0) atomicAdd(dba + c, sDb[2 * s.C + c]);
  }
}

int check_pool(const PoolDims& d, const char* who) {
  CTI_REQUIRE(d.B >= 0 && d.K > 0 && d.Q > 0 && d.A >= 0 && d.C > 0, "%s: bad dims", who);
  CTI_REQUIRE(d.C % kCC == 0, "%s: channel count %d must be a multiple of %d", who, d.C, kCC);
  CTI_REQUIRE(d.Q <= 16 && d.A <= 16, "%s: Q and A must be <= 16 (Q=%d A=%d)", who, d.Q, d.A);
  const PoolShape s = make_pool_shape(d);
  CTI_REQUIRE(s.QAT <= kMaxAcc, "%s: Q*A = %d exceeds %d", who, s.QA, kMaxAcc * 16);
  CTI_REQUIRE(s.MT * s.QAT <= 3 * kWarps, "%s: K*Q*A too large (K=%d, Q*A=%d)", who, d.K, s.QA);
  return 0;
}

}  // namespace

int tri_pool_fwd(const bf16* v, const bf16* q, const bf16* a, const float* w, long w_stride_b, float* out, PoolDims d,
                 cudaStream_t stream) {
  if (int rc = check_pool(d, "tri_pool_fwd")) return rc;
  if (d.B == 0) return 0;
  const PoolShape s = make_pool_shape(d);
  const PoolSmem lay = pool_smem(s, false);
  CTI_REQUIRE(lay.total <= 227 * 1024, "tri_pool_fwd: needs %zu bytes of shared memory", lay.total);
  cudaError_t e = cudaFuncSetAttribute(tri_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("tri_pool_fwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int per_sm = (lay.total <= 110 * 1024) ? 2 : 1;
  const int cap = kNumSMsB200 * per_sm;
  const int grid = d.B < cap ? d.B : cap;
  tri_pool_fwd_kernel<<<grid, kThreads, lay.total, stream>>>(v, q, a, w, w_stride_b, out, d);
  return check_launch("tri_pool_fwd_kernel");
}

int tri_pool_bwd(const bf16* v, const bf16* q, const bf16* a, const float* w, long w_stride_b, const float* dout,
                 bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dw, PoolDims d,
                 cudaStream_t stream) {
  if (int rc = check_pool(d, "tri_pool_bwd")) return rc;
  if (d.B == 0) return 0;
  const PoolShape s = make_pool_shape(d);
  const PoolSmem lay = pool_smem(s, true);
  CTI_REQUIRE(lay.total <= 227 * 1024, "tri_pool_bwd: needs %zu bytes of shared memory", lay.total);
  cudaError_t e = cudaFuncSetAttribute(tri_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("tri_pool_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  tri_pool_bwd_kernel<<<grid, kThreads, lay.total, stream>>>(v, q, a, w, w_stride_b, dout, dzv, dzq, dza, dbv, dbq, dba,
                                                           dw, d);
  return check_launch("tri_pool_bwd_kernel");
}

}  // namespace cti
