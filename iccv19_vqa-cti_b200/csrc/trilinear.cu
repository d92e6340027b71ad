// Rank-R trilinear logit map of TCNet.forward (reference src/tc.py:46-52 with the three
// mode products of src/Tensor.py:6-19), forward and backward, for sm_100a.
//
//   L[b,k,q,a,g] = sum_r sum_{i,j,l} T_eff[r,i,j,l,g] Vc[b,k,r,i] Qc[b,q,r,j] Ac[b,a,r,l]
//
// contracted in the minimal-FLOP order a -> q -> v (T_min):
//   (i)   N1[a,(i,g),j]   = sum_l Ac[a,r,l]  T[r,l,(i,g,j)]
//   (ii)  M [a,i,(g,q)]   = sum_j N1[a,(i,g),j] Qc[q,r,j]
//   (iii) L [k,(a,g,q)]  += sum_i Vc[k,r,i]  M[a,i,(g,q)]
// Every stage runs on tensor cores (bf16 in, fp32 accumulate); N1 and M live only in shared
// memory, L accumulates over the rank loop in registers, so the reference's (B,K,Q,A,G)
// accumulator that is re-read and re-written 32 times never touches HBM and the
// K x Q x A x (R d) intermediate is never materialised.
//
// Layouts (d = 16 fixed):
//   Vc (B,K,R*16) bf16, Qc (B,Q,R*16), Ac (B,A,R*16)        -- outputs of the per-rank projections
//   tpack (R,16,16*G*16) bf16 = T_eff[r][l][(i,g,j)]         -- packed core (host gathers it from T_g)
//   logits (B,G,K,Q,A) fp32, -inf where rowmask[b,k] != 0    -- returned as a (B,K,Q,A,G) view
//
// Backward (appendix B of SURVEY.md): rank-outer persistent loop so that dT[r] accumulates in
// registers over all samples a CTA owns and is flushed with one atomic pass per (CTA, r).
#include "cti_common.cuh"
#include "cti_kernels.h"

#include "wmma_tiles.cuh"

namespace cti {

namespace {

using namespace tiles;

struct TriShape {
  int B, K, Q, A, G, R;
  int RD;      // R * 16
  int MT;      // ceil(K / 16)
  int NT;      // A * G      (16-wide column tiles of L, one per (a, g))
  int NTI;     // 16 * G     (16-wide column tiles of T[r], one per (i, g))
  int LDH;     // RD + 8
  int LDT;     // NTI * 16 + 8
  int LDM;     // G * 16 + 8
  int LDL;     // NT * 16 + 8   (backward: dLm pitch in smem)
};

__host__ __device__ inline TriShape make_shape(TriDims d) {
  TriShape s;
  s.B = d.B; s.K = d.K; s.Q = d.Q; s.A = d.A; s.G = d.G; s.R = d.R;
  s.RD = d.R * 16;
  s.MT = (d.K + 15) / 16;
  s.NT = d.A * d.G;
  s.NTI = 16 * d.G;
  s.LDH = s.RD + 8;
  s.LDT = s.NTI * 16 + 8;
  s.LDM = d.G * 16 + 8;
  s.LDL = s.NT * 16 + 8;
  return s;
}

// --------------------------------------------------------------------------- //
// forward
// --------------------------------------------------------------------------- //
struct FwdSmem {
  size_t off_v, off_q, off_a, off_t, off_n1, off_m, off_scr, total;
};

__host__ __device__ inline FwdSmem fwd_smem(const TriShape& s) {
  FwdSmem m;
  size_t o = 0;
  const size_t v_bytes = (size_t)s.MT * 16 * s.LDH * 2;
  const size_t out_bytes = (size_t)s.MT * 16 * (s.NT * 16 + 4) * 4;     // epilogue staging reuses this region
  m.off_v = o; o = align_up(o + (v_bytes > out_bytes ? v_bytes : out_bytes), 128);
  m.off_q = o; o = align_up(o + (size_t)16 * s.LDH * 2, 128);
  m.off_a = o; o = align_up(o + (size_t)16 * s.LDH * 2, 128);
  m.off_t = o; o = align_up(o + (size_t)2 * 16 * s.LDT * 2, 128);
  m.off_n1 = o; o = align_up(o + (size_t)s.A * 16 * s.G * kLdS * 2, 128);
  m.off_m = o; o = align_up(o + (size_t)s.A * 16 * s.LDM * 2, 128);
  m.off_scr = o; o = align_up(o + (size_t)kWarps * kScrFloats * 4, 128);
  m.total = o;
  return m;
}

__global__ void __launch_bounds__(kThreads, 1)
trilinear_fwd_kernel(const bf16* __restrict__ vc, const bf16* __restrict__ qc, const bf16* __restrict__ ac,
                     const bf16* __restrict__ tpack, const uint8_t* __restrict__ rowmask, float* __restrict__ logits,
                     const TriDims dims) {
  pdl_prologue_done();
  extern __shared__ __align__(128) uint8_t smem[];
  const TriShape s = make_shape(dims);
  const FwdSmem lay = fwd_smem(s);
  bf16* sV = reinterpret_cast<bf16*>(smem + lay.off_v);
  float* sOut = reinterpret_cast<float*>(smem + lay.off_v);
  bf16* sQ = reinterpret_cast<bf16*>(smem + lay.off_q);
  bf16* sA = reinterpret_cast<bf16*>(smem + lay.off_a);
  bf16* sT = reinterpret_cast<bf16*>(smem + lay.off_t);
  bf16* sN1 = reinterpret_cast<bf16*>(smem + lay.off_n1);
  bf16* sM = reinterpret_cast<bf16*>(smem + lay.off_m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;

  const int t_chunks = 16 * (s.NTI * 16) / 8;          // 16-byte chunks of one T[r] slab
  const int t_row_chunks = s.NTI * 16 / 8;
  auto prefetch_T = [&](int r, int buf) {
    const bf16* src = tpack + (size_t)r * 16 * s.NTI * 16;
    for (int c = threadIdx.x; c < t_chunks; c += kThreads) {
      const int row = c / t_row_chunks, col = (c - row * t_row_chunks) * 8;
      cp_async16(smem_u32(sT + ((size_t)buf * 16 + row) * s.LDT + col), src + (size_t)row * s.NTI * 16 + col);
    }
    cp_async_commit();
  };

  // Q / A operand rows beyond Q / A stay zero for the whole kernel.
  for (int i = threadIdx.x; i < 16 * s.LDH; i += kThreads) {
    sQ[i] = __float2bfloat16(0.f);
    sA[i] = __float2bfloat16(0.f);
  }

  int it = 0;
  prefetch_T(0, 0);
  const int row_chunks = s.RD / 8;
  const int tiles = s.MT * s.NT;
  const int out_ld = s.NT * 16 + 4;

  for (int b = blockIdx.x; b < s.B; b += gridDim.x) {
    __syncthreads();   // previous sample's epilogue finished with sOut (aliases sV)
    // ---- stage the sample's projected operands (zero-filled pad rows) ----
    for (int c = threadIdx.x; c < s.MT * 16 * row_chunks; c += kThreads) {
      const int row = c / row_chunks, col = (c - row * row_chunks) * 8;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (row < s.K) val = __ldg(reinterpret_cast<const uint4*>(vc + ((size_t)(b / dims.VR) * s.K + row) * s.RD + col));
      *reinterpret_cast<uint4*>(sV + (size_t)row * s.LDH + col) = val;
    }
    for (int c = threadIdx.x; c < s.Q * row_chunks; c += kThreads) {
      const int row = c / row_chunks, col = (c - row * row_chunks) * 8;
      *reinterpret_cast<uint4*>(sQ + (size_t)row * s.LDH + col) =
          __ldg(reinterpret_cast<const uint4*>(qc + ((size_t)b * s.Q + row) * s.RD + col));
    }
    for (int c = threadIdx.x; c < s.A * row_chunks; c += kThreads) {
      const int row = c / row_chunks, col = (c - row * row_chunks) * 8;
      *reinterpret_cast<uint4*>(sA + (size_t)row * s.LDH + col) =
          __ldg(reinterpret_cast<const uint4*>(ac + ((size_t)b * s.A + row) * s.RD + col));
    }
    FragC acc[kMaxAcc];
#pragma unroll
    for (int i = 0; i < kMaxAcc; ++i) wmma::fill_fragment(acc[i], 0.f);

    for (int r = 0; r < s.R; ++r, ++it) {
      cp_async_wait_all();
      __syncthreads();                                   // T[r] landed; operands visible; stage (iii) of r-1 done
      prefetch_T((r + 1) % s.R, (it + 1) & 1);
      const bf16* sTr = sT + (size_t)(it & 1) * 16 * s.LDT;

      // ---- (i) N1[a,(i,g),j] = Ac_r (16 x 16) * T_r (16 x 16*NTI) ----
      {
        FragAR fa;
        wmma::load_matrix_sync(fa, sA + r * 16, s.LDH);
        for (int t = warp; t < s.NTI; t += kWarps) {
          FragBR fb;
          wmma::load_matrix_sync(fb, sTr + t * 16, s.LDT);
          FragC c;
          wmma::fill_fragment(c, 0.f);
          wmma::mma_sync(c, fa, fb, c);
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);
          __syncwarp();
          const int i = t / s.G, g = t - i * s.G;
          for (int e = lane; e < s.A * 16; e += 32) {
            const int a = e >> 4, j = e & 15;
            sN1[((size_t)(a * 16 + i) * s.G + g) * kLdS + j] = __float2bfloat16(scr[a * kScrLd + j]);
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- (ii) M[a,i,(g,q)] = N1 (A*16*G x 16) * Qc_r^T (16 x 16) ----
      {
        FragBC fb;
        wmma::load_matrix_sync(fb, sQ + r * 16, s.LDH);
        for (int t = warp; t < s.NT; t += kWarps) {
          FragAR fa;
          wmma::load_matrix_sync(fa, sN1 + (size_t)t * 16 * kLdS, kLdS);
          FragC c;
          wmma::fill_fragment(c, 0.f);
          wmma::mma_sync(c, fa, fb, c);
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);
          __syncwarp();
          for (int e = lane; e < 256; e += 32) {
            const int row = e >> 4, q = e & 15;
            const int rho = t * 16 + row;
            const int a = rho / (16 * s.G);
            const int rem = rho - a * 16 * s.G;
            const int i = rem / s.G, g = rem - i * s.G;
            sM[((size_t)a * 16 + i) * s.LDM + g * 16 + q] = __float2bfloat16(scr[row * kScrLd + q]);
          }
          __syncwarp();
        }
      }
      __syncthreads();
      // ---- (iii) L[k,(a,g,q)] += Vc_r (K x 16) * M (16 x NT*16) ----
#pragma unroll
      for (int u = 0; u < kMaxAcc; ++u) {
        const int t = warp + u * kWarps;
        if (t < tiles) {
          const int mt = t % s.MT, nt = t / s.MT;
          const int a = nt / s.G, g = nt - a * s.G;
          FragAR fa;
          FragBR fb;
          wmma::load_matrix_sync(fa, sV + (size_t)mt * 16 * s.LDH + r * 16, s.LDH);
          wmma::load_matrix_sync(fb, sM + (size_t)a * 16 * s.LDM + g * 16, s.LDM);
          wmma::mma_sync(acc[u], fa, fb, acc[u]);
        }
      }
    }
    // ---- epilogue: stage the tile, then coalesced (B,G,K,Q,A) writes with the row mask ----
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kMaxAcc; ++u) {
      const int t = warp + u * kWarps;
      if (t < tiles) {
        const int mt = t % s.MT, nt = t / s.MT;
        wmma::store_matrix_sync(sOut + (size_t)mt * 16 * out_ld + nt * 16, acc[u], out_ld, wmma::mem_row_major);
      }
    }
    __syncthreads();
    const int QA = s.Q * s.A;
    const int per_g = s.K * QA;
    float* dst = logits + (size_t)b * s.G * per_g;
    for (int e = threadIdx.x; e < s.G * per_g; e += kThreads) {
      const int g = e / per_g;
      const int rem = e - g * per_g;
      const int k = rem / QA;
      const int qa = rem - k * QA;
      const int q = qa / s.A, a = qa - q * s.A;
      float val = sOut[(size_t)k * out_ld + (a * s.G + g) * 16 + q];
      if (rowmask != nullptr && rowmask[(size_t)(b / dims.VR) * s.K + k]) val = -INFINITY;
      dst[e] = val;
    }
  }
  cp_async_wait_all();
}

// --------------------------------------------------------------------------- //
// backward
// --------------------------------------------------------------------------- //
// dLm[b][k][(a,g,q16)] bf16, zero padded in q: the matrix form of dlogits (B,G,K,Q,A) fp32.
// One warp per (b, k) row: the G segments of Q*A contiguous floats are read coalesced into shared memory and the
// row of A*G*16 bf16 is written coalesced (the first version gathered 4-byte elements with stride A: 0.75 TB/s).
constexpr int kDlmWarps = 8;
__global__ void __launch_bounds__(kDlmWarps * 32) dlogits_to_dlm_kernel(const float* __restrict__ dlogits, bf16* __restrict__ dlm,
                                                                        const TriDims d) {
  pdl_prologue_done();
  __shared__ float seg[kDlmWarps][4 * 16 * 16];          // G <= 4, Q <= 16, A <= 16
  const TriShape s = make_shape(d);
  const int NL = s.NT * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qa = s.Q * s.A;
  const long rows = (long)s.B * s.K;
  float* sm = seg[warp];
  for (long row = (long)blockIdx.x * kDlmWarps + warp; row < rows; row += (long)gridDim.x * kDlmWarps) {
    const long b = row / s.K;
    const int k = (int)(row - b * s.K);
    for (int g = 0; g < s.G; ++g) {
      const float* src = dlogits + (((size_t)b * s.G + g) * s.K + k) * qa;
      for (int i = lane; i < qa; i += 32) sm[g * 256 + i] = src[i];
    }
    __syncwarp();
    bf16* dst = dlm + (size_t)row * NL;
    for (int n = lane; n < NL; n += 32) {
      const int q = n & 15, ag = n >> 4;
      const int a = ag / s.G, g = ag - a * s.G;
      const float v = (q < s.Q && a < s.A) ? sm[g * 256 + q * s.A + a] : 0.f;
      dst[n] = __float2bfloat16(v);
    }
    __syncwarp();
  }
}

struct BwdSmem {
  size_t off_t, off_dl, off_vr, off_qr, off_ar, off_n1, off_m, off_d, off_dn1, off_scr, off_dacc, off_db, total;
};
__host__ __device__ inline BwdSmem bwd_smem(const TriShape& s) {
  BwdSmem m;
  size_t o = 0;
  m.off_t = o; o = align_up(o + (size_t)16 * s.LDT * 2, 128);
  m.off_dl = o; o = align_up(o + (size_t)2 * s.MT * 16 * s.LDL * 2, 128);
  m.off_vr = o; o = align_up(o + (size_t)2 * s.MT * 16 * kLdS * 2, 128);
  m.off_qr = o; o = align_up(o + (size_t)2 * 16 * kLdS * 2, 128);
  m.off_ar = o; o = align_up(o + (size_t)2 * 16 * kLdS * 2, 128);
  m.off_n1 = o; o = align_up(o + (size_t)s.A * 16 * s.G * kLdS * 2, 128);
  m.off_m = o; o = align_up(o + (size_t)s.A * 16 * s.LDM * 2, 128);
  m.off_d = o; o = align_up(o + (size_t)s.A * 16 * s.G * kLdS * 2, 128);
  m.off_dn1 = o; o = align_up(o + (size_t)16 * s.LDT * 2, 128);
  m.off_scr = o; o = align_up(o + (size_t)kWarps * kScrFloats * 4, 128);
  m.off_dacc = o; o = align_up(o + (size_t)16 * 16 * 4, 128);
  m.off_db = o; o = align_up(o + (size_t)3 * 16 * 4, 128);
  m.total = o;
  return m;
}

__global__ void __launch_bounds__(kThreads, 1)
trilinear_bwd_kernel(const bf16* __restrict__ vc, const bf16* __restrict__ qc, const bf16* __restrict__ ac,
                     const bf16* __restrict__ tpack, const bf16* __restrict__ dlm, bf16* __restrict__ dzv,
                     bf16* __restrict__ dzq, bf16* __restrict__ dza, float* __restrict__ dbv, float* __restrict__ dbq,
                     float* __restrict__ dba, float* __restrict__ dtpack, const TriDims dims) {
  pdl_prologue_done();
  extern __shared__ __align__(128) uint8_t smem[];
  const TriShape s = make_shape(dims);
  const BwdSmem lay = bwd_smem(s);
  bf16* sT = reinterpret_cast<bf16*>(smem + lay.off_t);
  bf16* sDL = reinterpret_cast<bf16*>(smem + lay.off_dl);
  bf16* sVr = reinterpret_cast<bf16*>(smem + lay.off_vr);
  bf16* sQr = reinterpret_cast<bf16*>(smem + lay.off_qr);
  bf16* sAr = reinterpret_cast<bf16*>(smem + lay.off_ar);
  bf16* sN1 = reinterpret_cast<bf16*>(smem + lay.off_n1);
  bf16* sM = reinterpret_cast<bf16*>(smem + lay.off_m);
  bf16* sD = reinterpret_cast<bf16*>(smem + lay.off_d);
  bf16* sdN1 = reinterpret_cast<bf16*>(smem + lay.off_dn1);
  float* sdAcc = reinterpret_cast<float*>(smem + lay.off_dacc);
  float* sDb = reinterpret_cast<float*>(smem + lay.off_db);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scr = reinterpret_cast<float*>(smem + lay.off_scr) + warp * kScrFloats;

  // zero everything once: pad rows / pad columns of the operand buffers must read as 0.
  for (size_t i = threadIdx.x; i < lay.total / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();

  const int NL = s.NT * 16;
  const int dl_row_chunks = NL / 8;
  const int KP = s.MT * 16;
  const int n_my = (s.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // samples this CTA owns
  if (n_my <= 0) return;

  auto prefetch_sample = [&](int b, int r, int buf) {
    bf16* dl = sDL + (size_t)buf * KP * s.LDL;
    const bf16* src = dlm + (size_t)b * s.K * NL;
    for (int c = threadIdx.x; c < s.K * dl_row_chunks; c += kThreads) {
      const int row = c / dl_row_chunks, col = (c - row * dl_row_chunks) * 8;
      cp_async16(smem_u32(dl + (size_t)row * s.LDL + col), src + (size_t)row * NL + col);
    }
    bf16* v = sVr + (size_t)buf * KP * kLdS;
    for (int c = threadIdx.x; c < s.K * 2; c += kThreads) {
      const int row = c >> 1, col = (c & 1) * 8;
      cp_async16(smem_u32(v + (size_t)row * kLdS + col), vc + ((size_t)(b / dims.VR) * s.K + row) * s.RD + r * 16 + col);
    }
    bf16* q = sQr + (size_t)buf * 16 * kLdS;
    for (int c = threadIdx.x; c < s.Q * 2; c += kThreads) {
      const int row = c >> 1, col = (c & 1) * 8;
      cp_async16(smem_u32(q + (size_t)row * kLdS + col), qc + ((size_t)b * s.Q + row) * s.RD + r * 16 + col);
    }
    bf16* a = sAr + (size_t)buf * 16 * kLdS;
    for (int c = threadIdx.x; c < s.A * 2; c += kThreads) {
      const int row = c >> 1, col = (c & 1) * 8;
      cp_async16(smem_u32(a + (size_t)row * kLdS + col), ac + ((size_t)b * s.A + row) * s.RD + r * 16 + col);
    }
    cp_async_commit();
  };

  const int t_row_chunks = s.NTI * 16 / 8;
  int it = 0;
  prefetch_sample(blockIdx.x, 0, 0);

  for (int r = 0; r < s.R; ++r) {
    __syncthreads();                       // everyone finished with sT / sDb of rank r-1
    {
      const bf16* src = tpack + (size_t)r * 16 * s.NTI * 16;
      for (int c = threadIdx.x; c < 16 * t_row_chunks; c += kThreads) {
        const int row = c / t_row_chunks, col = (c - row * t_row_chunks) * 8;
        *reinterpret_cast<uint4*>(sT + (size_t)row * s.LDT + col) =
            __ldg(reinterpret_cast<const uint4*>(src + (size_t)row * s.NTI * 16 + col));
      }
      if (threadIdx.x < 48) sDb[threadIdx.x] = 0.f;
    }
    FragC accT[kMaxAcc];
#pragma unroll
    for (int i = 0; i < kMaxAcc; ++i) wmma::fill_fragment(accT[i], 0.f);

    for (int si = 0; si < n_my; ++si, ++it) {
      const int b = blockIdx.x + si * gridDim.x;
      const int buf = it & 1;
      cp_async_wait_all();
      __syncthreads();                     // (1) this sample's operands landed; previous iteration fully done
      {
        int nb = b + gridDim.x, nr = r;
        if (si + 1 == n_my) { nb = blockIdx.x; nr = r + 1; }
        if (nr < s.R) prefetch_sample(nb, nr, buf ^ 1);
      }
      const bf16* dl = sDL + (size_t)buf * KP * s.LDL;
      const bf16* vr = sVr + (size_t)buf * KP * kLdS;
      const bf16* qr = sQr + (size_t)buf * 16 * kLdS;
      const bf16* ar = sAr + (size_t)buf * 16 * kLdS;

      // ---- phase 1: N1 = Ac_r * T_r ----
      {
        FragAR fa;
        wmma::load_matrix_sync(fa, ar, kLdS);
        for (int t = warp; t < s.NTI; t += kWarps) {
          FragBR fb;
          wmma::load_matrix_sync(fb, sT + t * 16, s.LDT);
          FragC c;
          wmma::fill_fragment(c, 0.f);
          wmma::mma_sync(c, fa, fb, c);
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);
          __syncwarp();
          const int i = t / s.G, g = t - i * s.G;
          for (int e = lane; e < s.A * 16; e += 32) {
            const int a = e >> 4, j = e & 15;
            sN1[((size_t)(a * 16 + i) * s.G + g) * kLdS + j] = __float2bfloat16(scr[a * kScrLd + j]);
          }
          __syncwarp();
        }
      }
      __syncthreads();                     // (2)
      // ---- phase 2: M = N1 * Qc_r^T  (NT items)  and  D[i,n] = Vc_r^T * dLm  (NT items) ----
      for (int w = warp; w < 2 * s.NT; w += kWarps) {
        if (w < s.NT) {
          const int t = w;
          FragBC fb;
          FragAR fa;
          wmma::load_matrix_sync(fb, qr, kLdS);
          wmma::load_matrix_sync(fa, sN1 + (size_t)t * 16 * kLdS, kLdS);
          FragC c;
          wmma::fill_fragment(c, 0.f);
          wmma::mma_sync(c, fa, fb, c);
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);
          __syncwarp();
          for (int e = lane; e < 256; e += 32) {
            const int row = e >> 4, q = e & 15;
            const int rho = t * 16 + row;
            const int a = rho / (16 * s.G);
            const int rem = rho - a * 16 * s.G;
            const int i = rem / s.G, g = rem - i * s.G;
            sM[((size_t)a * 16 + i) * s.LDM + g * 16 + q] = __float2bfloat16(scr[row * kScrLd + q]);
          }
          __syncwarp();
        } else {
          const int nt = w - s.NT;               // (a, g) column tile of dLm
          const int a = nt / s.G, g = nt - a * s.G;
          FragC c;
          wmma::fill_fragment(c, 0.f);
          for (int ks = 0; ks < s.MT; ++ks) {
            FragAC fa;                           // (m = i, k = k) at vr[k][i]
            FragBR fb;
            wmma::load_matrix_sync(fa, vr + (size_t)ks * 16 * kLdS, kLdS);
            wmma::load_matrix_sync(fb, dl + (size_t)ks * 16 * s.LDL + nt * 16, s.LDL);
            wmma::mma_sync(c, fa, fb, c);
          }
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);   // scr[i][q]
          __syncwarp();
          for (int e = lane; e < 256; e += 32) {
            const int i = e >> 4, q = e & 15;
            sD[((size_t)(a * 16 + i) * s.G + g) * kLdS + q] = __float2bfloat16(scr[i * kScrLd + q]);
          }
          __syncwarp();
        }
      }
      __syncthreads();                     // (3)
      // ---- phase 3: dVc (MT items), dQc (1 item), dN1 (NT items) ----
      for (int w = warp; w < s.MT + 1 + s.NT; w += kWarps) {
        if (w < s.MT) {
          const int mt = w;
          FragC c;
          wmma::fill_fragment(c, 0.f);
          for (int nt = 0; nt < s.NT; ++nt) {
            const int a = nt / s.G, g = nt - a * s.G;
            FragAR fa;
            FragBC fb;                           // (k = q, n = i) at sM[a][i][g*16 + q]
            wmma::load_matrix_sync(fa, dl + (size_t)mt * 16 * s.LDL + nt * 16, s.LDL);
            wmma::load_matrix_sync(fb, sM + (size_t)a * 16 * s.LDM + g * 16, s.LDM);
            wmma::mma_sync(c, fa, fb, c);
          }
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);   // scr[k_local][i]
          __syncwarp();
          {
            const int i = lane & 15;
            float colsum = 0.f;
            for (int kk = lane >> 4; kk < 16; kk += 2) {
              const int k = mt * 16 + kk;
              if (k < s.K) {
                float v = scr[kk * kScrLd + i];
                if (!(__bfloat162float(vr[(size_t)k * kLdS + i]) > 0.f)) v = 0.f;
                dzv[((size_t)b * s.K + k) * s.RD + r * 16 + i] = __float2bfloat16(v);
                colsum += v;
              }
            }
            colsum += __shfl_xor_sync(0xffffffffu, colsum, 16);
            if (lane < 16) atomicAdd(sDb + i, colsum);
          }
          __syncwarp();
        } else if (w == s.MT) {
          FragC c;
          wmma::fill_fragment(c, 0.f);
          for (int ks = 0; ks < s.NT; ++ks) {
            FragAC fa;                           // (m = q, k = rho) at sD[rho][q]
            FragBR fb;
            wmma::load_matrix_sync(fa, sD + (size_t)ks * 16 * kLdS, kLdS);
            wmma::load_matrix_sync(fb, sN1 + (size_t)ks * 16 * kLdS, kLdS);
            wmma::mma_sync(c, fa, fb, c);
          }
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);   // scr[q][j]
          __syncwarp();
          {
            const int j = lane & 15;
            float colsum = 0.f;
            for (int q = lane >> 4; q < s.Q; q += 2) {
              float v = scr[q * kScrLd + j];
              if (!(__bfloat162float(qr[(size_t)q * kLdS + j]) > 0.f)) v = 0.f;
              dzq[((size_t)b * s.Q + q) * s.RD + r * 16 + j] = __float2bfloat16(v);
              colsum += v;
            }
            colsum += __shfl_xor_sync(0xffffffffu, colsum, 16);
            if (lane < 16) atomicAdd(sDb + 16 + j, colsum);
          }
          __syncwarp();
        } else {
          const int t = w - s.MT - 1;            // 16-row block of rho
          FragAR fa;
          FragBR fb;
          wmma::load_matrix_sync(fa, sD + (size_t)t * 16 * kLdS, kLdS);
          wmma::load_matrix_sync(fb, qr, kLdS);
          FragC c;
          wmma::fill_fragment(c, 0.f);
          wmma::mma_sync(c, fa, fb, c);
          wmma::store_matrix_sync(scr, c, kScrLd, wmma::mem_row_major);   // scr[rho_local][j]
          __syncwarp();
          for (int e = lane; e < 256; e += 32) {
            const int row = e >> 4, j = e & 15;
            const int rho = t * 16 + row;
            const int a = rho / (16 * s.G);
            const int rem = rho - a * 16 * s.G;      // = i*G + g
            sdN1[(size_t)a * s.LDT + rem * 16 + j] = __float2bfloat16(scr[row * kScrLd + j]);
          }
          __syncwarp();
        }
      }
      __syncthreads();                     // (4)
      // ---- phase 4: dT_r += Ac_r^T * dN1 (NTI tiles, persistent accumulators); dAc partial sums ----
      {
        FragAC faT;                              // (m = l, k = a) at ar[a][l]
        wmma::load_matrix_sync(faT, ar, kLdS);
        FragC cA;
        wmma::fill_fragment(cA, 0.f);
#pragma unroll
        for (int u = 0; u < kMaxAcc; ++u) {
          const int t = warp + u * kWarps;
          if (t < s.NTI) {
            FragBR fb;
            wmma::load_matrix_sync(fb, sdN1 + t * 16, s.LDT);
            wmma::mma_sync(accT[u], faT, fb, accT[u]);
            // dAc[a][l] += dN1[a][x-chunk t] * T_r[l][x-chunk t]
            FragAR fa;
            FragBC fbt;                          // (k = x, n = l) at sT[l][t*16 + x]
            wmma::load_matrix_sync(fa, sdN1 + t * 16, s.LDT);
            wmma::load_matrix_sync(fbt, sT + t * 16, s.LDT);
            wmma::mma_sync(cA, fa, fbt, cA);
          }
        }
        wmma::store_matrix_sync(scr, cA, kScrLd, wmma::mem_row_major);     // scr[a][l] partial
        __syncwarp();
        for (int e = lane; e < s.A * 16; e += 32) atomicAdd(sdAcc + e, scr[(e >> 4) * kScrLd + (e & 15)]);
        __syncwarp();
      }
      __syncthreads();                     // (5)
      if (warp == 0) {
        const int l = lane & 15;
        float colsum = 0.f;
        for (int a = lane >> 4; a < s.A; a += 2) {
          float v = sdAcc[a * 16 + l];
          if (!(__bfloat162float(ar[(size_t)a * kLdS + l]) > 0.f)) v = 0.f;
          dza[((size_t)b * s.A + a) * s.RD + r * 16 + l] = __float2bfloat16(v);
          colsum += v;
        }
        colsum += __shfl_xor_sync(0xffffffffu, colsum, 16);
        if (lane < 16) atomicAdd(sDb + 32 + l, colsum);
        __syncwarp();
        for (int e = lane; e < 256; e += 32) sdAcc[e] = 0.f;
      }
    }
    // ---- flush dT[r] and the bias gradients of rank r ----
#pragma unroll
    for (int u = 0; u < kMaxAcc; ++u) {
      const int t = warp + u * kWarps;
      if (t < s.NTI) {
        wmma::store_matrix_sync(scr, accT[u], kScrLd, wmma::mem_row_major);     // scr[l][x_local]
        __syncwarp();
        float* dst = dtpack + (size_t)r * 16 * s.NTI * 16 + t * 16;
        for (int e = lane; e < 256; e += 32) {
          const int l = e >> 4, x = e & 15;
          atomicAdd(dst + (size_t)l * s.NTI * 16 + x, scr[l * kScrLd + x]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (threadIdx.x < 16) atomicAdd(dbv + r * 16 + threadIdx.x, sDb[threadIdx.x]);
    else if (threadIdx.x < 32) atomicAdd(dbq + r * 16 + threadIdx.x - 16, sDb[threadIdx.x]);
    else if (threadIdx.x < 48) atomicAdd(dba + r * 16 + threadIdx.x - 32, sDb[threadIdx.x]);
  }
  cp_async_wait_all();
}

int check_dims(const TriDims& d, const char* who) {
  CTI_REQUIRE(d.B >= 0 && d.K > 0 && d.Q > 0 && d.A > 0 && d.G > 0 && d.R > 0, "%s: bad dims", who);
  CTI_REQUIRE(d.VR >= 1 && d.B % d.VR == 0, "%s: B=%d rows do not divide into groups of v_rep=%d", who, d.B, d.VR);
  CTI_REQUIRE(d.Q <= 16, "%s: at most 16 question tokens are supported (Q=%d)", who, d.Q);
  CTI_REQUIRE(d.A <= 16, "%s: at most 16 answer tokens are supported (A=%d)", who, d.A);
  CTI_REQUIRE(d.G <= 4, "%s: at most 4 glimpses are supported (G=%d)", who, d.G);
  const int tiles = ((d.K + 15) / 16) * d.A * d.G;
  CTI_REQUIRE(tiles <= kWarps * kMaxAcc, "%s: K*A*G too large for the on-chip accumulator (K=%d A=%d G=%d)", who, d.K,
              d.A, d.G);
  return 0;
}

}  // namespace

int trilinear_fwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack_perm, const uint8_t* rowmask,
                     float* logits, void* n1_save, TriDims d, cudaStream_t stream);   // trilinear_tc.cu

int trilinear_fwd(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const bf16* tpack_perm,
                  const uint8_t* rowmask, float* logits, void* n1_save, TriDims d, cudaStream_t stream) {
  if (int rc = check_dims(d, "trilinear_fwd")) return rc;
  if (d.B == 0) return 0;
  {   // tcgen05 fast path (G == 2, K <= 64, A <= 6, tpack_perm given); other shapes use the generic kernel below
    const int rc = trilinear_fwd_tc(vc, qc, ac, tpack_perm, rowmask, logits, n1_save, d, stream);
    if (rc != -100) return rc;
  }
  CTI_REQUIRE(n1_save == nullptr, "trilinear_fwd: n1_save is only written by the tcgen05 path (pass tpack_perm; "
                                  "cti_trilinear_n1_bytes() == 0 for shapes outside it)");
  const TriShape s = make_shape(d);
  const FwdSmem lay = fwd_smem(s);
  CTI_REQUIRE(lay.total <= 227 * 1024, "trilinear_fwd: needs %zu bytes of shared memory (> 227 KB)", lay.total);
  cudaError_t e = cudaFuncSetAttribute(trilinear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("trilinear_fwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(trilinear_fwd_kernel, dim3(grid), dim3(kThreads), lay.total, stream, vc, qc, ac, tpack, rowmask, logits, d);
  return check_launch("trilinear_fwd_kernel");
}

size_t trilinear_bwd_tc_workspace(TriDims d);                                   // trilinear_bwd_tc.cu
int trilinear_bwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const bf16* dlm, const void* n1,
                     bf16* dn1, bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dtpack,
                     TriDims d, cudaStream_t stream);

static size_t dlm_bytes(TriDims d) {
  const TriShape s = make_shape(d);
  return align_up((size_t)d.B * d.K * s.NT * 16 * sizeof(bf16), 1024);
}

size_t trilinear_bwd_workspace(TriDims d) {      // dLm (bf16 matrix form of dlogits) + dN1 of the tcgen05 path
  return dlm_bytes(d) + trilinear_bwd_tc_workspace(d);
}

int trilinear_bwd(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const float* dlogits,
                  const void* n1_saved, bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dtpack,
                  void* workspace, size_t workspace_bytes, TriDims d, cudaStream_t stream) {
  if (int rc = check_dims(d, "trilinear_bwd")) return rc;
  if (d.B == 0) return 0;
  CTI_REQUIRE(workspace != nullptr && workspace_bytes >= trilinear_bwd_workspace(d),
              "trilinear_bwd: workspace too small (%zu < %zu)", workspace_bytes, trilinear_bwd_workspace(d));
  const TriShape s = make_shape(d);
  const BwdSmem lay = bwd_smem(s);
  CTI_REQUIRE(lay.total <= 227 * 1024, "trilinear_bwd: needs %zu bytes of shared memory (> 227 KB)", lay.total);
  bf16* dlm = static_cast<bf16*>(workspace);
  launch_pdl(dlogits_to_dlm_kernel, dim3(kNumSMsB200 * 8), dim3(kDlmWarps * 32), 0, stream, dlogits, dlm, d);
  if (int rc = check_launch("dlogits_to_dlm_kernel")) return rc;
  {   // tcgen05 fast path (G == 2, A <= 6, K <= 64, N1 tiles saved by the forward); otherwise the generic kernel below
    bf16* dn1 = reinterpret_cast<bf16*>(static_cast<uint8_t*>(workspace) + dlm_bytes(d));
    const int rc = trilinear_bwd_tc(vc, qc, ac, tpack, dlm, n1_saved, dn1, dzv, dzq, dza, dbv, dbq, dba, dtpack, d, stream);
    if (rc != -100) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(trilinear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
  if (e != cudaSuccess) { set_error("trilinear_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
  const int grid = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(trilinear_bwd_kernel, dim3(grid), dim3(kThreads), lay.total, stream, vc, qc, ac, tpack, dlm, dzv, dzq, dza, dbv, dbq, dba,
                                                            dtpack, d);
  return check_launch("trilinear_bwd_kernel");
}

}  // namespace cti
