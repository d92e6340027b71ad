// Backward of the rank-R trilinear logit map on tcgen05 (glimpse-2 fast path).
//
//   L[b,k,(a,g,q)] = sum_r sum_i Vc_r[k,i] M_r[i,(a,g,q)],   M_r = N1_r x_j Qc_r,   N1_r = T_r x_l Ac_r
//
// Two kernels, both built from per-sample tcgen05.mma stages with TMEM accumulators and the dual-use
// swizzled operand tiles of tc_tiles.cuh:
//
// 1. trilinear_bwd1_tc_kernel -- sample-outer, rank-inner (like the forward).  Per (sample, rank):
//      F1, F2           recompute N1_r and M_r (same stages as the forward)
//      B1  dVc[k,i]     = dL[k,n] . M_r[i,n]^T                      -> ReLU mask, dzv, dbv
//      B2  D^T[n,i]     = dL[k,n]^T . Vc_r[k,i]                     -> D tile [q][(a,g,i)]
//      B3  dQc[q,j]     = D[q,(a,g,i)] . N1_r[(a,g,i),j]            -> ReLU mask, dzq, dbq
//      B4  dN1[(a,g,i),j] = D[q,(a,g,i)]^T . Qc_r[q,j]              -> bf16, written to the workspace
//    dL (the gradient of the logits, bf16, [k][(a,g,q16)]) is loaded once per sample by TMA.
// 2. trilinear_bwd2_tc_kernel -- rank-outer: CTA (r, sample chunk) keeps T_r resident and streams the
//    dN1_r tiles of its samples (two samples per step):
//      B5  dAc[l,(s,a)] = T_r[l,(i,g,j)] . dN1[(s,a),(i,g,j)]^T     -> ReLU mask, dza, dba
//      B6  dT_r[(i,g,j),l] += dN1[(s,a),(i,g,j)]^T . Ac_r[(s,a),l]   accumulated in TMEM over all samples,
//                                                                   one atomic flush per CTA
//    The reduction over the batch that dT needs never leaves TMEM until the end of the CTA.
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

namespace cti {

namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// Column sums of a 16-value-per-lane tile over the 32 lanes of a warp with 16 shuffles: afterwards the
// lanes with (lane & 1) == 0 hold the total of value index (lane >> 1).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
  float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  return w1;      // value index = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0) = lane >> 1
}

// =========================================================================== //
// kernel 1
// =========================================================================== //
constexpr int kThreads1 = 704;     // 22 warps: TMA, 5 single-thread MMA issuers (1, 2, 3, 16, 17), 4 x 4 converter / epilogue warps
constexpr int T_BYTES = 16 * 1024, T_RING = 2;
constexpr int OP_V = 0, OP_Q = 8192, OP_A = 10240, OP_BYTES = 12288, OP_RING = 3;
constexpr int N1_BYTES = 32 * 1024;
constexpr int M_BYTES = 8 * 1024;
constexpr int DL_CHUNKS = 3, DL_BYTES = DL_CHUNKS * 8192;     // dL tile [64 k][192 n]
constexpr int D_BYTES = 6 * 1024;                             // D tile [16 q][192 (a,g,i)]
constexpr int DB_FLOATS = 512;                                // bias-gradient accumulators (R * 16 <= 512)
constexpr uint32_t TM_F1 = 0, TM_F2 = 128, TM_B1 = 192, TM_B2 = 224, TM_B3 = 288, TM_B4 = 320;   // all double buffered

constexpr int N1_RING = 8;     // one 32 KB N1 tile holds the 16-column blocks of 4 consecutive ranks: 2 tiles = 8 units in flight
enum { A_TFULL = 0, A_TEMPTY = 2, A_OPFULL = 4, A_OPEMPTY = 7, A_DLFULL = 10, A_DLEMPTY = 12, A_F1FULL = 14, A_F1EMPTY = 16,
       A_N1FULL = 18, A_N1EMPTY = A_N1FULL + N1_RING, A_F2FULL = A_N1EMPTY + N1_RING, A_F2EMPTY = A_F2FULL + 2,
       A_MFULL = A_F2EMPTY + 2, A_MEMPTY = A_MFULL + 2, A_B1FULL = A_MEMPTY + 2, A_B1EMPTY = A_B1FULL + 2,
       A_B2FULL = A_B1EMPTY + 2, A_B2EMPTY = A_B2FULL + 2, A_DFULL = A_B2EMPTY + 2, A_DEMPTY = A_DFULL + 2,
       A_B3FULL = A_DEMPTY + 2, A_B3EMPTY = A_B3FULL + 2, A_B4FULL = A_B3EMPTY + 2, A_B4EMPTY = A_B4FULL + 2,
       A_COUNT = A_B4EMPTY + 2 };

struct Bwd1Params {
  bf16 *dzv, *dzq, *dn1;
  float *dbv, *dbq;
  int B, K, Q, A, R, N;
  int VR;        // rows b share the v operand of row b / VR (dzv stays per row b)
};

constexpr size_t kSmem1 = (size_t)T_RING * T_BYTES + OP_RING * OP_BYTES + 2 * N1_BYTES + 2 * M_BYTES + 2 * DL_BYTES +
                          2 * D_BYTES + 2 * DB_FLOATS * 4 + A_COUNT * 8 + 16 + 1024;

__global__ void __launch_bounds__(kThreads1, 1)
trilinear_bwd1_tc_kernel(const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_v,
                         const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_dl, const Bwd1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // Tiles that are read as a 128-row A operand although they hold fewer rows (Vc 64, dL 64, D 16: the extra
  // accumulator lanes are never read back) come first, so the over-read stays inside this allocation.
  const uint32_t sT = base;
  const uint32_t sOp = sT + T_RING * T_BYTES;
  const uint32_t sDL = sOp + OP_RING * OP_BYTES;
  const uint32_t sD = sDL + 2 * DL_BYTES;
  const uint32_t sM = sD + 2 * D_BYTES;
  const uint32_t sN1 = sM + 2 * M_BYTES;
  const uint32_t sDb = sN1 + 2 * N1_BYTES;
  const uint32_t sBar = sDb + 2 * DB_FLOATS * 4;
  const uint32_t tmem_slot = sBar + A_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };
  float* db_acc = reinterpret_cast<float*>(smem_raw + (sDb - smem_u32(smem_raw)));     // [0,512) dbv, [512,1024) dbq

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 2 * DB_FLOATS; i += kThreads1) db_acc[i] = 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_t);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_dl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(A_TFULL + s), 1);
      mbar_init(bar(A_TEMPTY + s), 1);
      mbar_init(bar(A_DLFULL + s), 1);
      mbar_init(bar(A_DLEMPTY + s), 2);       // B1 and B2 issuers
      mbar_init(bar(A_F1FULL + s), 1);
      mbar_init(bar(A_F1EMPTY + s), 4);
      mbar_init(bar(A_F2FULL + s), 1);
      mbar_init(bar(A_F2EMPTY + s), 4);
      mbar_init(bar(A_MFULL + s), 4);
      mbar_init(bar(A_MEMPTY + s), 1);
      mbar_init(bar(A_B1FULL + s), 1);
      mbar_init(bar(A_B1EMPTY + s), 4);
      mbar_init(bar(A_B2FULL + s), 1);
      mbar_init(bar(A_B2EMPTY + s), 4);
      mbar_init(bar(A_DFULL + s), 4);
      mbar_init(bar(A_DEMPTY + s), 1);
      mbar_init(bar(A_B3FULL + s), 1);
      mbar_init(bar(A_B3EMPTY + s), 4);
      mbar_init(bar(A_B4FULL + s), 1);
      mbar_init(bar(A_B4EMPTY + s), 4);
    }
    for (int s = 0; s < N1_RING; ++s) {
      mbar_init(bar(A_N1FULL + s), 4);
      mbar_init(bar(A_N1EMPTY + s), 2);       // F2 and B3 issuers
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(bar(A_OPFULL + s), 1);
      mbar_init(bar(A_OPEMPTY + s), 10);      // F2 and B4 issuers + the 8 epilogue warps that read Vc / Qc (ReLU masks)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_prologue_done();      // everything above touched only this CTA's shared memory / TMEM

  const int n_my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int U = n_my * p.R;
  const int nt2 = (p.A + 3) >> 2;               // 128-row tiles over (a,g,i)
  const int ntn = (p.N + 127) >> 7;             // 128-row tiles over n = (a,g,q16)
  const int kn = p.N >> 4;                      // K steps over n / over (a,g,i)
  const int ktok = (p.K + 15) >> 4;             // K steps over tokens

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t tslot = 0, tph = 0, oslot = 0, oph = 0;
      int r = 0, sl = 0, b = blockIdx.x;
      auto load_dl = [&](int s_local, int bb) {
        const int slot = s_local & 1;
        mbar_wait(bar(A_DLEMPTY + slot), ((s_local >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(A_DLFULL + slot), DL_BYTES);
        for (int c = 0; c < DL_CHUNKS; ++c) tma_load_3d(&tmap_dl, bar(A_DLFULL + slot), sDL + slot * DL_BYTES + c * 8192, c * 64, 0, bb);
      };
      for (int u = 0; u < U; ++u) {
        if (u == 0) load_dl(0, b);
        if (r == p.R / 2 && sl + 1 < n_my) load_dl(sl + 1, b + gridDim.x);
        if ((r & 3) == 0) {
          mbar_wait(bar(A_OPEMPTY + oslot), oph ^ 1u);
          mbar_arrive_expect_tx(bar(A_OPFULL + oslot), OP_BYTES);
          const uint32_t dst = sOp + oslot * OP_BYTES;
          tma_load_3d(&tmap_v, bar(A_OPFULL + oslot), dst + OP_V, r * 16, 0, b / p.VR);
          tma_load_3d(&tmap_q, bar(A_OPFULL + oslot), dst + OP_Q, r * 16, 0, b);
          tma_load_3d(&tmap_a, bar(A_OPFULL + oslot), dst + OP_A, r * 16, 0, b);
          if (++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        }
        mbar_wait(bar(A_TEMPTY + tslot), tph ^ 1u);
        mbar_arrive_expect_tx(bar(A_TFULL + tslot), T_BYTES);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          tma_load_3d(&tmap_t, bar(A_TFULL + tslot), sT + tslot * T_BYTES + c * 2048, c * 64, r * 16, 0);
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) { r = 0; ++sl; b += gridDim.x; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ issuer: F1 ------------------------------------
    if (lane == 0) {
      const uint32_t id_f1 = make_idesc_rt(128, 16, 1, 0);
      uint32_t tslot = 0, tph = 0, oslot = 0, oph = 0;
      int r = 0;
      for (int u = 0; u < U; ++u) {
        mbar_wait(bar(A_TFULL + tslot), tph);
        if ((r & 3) == 0) mbar_wait(bar(A_OPFULL + oslot), oph);
        mbar_wait(bar(A_F1EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint64_t db = desc_kmajor(sOp + oslot * OP_BYTES + OP_A, r & 3);
        const uint64_t da = desc_mnmajor(sT + tslot * T_BYTES, 0, 2048);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          umma_bf16_ss(tmem_base + TM_F1 + (u & 1) * 64 + t * 16, da + (uint64_t)(2 * t * 2048 >> 4), db, id_f1, 0u);
        umma_commit(bar(A_F1FULL + (u & 1)));
        umma_commit(bar(A_TEMPTY + tslot));
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------ issuer: B2 (D^T = dL^T Vc) -- depends on TMA data only, runs ahead -------
    if (lane == 0) {
      const uint32_t id_b2 = make_idesc_rt(128, 16, 1, 1);
      uint32_t oslot = 0, oph = 0;
      int r = 0, sl = 0;
      for (int u = 0; u < U; ++u) {
        const uint32_t dl = sDL + (sl & 1) * DL_BYTES;
        if ((r & 3) == 0) mbar_wait(bar(A_OPFULL + oslot), oph);
        if (r == 0) mbar_wait(bar(A_DLFULL + (sl & 1)), (sl >> 1) & 1);
        mbar_wait(bar(A_B2EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint64_t db0 = desc_mnmajor(sOp + oslot * OP_BYTES + OP_V + (r & 3) * 32, 0, 0);
        for (int t = 0; t < ntn; ++t) {
          const uint64_t da0 = desc_mnmajor(dl + 2 * t * 8192, 0, 8192);
          for (int ks = 0; ks < ktok; ++ks)
            umma_bf16_ss(tmem_base + TM_B2 + (u & 1) * 32 + t * 16, da0 + (uint64_t)(ks * 2048 >> 4),
                         db0 + (uint64_t)(ks * 2048 >> 4), id_b2, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar(A_B2FULL + (u & 1)));
        if (r == p.R - 1) umma_commit(bar(A_DLEMPTY + (sl & 1)));
        if (++r == p.R) { r = 0; ++sl; }
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
      }
    }
  } else if (warp == 16) {
    // ------------------------------ issuer: B1 (dVc = dL M^T) ---------------------------------------------
    if (lane == 0) {
      const uint32_t id_b1 = make_idesc_rt(128, 16, 0, 0);
      int r = 0, sl = 0;
      for (int u = 0; u < U; ++u) {
        const uint32_t dl = sDL + (sl & 1) * DL_BYTES;
        if (r == 0) mbar_wait(bar(A_DLFULL + (sl & 1)), (sl >> 1) & 1);
        mbar_wait(bar(A_MFULL + (u & 1)), (u >> 1) & 1);
        mbar_wait(bar(A_B1EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint64_t da0 = desc_kmajor(dl, 0), db0 = desc_kmajor(sM + (u & 1) * M_BYTES, 0);
        for (int ks = 0; ks < kn; ++ks)
          umma_bf16_ss(tmem_base + TM_B1 + (u & 1) * 16, da0 + (uint64_t)(((ks >> 2) * 8192 + (ks & 3) * 32) >> 4),
                       db0 + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4), id_b1, ks > 0 ? 1u : 0u);
        umma_commit(bar(A_B1FULL + (u & 1)));
        umma_commit(bar(A_MEMPTY + (u & 1)));
        if (r == p.R - 1) umma_commit(bar(A_DLEMPTY + (sl & 1)));
        if (++r == p.R) { r = 0; ++sl; }
      }
    }
  } else if (warp == 3) {
    // ------------------------------ issuer: F2 (M = N1 Qc^T) ------------------------------------------------
    if (lane == 0) {
      const uint32_t id_f2 = make_idesc_rt(128, 16, 0, 0);
      uint32_t oslot = 0, oph = 0;
      int r = 0;
      for (int u = 0; u < U; ++u) {
        if ((r & 3) == 0) mbar_wait(bar(A_OPFULL + oslot), oph);
        mbar_wait(bar(A_N1FULL + (u & 7)), (u >> 3) & 1);
        mbar_wait(bar(A_F2EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint64_t db = desc_kmajor(sOp + oslot * OP_BYTES + OP_Q, r & 3);
        const uint64_t da = desc_kmajor(sN1 + ((u >> 2) & 1) * N1_BYTES, u & 3);     // K step = rank within the quad tile
        for (int t2 = 0; t2 < nt2; ++t2)
          umma_bf16_ss(tmem_base + TM_F2 + (u & 1) * 32 + t2 * 16, da + (uint64_t)(t2 * 16384 >> 4), db, id_f2, 0u);
        umma_commit(bar(A_F2FULL + (u & 1)));
        umma_commit(bar(A_N1EMPTY + (u & 7)));
        if ((r & 3) == 3) umma_commit(bar(A_OPEMPTY + oslot));
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
      }
    }
  } else if (warp == 17) {
    // ------------------------------ issuer: B3 (dQc = D N1), B4 (dN1 = D^T Qc) ---------------------------------
    if (lane == 0) {
      const uint32_t id_b3 = make_idesc_rt(128, 16, 0, 1);
      const uint32_t id_b4 = make_idesc_rt(128, 16, 1, 1);
      uint32_t oslot = 0, oph = 0;
      int r = 0;
      for (int u = 0; u < U; ++u) {
        const uint32_t op = sOp + oslot * OP_BYTES;
        if ((r & 3) == 0) mbar_wait(bar(A_OPFULL + oslot), oph);
        mbar_wait(bar(A_N1FULL + (u & 7)), (u >> 3) & 1);
        mbar_wait(bar(A_DFULL + (u & 1)), (u >> 1) & 1);
        mbar_wait(bar(A_B3EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        mbar_wait(bar(A_B4EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t dt = sD + (u & 1) * D_BYTES;
        const uint64_t da0 = desc_kmajor(dt, 0), db0 = desc_mnmajor(sN1 + ((u >> 2) & 1) * N1_BYTES + (u & 3) * 32, 0, 0);
        for (int ks = 0; ks < kn; ++ks)
          umma_bf16_ss(tmem_base + TM_B3 + (u & 1) * 16, da0 + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4),
                       db0 + (uint64_t)((ks * 2048) >> 4), id_b3, ks > 0 ? 1u : 0u);
        umma_commit(bar(A_B3FULL + (u & 1)));
        const uint64_t dq = desc_mnmajor(op + OP_Q + (r & 3) * 32, 0, 0);
        for (int t = 0; t < nt2; ++t)
          umma_bf16_ss(tmem_base + TM_B4 + (u & 1) * 32 + t * 16, desc_mnmajor(dt + 2 * t * 2048, 0, 2048), dq, id_b4, 0u);
        umma_commit(bar(A_B4FULL + (u & 1)));
        umma_commit(bar(A_DEMPTY + (u & 1)));
        umma_commit(bar(A_N1EMPTY + (u & 7)));
        if ((r & 3) == 3) umma_commit(bar(A_OPEMPTY + oslot));
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
      }
    }
  } else if (warp < 8) {
    // ------------------------------ G1: N1^T (TMEM) -> N1 tile rows (a,g,i), columns j --------------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int j = L & 15, g = (L >> 4) & 1;
    for (int u = 0; u < U; ++u) {
      const int slot = u & 1;
      mbar_wait(bar(A_F1FULL + slot), (u >> 1) & 1);
      tcgen05_fence_after();
      mbar_wait(bar(A_N1EMPTY + (u & 7)), ((u >> 3) & 1) ^ 1);
      const uint32_t n1 = sN1 + ((u >> 2) & 1) * N1_BYTES;
      const uint32_t sub = u & 3;                                // rank within the quad tile: columns sub * 16 + j
      uint32_t v[4][8];
#pragma unroll
      for (int t = 0; t < 4; ++t) tmem_ld_32x32b_x8(tmem_base + lane_addr + TM_F1 + slot * 64 + t * 16, v[t]);
      tmem_wait_ld();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = 4 * t + (L >> 5);
        const uint32_t off = (g * 2 + (i >> 3)) * 1024u + (i & 7) * 128u + (((sub * 2 + ((j >> 3) & 1)) ^ (i & 7)) << 4) + (j & 7) * 2u;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          if (a < p.A) {
            const __nv_bfloat16 h = __float2bfloat16(__uint_as_float(v[t][a]));
            st_shared_u16(n1 + a * 4096u + off, *reinterpret_cast<const uint16_t*>(&h));
          }
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(A_N1FULL + (u & 7)));
        mbar_arrive(bar(A_F1EMPTY + slot));
      }
    }
  } else if (warp < 12) {
    // ------------------------------ G2: M (TMEM) -> M tile [i][(a,g,q16)].  (The D^T -> D tile conversion used to
    // live here too and made this group the busiest stage of the pipeline; it now runs on the E1 warps, which
    // otherwise wait for B1 most of the time.) -------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    for (int u = 0; u < U; ++u) {
      const int slot = u & 1;
      {   // C2
        mbar_wait(bar(A_F2FULL + slot), (u >> 1) & 1);
        tcgen05_fence_after();
        mbar_wait(bar(A_MEMPTY + slot), ((u >> 1) & 1) ^ 1);
        const uint32_t mt = sM + slot * M_BYTES;
        for (int t2 = 0; t2 < nt2; ++t2) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + slot * 32 + t2 * 16, v);
          tmem_wait_ld();
          const int rho = t2 * 128 + L;
          const int a = rho >> 5, i = rho & 15;
          if (a < p.A) {
            const int ag = rho >> 4;
            uint32_t pk[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
            const uint32_t tile = mt + (ag >> 2) * 2048u;
            const uint32_t c0 = (ag & 3) * 16;
            st_shared_v4(tile + sw128_off(i, c0), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(tile + sw128_off(i, c0 + 8), pk[4], pk[5], pk[6], pk[7]);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(A_MFULL + slot));
          mbar_arrive(bar(A_F2EMPTY + slot));
        }
      }
    }
  } else if (warp < 16 || warp >= 18) {
    // warps 12-15: E1 (dVc -> dzv); warps 18-21: E3 (dQc -> dzq) and E4 (dN1 -> workspace)
    const bool do_e1 = warp < 16;
    // ------------------------------ G3 / G4: epilogues  dVc -> dzv, dQc -> dzq, dN1 -> workspace -----------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int RD = p.R * 16;
    uint32_t oslot = 0, oph = 0;
    int r = 0, b = blockIdx.x;
    for (int u = 0; u < U; ++u) {
      const int slot = u & 1;
      const uint32_t op = sOp + oslot * OP_BYTES;
      if ((r & 3) == 0) mbar_wait(bar(A_OPFULL + oslot), oph);
      // activation chunk (16 values of rank r) of row `row` of an OP tile, de-swizzled
      auto act16 = [&](uint32_t tile, int row, float (&out)[16]) {
        uint32_t w0[4], w1[4];
        const uint32_t rb = tile + (row >> 3) * 1024u + (row & 7) * 128u;
        ld_shared_v4(rb + ((((r & 3) * 2) ^ (row & 7)) << 4), w0);
        ld_shared_v4(rb + ((((r & 3) * 2 + 1) ^ (row & 7)) << 4), w1);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float2 f0 = unpack_bf16x2(w0[x]), f1 = unpack_bf16x2(w1[x]);
          out[2 * x] = f0.x; out[2 * x + 1] = f0.y; out[8 + 2 * x] = f1.x; out[8 + 2 * x + 1] = f1.y;
        }
      };
      if (do_e1) {   // C3: D^T (TMEM) -> D tile [q][(a,g,i)]; B2 runs ahead of the chain, so this is ready early
        mbar_wait(bar(A_B2FULL + slot), (u >> 1) & 1);
        tcgen05_fence_after();
        mbar_wait(bar(A_DEMPTY + slot), ((u >> 1) & 1) ^ 1);
        const uint32_t dt = sD + slot * D_BYTES;
        for (int t = 0; t < ntn; ++t) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B2 + slot * 32 + t * 16, v);
          tmem_wait_ld();
          const int n = t * 128 + L;
          if (n < p.N) {
            const int ag = n >> 4, q = n & 15;
            uint32_t pk[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
            const uint32_t tile = dt + (ag >> 2) * 2048u;
            const uint32_t c0 = (ag & 3) * 16;
            st_shared_v4(tile + sw128_off(q, c0), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(tile + sw128_off(q, c0 + 8), pk[4], pk[5], pk[6], pk[7]);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(A_DFULL + slot));
          mbar_arrive(bar(A_B2EMPTY + slot));
        }
      }
      if (do_e1) {   // E1: dVc[k, i]
        mbar_wait(bar(A_B1FULL + slot), (u >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B1 + slot * 16, v);
        tmem_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(A_B1EMPTY + slot));
        float gv[16];
        if (L < p.K) {
          float act[16];
          act16(op + OP_V, L, act);
#pragma unroll
          for (int x = 0; x < 16; ++x) gv[x] = act[x] > 0.f ? __uint_as_float(v[x]) : 0.f;
          uint4 o0, o1;
          o0.x = pack_bf16x2(gv[0], gv[1]); o0.y = pack_bf16x2(gv[2], gv[3]); o0.z = pack_bf16x2(gv[4], gv[5]); o0.w = pack_bf16x2(gv[6], gv[7]);
          o1.x = pack_bf16x2(gv[8], gv[9]); o1.y = pack_bf16x2(gv[10], gv[11]); o1.z = pack_bf16x2(gv[12], gv[13]); o1.w = pack_bf16x2(gv[14], gv[15]);
          uint4* dst = reinterpret_cast<uint4*>(p.dzv + ((size_t)b * p.K + L) * RD + r * 16);
          dst[0] = o0;
          dst[1] = o1;
        } else {
#pragma unroll
          for (int x = 0; x < 16; ++x) gv[x] = 0.f;
        }
        if (qd * 32 < p.K) {          // warp-uniform: this warp holds valid regions
          const float s = warp_colsum16(gv, lane);
          if ((lane & 1) == 0) atomicAdd(db_acc + r * 16 + (lane >> 1), s);
        }
      }
      if (!do_e1) {   // E3: dQc[q, j]
        mbar_wait(bar(A_B3FULL + slot), (u >> 1) & 1);
        tcgen05_fence_after();
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B3 + slot * 16, v);
        tmem_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(A_B3EMPTY + slot));
        if (qd == 0) {
          float gv[16];
          if (L < p.Q) {
            float act[16];
            act16(op + OP_Q, L, act);
#pragma unroll
            for (int x = 0; x < 16; ++x) gv[x] = act[x] > 0.f ? __uint_as_float(v[x]) : 0.f;
            uint4 o0, o1;
            o0.x = pack_bf16x2(gv[0], gv[1]); o0.y = pack_bf16x2(gv[2], gv[3]); o0.z = pack_bf16x2(gv[4], gv[5]); o0.w = pack_bf16x2(gv[6], gv[7]);
            o1.x = pack_bf16x2(gv[8], gv[9]); o1.y = pack_bf16x2(gv[10], gv[11]); o1.z = pack_bf16x2(gv[12], gv[13]); o1.w = pack_bf16x2(gv[14], gv[15]);
            uint4* dst = reinterpret_cast<uint4*>(p.dzq + ((size_t)b * p.Q + L) * RD + r * 16);
            dst[0] = o0;
            dst[1] = o1;
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x) gv[x] = 0.f;
          }
          const float s = warp_colsum16(gv, lane);
          if ((lane & 1) == 0) atomicAdd(db_acc + DB_FLOATS + r * 16 + (lane >> 1), s);
        }
      }
      if (!do_e1) {   // E4: dN1[(a,g,i), j] -> workspace [b][r][a][(i,g,j)] bf16
        mbar_wait(bar(A_B4FULL + slot), (u >> 1) & 1);
        tcgen05_fence_after();
        for (int t = 0; t < nt2; ++t) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B4 + slot * 32 + t * 16, v);
          tmem_wait_ld();
          const int rho = t * 128 + L;
          const int a = rho >> 5, g = (rho >> 4) & 1, i = rho & 15;
          if (a < p.A) {
            uint4 o0, o1;
            o0.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));
            o0.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
            o0.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));
            o0.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
            o1.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));
            o1.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
            o1.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13]));
            o1.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
            uint4* dst = reinterpret_cast<uint4*>(p.dn1 + (((size_t)b * p.R + r) * p.A + a) * 512 + i * 32 + g * 16);
            dst[0] = o0;
            dst[1] = o1;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(A_B4EMPTY + slot));
      }
      if ((r & 3) == 3) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(A_OPEMPTY + oslot));
      }
      if (++r == p.R) { r = 0; b += gridDim.x; }
      if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  for (int i = threadIdx.x; i < p.R * 16; i += kThreads1) {
    atomicAdd(p.dbv + i, db_acc[i]);
    atomicAdd(p.dbq + i, db_acc[DB_FLOATS + i]);
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =========================================================================== //
// kernel 2
// =========================================================================== //
// Rank-outer: CTA (r, sample chunk) keeps T_r resident and streams the dN1_r rows of its samples EIGHT samples per stage:
//   B5  dAc^T[l, (s,a)] = T_r[l, x] . dN1[(s,a), x]^T        M = 64 (16 valid rows l), N = 64 (8 samples x 8), K = x = 512
//   B6  dT_r^T[x, l]   += dN1[(s,a), x]^T . Ac_r[(s,a), l]    4 tiles of 128 x, N = 16, K = (s,a) = 64 per stage
// Round 1 streamed two samples per stage from ONE issuing thread: 36 MMAs per pair at ~100 cycles of issue latency each
// made that thread the bottleneck (167 us for 1024 rows; tools/ubench).  Now four threads issue, each 8 of the 32 K steps
// of B5 into its own partial accumulator (two partials share a column range through the 16-lane interleave of M = 64
// accumulators) and one of the four B6 tiles: 12 MMAs per thread per 8 samples instead of 144 on one thread.
constexpr int kThreads2 = 320;     // warp 0 TMA | warps 1, 2, 3, 9 issuers (2 also owns TMEM) | warps 4, 8 dAc epilogue | 4-7 final dT flush
constexpr int OCT = 8;                                   // samples per stage
constexpr int S2_DN1 = 0, S2_AC = OCT * 8192, S2_BYTES = OCT * (8192 + 1024), S2_RING = 2;
constexpr uint32_t TM2_D5 = 0, TM2_D6 = 256;             // D5: 2 slots x 2 column ranges x 64 columns (x 2 lane halves)
enum { C_TTFULL = 0, C_SFULL = 1, C_SEMPTY = C_SFULL + S2_RING, C_D5FULL = C_SEMPTY + S2_RING, C_D5EMPTY = C_D5FULL + 2,
       C_D6FULL = C_D5EMPTY + 2, C_COUNT = C_D6FULL + 1 };

struct Bwd2Params {
  bf16* dza;
  float *dba, *dtpack;
  int B, A, R, n_chunks;
};
constexpr size_t kSmem2 = (size_t)T_BYTES + S2_RING * S2_BYTES + C_COUNT * 8 + 16 + 1024;

__global__ void __launch_bounds__(kThreads2, 1)
trilinear_bwd2_tc_kernel(const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_dn1,
                         const __grid_constant__ CUtensorMap tmap_a8, const Bwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sT = base;
  const uint32_t sS = sT + T_BYTES;
  const uint32_t sBar = sS + S2_RING * S2_BYTES;
  const uint32_t tmem_slot = sBar + C_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x % p.R, chunk = blockIdx.x / p.R;
  const int RD = p.R * 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_t);
    tma_prefetch_desc(&tmap_dn1);
    tma_prefetch_desc(&tmap_a8);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar(C_TTFULL), 1);
    for (int s = 0; s < S2_RING; ++s) {
      mbar_init(bar(C_SFULL + s), 1);
      mbar_init(bar(C_SEMPTY + s), 5);          // four issuers (commit after their last MMA on the stage) + the epilogue warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(C_D5FULL + s), 4);
      mbar_init(bar(C_D5EMPTY + s), 1);
    }
    mbar_init(bar(C_D6FULL), 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_prologue_done();      // everything above touched only this CTA's shared memory / TMEM

  const int n_my = (p.B - chunk + p.n_chunks - 1) / p.n_chunks;      // samples chunk, chunk + n_chunks, ...
  const int n_oct = (n_my + OCT - 1) / OCT;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(C_TTFULL), T_BYTES);
#pragma unroll
      for (int c = 0; c < 8; ++c) tma_load_3d(&tmap_t, bar(C_TTFULL), sT + c * 2048, c * 64, r * 16, 0);
      for (int oi = 0; oi < n_oct; ++oi) {
        const int slot = oi % S2_RING;
        mbar_wait(bar(C_SEMPTY + slot), ((oi / S2_RING) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(C_SFULL + slot), S2_BYTES);
        const uint32_t dst = sS + slot * S2_BYTES;
#pragma unroll
        for (int s = 0; s < OCT; ++s) {
          const int b = chunk + (OCT * oi + s) * p.n_chunks;            // b >= B: the boxes are out of range -> zeros
          // one box (64 x, 8 rows a, 8 chunks) per sample: lands as [chunk][8 rows][128 B] (rows a >= A zero-filled)
          tma_load_4d(&tmap_dn1, bar(C_SFULL + slot), dst + S2_DN1 + s * 8192, 0, 0, 0, b * p.R + r);
          tma_load_3d(&tmap_a8, bar(C_SFULL + slot), dst + S2_AC + s * 1024, (r >> 2) * 64, 0, b);
        }
      }
    }
  } else if (warp == 1 || warp == 2 || warp == 3 || warp == 9) {
    if (lane == 0) {
      const int w = warp == 9 ? 3 : warp - 1;                 // issuer index: K steps 8w .. 8w+7 of B5, tile w of B6
      const uint32_t id_b5 = make_idesc_rt(64, 64, 0, 0);     // only rows l < 16 are valid: M = 64 halves the A-operand read
      const uint32_t id_b6 = make_idesc_rt(128, 16, 1, 1);
      mbar_wait(bar(C_TTFULL), 0);
      // partial accumulator of this issuer: column range w >> 1, lane half w & 1 (M = 64 accumulators use 16 lanes of
      // every 32-lane quarter; the second one sits in the other 16 -- tcgen05 "interleaved" allocation)
      const uint32_t d5_off = TM2_D5 + (w >> 1) * 64 + (static_cast<uint32_t>((w & 1) * 16) << 16);
      for (int oi = 0; oi < n_oct; ++oi) {
        const int slot = oi % S2_RING;
        const uint32_t st = sS + slot * S2_BYTES;
        mbar_wait(bar(C_SFULL + slot), (oi / S2_RING) & 1);
        mbar_wait(bar(C_D5EMPTY + (oi & 1)), ((oi >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        // B5: dAc^T[l, (s,a)] = T_r[l, x] . dN1[(s,a), x]^T over this issuer's 128 x (chunks 2w, 2w+1).
        // dN1 stage layout [sample][chunk][8 rows][128 B]: the 8-row groups of the N operand are 8192 B apart (SBO)
        {
          const uint64_t da0 = make_smem_desc_sw128(sT + 2 * w * 2048, 16u, 1024u);
          const uint64_t db0 = make_smem_desc_sw128(st + S2_DN1 + 2 * w * 1024, 16u, 8192u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_ss(tmem_base + d5_off + (oi & 1) * 128, da0 + (uint64_t)(((ks >> 2) * 2048 + (ks & 3) * 32) >> 4),
                         db0 + (uint64_t)(((ks >> 2) * 1024 + (ks & 3) * 32) >> 4), id_b5, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar(C_D5FULL + (oi & 1)));
        // B6: dT_r^T[x, l] += dN1[(s,a), x]^T . Ac_r[(s,a), l], tile w (x = 128 w ..): K = (s,a) = 64 = 4 steps of 2 samples.
        // A: MN-major, 64 x per 128-byte row (chunks 1024 B apart = LBO), 8 K-rows per atom, next sample 8192 B (SBO)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16_ss(tmem_base + TM2_D6 + w * 16, make_smem_desc_sw128(st + S2_DN1 + 2 * w * 1024 + ks * 16384, 1024u, 8192u),
                       make_smem_desc_sw128(st + S2_AC + (r & 3) * 32 + ks * 2048, 0u, 1024u), id_b6, (oi > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar(C_SEMPTY + slot));
      }
      umma_commit(bar(C_D6FULL));
    }
  }
  if (warp == 4 || warp == 8) {
    // ---- dAc epilogue: warp 4 takes the even stages, warp 8 the odd ones (both sit on TMEM lanes 0-31).  Lanes 0-15 hold
    // l for the partials of issuers 0 / 2, lanes 16-31 the same l for issuers 1 / 3: add the two column ranges, fold the
    // lane halves with one shuffle, then lanes 0-15 write the even (s,a) columns and lanes 16-31 the odd ones. ----
    const int l = lane & 15;
    float dba_acc = 0.f;
    for (int oi = warp == 4 ? 0 : 1; oi < n_oct; oi += 2) {
      const int slot = oi % S2_RING;
      const uint32_t st = sS + slot * S2_BYTES;
      mbar_wait(bar(C_SFULL + slot), (oi / S2_RING) & 1);
      mbar_wait(bar(C_D5FULL + (oi & 1)), (oi >> 1) & 1);
      tcgen05_fence_after();
      float acc[64];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w0[32], w1[32];
        tmem_ld_32x32b_x32(tmem_base + TM2_D5 + (oi & 1) * 128 + h * 32, w0);
        tmem_ld_32x32b_x32(tmem_base + TM2_D5 + (oi & 1) * 128 + 64 + h * 32, w1);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[h * 32 + c] = __uint_as_float(w0[c]) + __uint_as_float(w1[c]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C_D5EMPTY + (oi & 1)));
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        // fold the lane halves; lanes 0-15 keep column c, lanes 16-31 column c + 1
        const float mine = lane < 16 ? acc[c] : acc[c + 1];
        const float send = lane < 16 ? acc[c + 1] : acc[c];
        const float tot = mine + __shfl_xor_sync(0xffffffffu, send, 16);
        const int sa = c + (lane >> 4);
        const int s = sa >> 3, a = sa & 7;
        const int b = chunk + (OCT * oi + s) * p.n_chunks;
        if (a < p.A && b < p.B) {
          const int col = (r & 3) * 16 + l;
          const float act = bf16_bits_to_float(ld_shared_u16(st + S2_AC + s * 1024 + sw128_off(a, col)));
          const float gv = act > 0.f ? tot : 0.f;
          p.dza[((size_t)b * p.A + a) * RD + r * 16 + l] = __float2bfloat16(gv);
          dba_acc += gv;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(C_SEMPTY + slot));
    }
    dba_acc += __shfl_xor_sync(0xffffffffu, dba_acc, 16);
    if (lane < 16) atomicAdd(p.dba + r * 16 + l, dba_acc);
  }
  if (warp >= 4 && warp < 8) {
    // ---- final flush of dT_r (accumulated over every sample of this CTA) ----
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    mbar_wait(bar(C_D6FULL), 0);
    tcgen05_fence_after();
    if (n_oct > 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM2_D6 + t * 16, v);
        tmem_wait_ld();
        float* dst = p.dtpack + (size_t)r * 16 * 512 + t * 128 + L;
#pragma unroll
        for (int l = 0; l < 16; ++l) atomicAdd(dst + l * 512, __uint_as_float(v[l]));
      }
    }
    tcgen05_fence_before();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

size_t trilinear_bwd_tc_workspace(TriDims d) {        // dN1 [B][R][A][512] bf16
  return (size_t)d.B * d.R * d.A * 512 * sizeof(bf16);
}

// Returns -100 when the shape is outside the fast path.  dlm: [B][K][32 A] bf16 (gradient of the logits).
int trilinear_bwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const bf16* dlm, bf16* dn1,
                     bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dtpack, TriDims d,
                     cudaStream_t stream) {
  if (d.G != 2 || d.K > 64 || d.Q > 16 || d.A > 6 || (d.R & 3) != 0 || d.R > 32) return -100;
  const int RD = d.R * 16, N = 32 * d.A;
  CUtensorMap tt, tv, tq, ta, tdl, tdn, ta8;
  if (int rc = make_tmap_3d(&tt, tpack, 512, (uint64_t)d.R * 16, 1, 512, (uint64_t)d.R * 16 * 512, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&tv, vc, RD, d.K, d.B / d.VR, RD, (uint64_t)d.K * RD, 64, 64)) return rc;
  if (int rc = make_tmap_3d(&tq, qc, RD, d.Q, d.B, RD, (uint64_t)d.Q * RD, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&ta, ac, RD, d.A, d.B, RD, (uint64_t)d.A * RD, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&tdl, dlm, N, d.K, d.B, N, (uint64_t)d.K * N, 64, 64)) return rc;
  {   // dN1 [b][r][a][512] viewed as (64 x, A rows, 8 chunks, B*R): one box (64, 8, 8, 1) per (sample, rank)
    const uint64_t dims[4] = {64, (uint64_t)d.A, 8, (uint64_t)d.B * d.R};
    const uint64_t strides[3] = {512, 64, (uint64_t)d.A * 512};
    const uint32_t box[4] = {64, 8, 8, 1};
    if (int rc = make_tmap_4d(&tdn, dn1, dims, strides, box)) return rc;
  }
  if (int rc = make_tmap_3d(&ta8, ac, RD, d.A, d.B, RD, (uint64_t)d.A * RD, 64, 8)) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(trilinear_bwd1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem1);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(trilinear_bwd2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
    if (e != cudaSuccess) {
      set_error("trilinear_bwd_tc smem attr: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  Bwd1Params p1{dzv, dzq, dn1, dbv, dbq, d.B, d.K, d.Q, d.A, d.R, N, d.VR};
  const int grid1 = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(trilinear_bwd1_tc_kernel, dim3(grid1), dim3(kThreads1), kSmem1, stream, tt, tv, tq, ta, tdl, p1);
  if (int rc = check_launch("trilinear_bwd1_tc_kernel")) return rc;
  int n_chunks = kNumSMsB200 / d.R;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > (d.B + OCT - 1) / OCT) n_chunks = (d.B + OCT - 1) / OCT;
  if (n_chunks < 1) n_chunks = 1;
  Bwd2Params p2{dza, dba, dtpack, d.B, d.A, d.R, n_chunks};
  launch_pdl(trilinear_bwd2_tc_kernel, dim3(d.R * n_chunks), dim3(kThreads2), kSmem2, stream, tt, tdn, ta8, p2);
  return check_launch("trilinear_bwd2_tc_kernel");
}

}  // namespace cti
