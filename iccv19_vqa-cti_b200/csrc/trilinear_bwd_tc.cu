// Backward of the rank-R trilinear logit map on tcgen05 (glimpse-2 fast path).
//
//   L[b,k,(a,g,q)] = sum_r sum_i Vc_r[k,i] M_r[i,(a,g,q)],   M_r = N1_r x_j Qc_r,   N1_r = T_r x_l Ac_r
//
// Two kernels, both built from tcgen05.mma stages with TMEM accumulators and the dual-use swizzled operand tiles of
// tc_tiles.cuh.  What shapes them (round 2, tools/ubench + the launch list): every tcgen05.mma / tcgen05.commit is one
// entry of an in-order queue that drains at ~75 cycles per entry in kernels like these, whatever the tile size; the
// first version issued 40 small MMAs + 14 commits per (sample, rank) and ran at exactly 54 x 75 cycles per unit.  So
// the stages are batched until the instruction count, not the FLOP count, is small:
//
// 1. trilinear_bwd1_tc_kernel -- sample-outer, one step per rank QUAD t (4 ranks = 64 operand columns):
//      (N1 tiles of the quad come back from the forward: one 24 KB bulk copy instead of F1 + the core re-load)
//      B2  dM[(r4,i), n]      = Vc_t[k,(r4,i)]^T . dL[k,n]            4 MMAs 128 x 192 x 16 per quad   -> Dt tile [(a,g,i)][(r4,q)]
//      F2  M_r[(a,g,i), q]    = N1_r[(a,g,i),j] . Qc_r[q,j]^T         2 per rank                        -> M quad tile [(r4,i)][n]
//      B1  dVc[k,(r4,i)]      = dL[k,n] . Mq[(r4,i),n]^T              12 MMAs 128 x 64 x 16 per quad    -> ReLU mask, dzv
//      B3  X[(r4',q),(r4,j)]  = Dt[(a,g,i),(r4',q)]^T . N1[(a,g,i),(r4,j)]   12 per quad; the diagonal blocks r' = r are
//                                                                     dQc_r[q,j] (4x redundant FLOPs, a quarter of the MMAs)
//      B4  dN1_r[(a,g,i), j]  = Dt_r[(a,g,i),q] . Qc_r[q,j]           2 per rank                        -> bf16 workspace
//    11 MMAs + ~3 commits per (sample, rank).  Buffers are released by the converter warps once they have seen the
//    consuming MMAs' own completion barrier (no extra commits).  dL (bf16, [k][(a,g,q16)]) is loaded once per sample.
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

// Debug build (CTI_PROF=1 python build.py): per-role cycle accounting of kernel 1, read back with cti_debug_prof_read_bwd1().
// Slot layout: [block][role 0..15][counter 0..7].
#ifdef CTI_PROF
__device__ unsigned long long g_prof1[148 * 128];
__device__ unsigned long long g_trace1[16 * 128];        // block 0: time at which event e of quad c happens
#define TRACE(e, c) do { if (blockIdx.x == 0 && (c) < 128 && (threadIdx.x & 31) == 0) g_trace1[(e) * 128 + (c)] = clock64(); } while (0)
#define PROF_DECL unsigned long long prof_t0 = 0, prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; (void)prof_t0;
#define PROF_T0() (prof_t0 = clock64())
#define PROF_ADD(i) (prof_acc[i] += clock64() - prof_t0)
#define PROF_FLUSH(role)                                                                           \
  do {                                                                                             \
    if ((threadIdx.x & 31) == 0)                                                                   \
      for (int i_ = 0; i_ < 8; ++i_) g_prof1[(blockIdx.x * 16 + (role)) * 8 + i_] = prof_acc[i_]; \
  } while (0)
#else
#define PROF_DECL
#define PROF_T0()
#define PROF_ADD(i)
#define PROF_FLUSH(role)
#define TRACE(e, c)
#endif
#define PWAIT(i, ...) do { PROF_T0(); mbar_wait(__VA_ARGS__); PROF_ADD(i); } while (0)

// =========================================================================== //
// kernel 1
// =========================================================================== //
// 24 warps: 0 TMA | 1 B2 issuer | 2 F2 issuer (+ TMEM owner) | 3 B1 issuer | 22 B3 issuer | 23 B4 issuer |
//           4-7 C3 (dM -> Dt tile) | 8-11 C2 (M -> M tile) | 12-15 E1 (dVc -> dzv) | 16-19 E4 (dN1 -> workspace) | 20-21 E3 (dQc -> dzq)
constexpr int kThreads1 = 768;
constexpr int T_BYTES = 16 * 1024;
constexpr int OP_V = 0, OP_Q = 8192, OP_BYTES = 10240, OP_RING = 2;     // per quad: Vc [64 k][64 (r4,i)], Qc [16 q][64 (r4,j)]
constexpr int N1_BYTES = 24 * 1024;                           // [(a,g,i) 192][64 (r4,j)]   (the forward's tile image)
constexpr int DT_BYTES = 24 * 1024;                           // [(a,g,i) 192][64 (r4,q16)]
constexpr int MQ_BYTES = 24 * 1024;                           // 3 chunks x [64 (r4,i)][64 n]
constexpr int DL_CHUNKS = 3, DL_BYTES = DL_CHUNKS * 8192;     // dL tile: 3 chunks x [64 k][64 n]
// TMEM columns: B2 [0,192) | F2 2 x 32 | B4 2 x 32 | B1 2 x 64 | B3 64
constexpr uint32_t TM_B2 = 0, TM_F2 = 192, TM_B4 = 256, TM_B1 = 320, TM_B3 = 448;

enum { A_OPFULL = 0, A_OPEMPTY = 2, A_N1FULL = 4, A_N1EMPTY = 6, A_DLFULL = 8, A_DLEMPTY = 10, A_B2FULL = 12, A_B2EMPTY = 13,
       A_DTFULL = 14, A_DTEMPTY = 16, A_F2FULL = 18, A_F2EMPTY = 20, A_MFULL = 22, A_MEMPTY = 24, A_B1FULL = 26,
       A_B1EMPTY = 28, A_B3FULL = 30, A_B3EMPTY = 31, A_B4FULL = 32, A_B4EMPTY = 34, A_COUNT = 36 };

struct Bwd1Params {
  const uint8_t* n1;         // the forward's N1 quad tiles: [b][R / 4][A * 4096 bytes]
  bf16 *dzv, *dzq, *dn1;
  int B, K, Q, A, R, N;
  int VR;        // rows b share the v operand of row b / VR (dzv stays per row b)
};

// Regions that are read as a 128-row A operand although they hold fewer rows (dL chunks 64, N1 / Dt tiles 192 of 256:
// the extra accumulator lanes are never read back) come first, so the over-read stays inside this allocation.
constexpr size_t kSmem1 = (size_t)2 * DL_BYTES + 2 * N1_BYTES + 2 * DT_BYTES + 2 * MQ_BYTES + OP_RING * OP_BYTES +
                          A_COUNT * 8 + 16 + 1024;

__global__ void __launch_bounds__(kThreads1, 1)
trilinear_bwd1_tc_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q,
                         const __grid_constant__ CUtensorMap tmap_dl, const Bwd1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sDL = base;
  const uint32_t sN1 = sDL + 2 * DL_BYTES;
  const uint32_t sDT = sN1 + 2 * N1_BYTES;
  const uint32_t sM = sDT + 2 * DT_BYTES;
  const uint32_t sOp = sM + 2 * MQ_BYTES;
  const uint32_t sBar = sOp + OP_RING * OP_BYTES;
  const uint32_t tmem_slot = sBar + A_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_dl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(A_OPFULL + s), 1);
      mbar_init(bar(A_OPEMPTY + s), 18);      // C3 (for B2), C2 (for F2), E1 (own reads), E4 (for B4): 4 warps each; E3 (own reads): 2
      mbar_init(bar(A_N1FULL + s), 1);
      mbar_init(bar(A_N1EMPTY + s), 6);       // C2 (for F2) + E3 (for B3)
      mbar_init(bar(A_DLFULL + s), 1);
      mbar_init(bar(A_DLEMPTY + s), 8);       // C3 (for B2) + E1 (for B1), at the sample's last quad
      mbar_init(bar(A_DTFULL + s), 4);
      mbar_init(bar(A_DTEMPTY + s), 6);       // E4 (for B4) + E3 (for B3)
      mbar_init(bar(A_F2FULL + s), 1);
      mbar_init(bar(A_F2EMPTY + s), 4);
      mbar_init(bar(A_MFULL + s), 4);
      mbar_init(bar(A_MEMPTY + s), 4);        // E1 (for B1)
      mbar_init(bar(A_B1FULL + s), 1);
      mbar_init(bar(A_B1EMPTY + s), 4);
      mbar_init(bar(A_B4FULL + s), 1);
      mbar_init(bar(A_B4EMPTY + s), 4);
    }
    mbar_init(bar(A_B2FULL), 1);
    mbar_init(bar(A_B2EMPTY), 4);
    mbar_init(bar(A_B3FULL), 1);
    mbar_init(bar(A_B3EMPTY), 2);
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
  const int NQ = p.R >> 2;                      // rank quads per sample
  const int CQ = n_my * NQ;                     // quads of this CTA
  const int nt2 = (p.A + 3) >> 2;               // 128-row tiles over (a,g,i)
  const int kn = p.N >> 4;                      // K steps over n = (a,g,q16) / over (a,g,i)
  const int ktok = (p.K + 15) >> 4;             // K steps over regions
  const uint32_t n1_bytes = static_cast<uint32_t>(p.A) * 4096u;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one_sync()) {
      PROF_DECL
      auto load_dl = [&](int s_local, int bb) {
        const int slot = s_local & 1;
        PWAIT(2, bar(A_DLEMPTY + slot), ((s_local >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(A_DLFULL + slot), DL_BYTES);
        for (int c = 0; c < DL_CHUNKS; ++c) tma_load_3d(&tmap_dl, bar(A_DLFULL + slot), sDL + slot * DL_BYTES + c * 8192, c * 64, 0, bb);
      };
      int t = 0, sl = 0, b = blockIdx.x;
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        if (c == 0) load_dl(0, b);
        if (t == NQ / 2 && sl + 1 < n_my) load_dl(sl + 1, b + gridDim.x);
        PWAIT(0, bar(A_OPEMPTY + slot), ph ^ 1);
        TRACE(0, c);
        mbar_arrive_expect_tx(bar(A_OPFULL + slot), OP_BYTES);
        tma_load_3d(&tmap_v, bar(A_OPFULL + slot), sOp + slot * OP_BYTES + OP_V, t * 64, 0, b / p.VR);
        tma_load_3d(&tmap_q, bar(A_OPFULL + slot), sOp + slot * OP_BYTES + OP_Q, t * 64, 0, b);
        PWAIT(1, bar(A_N1EMPTY + slot), ph ^ 1);
        mbar_arrive_expect_tx(bar(A_N1FULL + slot), n1_bytes);
        bulk_load_1d(sN1 + slot * N1_BYTES, p.n1 + ((size_t)b * NQ + t) * n1_bytes, n1_bytes, bar(A_N1FULL + slot));
        if (++t == NQ) { t = 0; ++sl; b += gridDim.x; }
      }
      PROF_FLUSH(0);
    }
  } else if (warp == 1) {
    // ------------------------------ issuer: B2  dM[(r4,i), n] = Vc^T dL  (depends on TMA data only, runs ahead).
    // M = 128 with LBO = 0: accumulator lanes 64-127 repeat lanes 0-63, so all four C3 warps share the read-out ------
    if (elect_one_sync()) {
      const uint32_t id_b2 = make_idesc_rt(128, p.N, 1, 1);
      PROF_DECL
      int t = 0, sl = 0;
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        PWAIT(0, bar(A_OPFULL + slot), ph);
        if (t == 0) PWAIT(1, bar(A_DLFULL + (sl & 1)), (sl >> 1) & 1);
        PWAIT(2, bar(A_B2EMPTY), (c & 1) ^ 1);
        TRACE(1, c);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t da0 = desc_mnmajor(sOp + slot * OP_BYTES + OP_V, 0, 0);
        const uint64_t db0 = desc_mnmajor(sDL + (sl & 1) * DL_BYTES, 0, 8192);
#pragma unroll 1
        for (int ks = 0; ks < ktok; ++ks)
          umma_bf16_ss(tmem_base + TM_B2, da0 + (uint64_t)(ks * 2048 >> 4), db0 + (uint64_t)(ks * 2048 >> 4), id_b2, ks > 0 ? 1u : 0u);
        umma_commit(bar(A_B2FULL));
        PROF_ADD(3);
        if (++t == NQ) { t = 0; ++sl; }
      }
      PROF_FLUSH(1);
    }
  } else if (warp == 2) {
    // ------------------------------ issuer: F2  M_r[(a,g,i), q] = N1_r Qc_r^T  (per rank) ------------------------
    if (elect_one_sync()) {
      const uint32_t id_f2 = make_idesc_rt(128, 16, 0, 0);
      PROF_DECL
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        PWAIT(0, bar(A_OPFULL + slot), ph);
        PWAIT(1, bar(A_N1FULL + slot), ph);
        TRACE(7, c);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          const int u = c * 4 + s;
          PWAIT(2, bar(A_F2EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
          PROF_T0();
          tcgen05_fence_after();
          const uint64_t da = desc_kmajor(sN1 + slot * N1_BYTES, s);
          const uint64_t db = desc_kmajor(sOp + slot * OP_BYTES + OP_Q, s);
#pragma unroll 1
          for (int t2 = 0; t2 < nt2; ++t2)
            umma_bf16_ss(tmem_base + TM_F2 + (u & 1) * 32 + t2 * 16, da + (uint64_t)(t2 * 16384 >> 4), db, id_f2, 0u);
          umma_commit(bar(A_F2FULL + (u & 1)));
          PROF_ADD(3);
        }
      }
      PROF_FLUSH(2);
    }
  } else if (warp == 3) {
    // ------------------------------ issuer: B1  dVc[k, (r4,i)] = dL Mq^T  (per quad) ------------------------------
    if (elect_one_sync()) {
      const uint32_t id_b1 = make_idesc_rt(128, 64, 0, 0);
      PROF_DECL
      int t = 0, sl = 0;
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        if (t == 0) PWAIT(0, bar(A_DLFULL + (sl & 1)), (sl >> 1) & 1);
        PWAIT(1, bar(A_MFULL + slot), ph);
        PWAIT(2, bar(A_B1EMPTY + slot), ph ^ 1);
        TRACE(9, c);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t da0 = desc_kmajor(sDL + (sl & 1) * DL_BYTES, 0), db0 = desc_kmajor(sM + slot * MQ_BYTES, 0);
#pragma unroll 1
        for (int ks = 0; ks < kn; ++ks) {
          const uint64_t o = (uint64_t)(((ks >> 2) * 8192 + (ks & 3) * 32) >> 4);
          umma_bf16_ss(tmem_base + TM_B1 + slot * 64, da0 + o, db0 + o, id_b1, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar(A_B1FULL + slot));
        PROF_ADD(3);
        if (++t == NQ) { t = 0; ++sl; }
      }
      PROF_FLUSH(3);
    }
  } else if (warp == 22) {
    // ------------------------------ issuer: B3  dQc^T-blocks[(r4',q), (r4,j)] = Dt^T N1 (per quad; the diagonal blocks
    // r' = r are dQc_r) ------------------------------------------------------------------------------------------
    if (elect_one_sync()) {
      const uint32_t id_b3 = make_idesc_rt(128, 64, 1, 1);
      PROF_DECL
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        PWAIT(1, bar(A_N1FULL + slot), ph);
        PWAIT(2, bar(A_DTFULL + slot), ph);
        PWAIT(3, bar(A_B3EMPTY), (c & 1) ^ 1);
        TRACE(4, c);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t da0 = desc_mnmajor(sDT + slot * DT_BYTES, 0, 0), db0 = desc_mnmajor(sN1 + slot * N1_BYTES, 0, 0);
#pragma unroll 1
        for (int ks = 0; ks < kn; ++ks)
          umma_bf16_ss(tmem_base + TM_B3, da0 + (uint64_t)(ks * 2048 >> 4), db0 + (uint64_t)(ks * 2048 >> 4), id_b3, ks > 0 ? 1u : 0u);
        umma_commit(bar(A_B3FULL));
        PROF_ADD(5);
      }
      PROF_FLUSH(4);
    }
  } else if (warp == 23) {
    // ------------------------------ issuer: B4  dN1_r[(a,g,i), j] = Dt_r Qc_r (per rank) ---------------------------
    if (elect_one_sync()) {
      const uint32_t id_b4 = make_idesc_rt(128, 16, 0, 1);
      PROF_DECL
      for (int c = 0; c < CQ; ++c) {
        const int slot = c & 1, ph = (c >> 1) & 1;
        const uint32_t dt = sDT + slot * DT_BYTES, op = sOp + slot * OP_BYTES;
        PWAIT(0, bar(A_OPFULL + slot), ph);
        PWAIT(2, bar(A_DTFULL + slot), ph);
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          const int u = c * 4 + s;
          PWAIT(4, bar(A_B4EMPTY + (u & 1)), ((u >> 1) & 1) ^ 1);
          PROF_T0();
          tcgen05_fence_after();
          const uint64_t dq = desc_mnmajor(op + OP_Q + s * 32, 0, 0);
#pragma unroll 1
          for (int t2 = 0; t2 < nt2; ++t2)
            umma_bf16_ss(tmem_base + TM_B4 + (u & 1) * 32 + t2 * 16, desc_kmajor(dt + t2 * 16384, s), dq, id_b4, 0u);
          umma_commit(bar(A_B4FULL + (u & 1)));
          PROF_ADD(5);
        }
      }
      PROF_FLUSH(9);
    }
  } else if (warp < 8) {
    // ------------------------------ C3: dM (TMEM, lane (r4,i), column n = (a,g,q16)) -> Dt tile rows (a,g,i), columns
    // (r4,q16).  Lanes 64-127 repeat lanes 0-63: warps on lanes 0-63 convert (a,g) < A, the others (a,g) >= A. ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int r4 = (L & 63) >> 4, i = L & 15, ag0 = (L >> 6) * p.A;
    PROF_DECL
    int t = 0, sl = 0;
    for (int c = 0; c < CQ; ++c) {
      const int slot = c & 1, ph = (c >> 1) & 1;
      PWAIT(0, bar(A_B2FULL), c & 1);
      tcgen05_fence_after();
      PWAIT(1, bar(A_DTEMPTY + slot), ph ^ 1);
      if (warp == 4) TRACE(2, c);
      PROF_T0();
      const uint32_t dt = sDT + slot * DT_BYTES;
      for (int x = 0; x < p.A; x += 2) {
        const int ag = ag0 + x;
        uint32_t vv[2][16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B2 + ag * 16, vv[0]);
        if (x + 1 < p.A) tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B2 + ag * 16 + 16, vv[1]);
        tmem_wait_ld();
#pragma unroll
        for (int y2 = 0; y2 < 2; ++y2) {
          if (x + y2 < p.A) {
            uint32_t pk[8];
#pragma unroll
            for (int y = 0; y < 8; ++y) pk[y] = pack_bf16x2(__uint_as_float(vv[y2][2 * y]), __uint_as_float(vv[y2][2 * y + 1]));
            const uint32_t row = (ag + y2) * 16 + i;
            st_shared_v4(dt + sw128_off(row, r4 * 16), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(dt + sw128_off(row, r4 * 16 + 8), pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(A_DTFULL + slot));
        mbar_arrive(bar(A_B2EMPTY));
        mbar_arrive(bar(A_OPEMPTY + slot));                        // B2 has retired: it no longer reads the Vc chunk ...
        if (t == NQ - 1) mbar_arrive(bar(A_DLEMPTY + (sl & 1)));   // ... nor, after the sample's last quad, the dL tile
      }
      PROF_ADD(2);
      if (warp == 4) TRACE(3, c);
      if (++t == NQ) { t = 0; ++sl; }
    }
    if (warp == 4) PROF_FLUSH(5);
  } else if (warp < 12) {
    // ------------------------------ C2: M_r (TMEM, lane (a,g,i), column q16) -> M quad tile rows (r4,i), columns n ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    PROF_DECL
    for (int c = 0; c < CQ; ++c) {
      const int slot = c & 1, ph = (c >> 1) & 1;
      const uint32_t mt = sM + slot * MQ_BYTES;
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const int u = c * 4 + s;
        PWAIT(0, bar(A_F2FULL + (u & 1)), (u >> 1) & 1);
        tcgen05_fence_after();
        if (s == 0) PWAIT(1, bar(A_MEMPTY + slot), ph ^ 1);
        PROF_T0();
        uint32_t vv[2][16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + (u & 1) * 32, vv[0]);
        if (nt2 > 1) tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + (u & 1) * 32 + 16, vv[1]);
        tmem_wait_ld();
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const int rho = t2 * 128 + L;
          const int a = rho >> 5, i = rho & 15;
          if (t2 < nt2 && a < p.A) {
            const int ag = rho >> 4;
            uint32_t pk[8];
#pragma unroll
            for (int y = 0; y < 8; ++y) pk[y] = pack_bf16x2(__uint_as_float(vv[t2][2 * y]), __uint_as_float(vv[t2][2 * y + 1]));
            const uint32_t tile = mt + (ag >> 2) * 8192u;
            const uint32_t c0 = (ag & 3) * 16;
            st_shared_v4(tile + sw128_off(s * 16 + i, c0), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(tile + sw128_off(s * 16 + i, c0 + 8), pk[4], pk[5], pk[6], pk[7]);
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(A_F2EMPTY + (u & 1)));
          if (s == 3) {
            mbar_arrive(bar(A_MFULL + slot));
            mbar_arrive(bar(A_N1EMPTY + slot));      // F2 of the whole quad has retired: N1 tile and Qc chunk are free of it
            mbar_arrive(bar(A_OPEMPTY + slot));
          }
        }
        PROF_ADD(2);
      }
      if (warp == 8) TRACE(8, c);
    }
    if (warp == 8) PROF_FLUSH(6);
  } else if (warp < 16) {
    // ------------------------------ E1: dVc[k, (r4,i)] -> ReLU mask -> dzv (128 contiguous bytes per region) ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int RD = p.R * 16;
    PROF_DECL
    int t = 0, sl = 0, b = blockIdx.x;
    for (int c = 0; c < CQ; ++c) {
      const int slot = c & 1, ph = (c >> 1) & 1;
      PWAIT(0, bar(A_OPFULL + slot), ph);
      PWAIT(1, bar(A_B1FULL + slot), ph);
      if (warp == 12) TRACE(10, c);
      PROF_T0();
      tcgen05_fence_after();
      const uint32_t rb = sOp + slot * OP_BYTES + OP_V + (L >> 3) * 1024u + (L & 7) * 128u;      // row k = L of the Vc chunk
      bf16* dst = p.dzv + ((size_t)b * p.K + L) * RD + t * 64;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
#ifdef CTI_PROF
        const unsigned long long tq0 = clock64();
#endif
        tmem_ld_32x32b_x32(tmem_base + lane_addr + TM_B1 + slot * 64 + hh * 32, v);
        tmem_wait_ld();
#ifdef CTI_PROF
        prof_acc[3] += clock64() - tq0;
#endif
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int h = hh * 2 + h2;
          if (L < p.K) {
            float gv[16];
            uint32_t w0[4], w1[4];
            ld_shared_v4(rb + (((h * 2) ^ (L & 7)) << 4), w0);
            ld_shared_v4(rb + (((h * 2 + 1) ^ (L & 7)) << 4), w1);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const float2 f0 = unpack_bf16x2(w0[x]), f1 = unpack_bf16x2(w1[x]);
              gv[2 * x] = f0.x > 0.f ? __uint_as_float(v[h2 * 16 + 2 * x]) : 0.f;
              gv[2 * x + 1] = f0.y > 0.f ? __uint_as_float(v[h2 * 16 + 2 * x + 1]) : 0.f;
              gv[8 + 2 * x] = f1.x > 0.f ? __uint_as_float(v[h2 * 16 + 8 + 2 * x]) : 0.f;
              gv[8 + 2 * x + 1] = f1.y > 0.f ? __uint_as_float(v[h2 * 16 + 8 + 2 * x + 1]) : 0.f;
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(gv[0], gv[1]); o0.y = pack_bf16x2(gv[2], gv[3]); o0.z = pack_bf16x2(gv[4], gv[5]); o0.w = pack_bf16x2(gv[6], gv[7]);
            o1.x = pack_bf16x2(gv[8], gv[9]); o1.y = pack_bf16x2(gv[10], gv[11]); o1.z = pack_bf16x2(gv[12], gv[13]); o1.w = pack_bf16x2(gv[14], gv[15]);
            reinterpret_cast<uint4*>(dst + h * 16)[0] = o0;
            reinterpret_cast<uint4*>(dst + h * 16)[1] = o1;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(A_B1EMPTY + slot));
        mbar_arrive(bar(A_MEMPTY + slot));                         // B1 has retired: the M quad tile is free ...
        mbar_arrive(bar(A_OPEMPTY + slot));
        if (t == NQ - 1) mbar_arrive(bar(A_DLEMPTY + (sl & 1)));   // ... and, after the last quad, the dL tile
      }
      PROF_ADD(2);
      if (warp == 12) TRACE(11, c);
      if (++t == NQ) { t = 0; ++sl; b += gridDim.x; }
    }
    if (warp == 12) PROF_FLUSH(7);
  } else if (warp < 20) {
    // ------------------------------ E4: dN1_r[(a,g,i), j] -> bf16 workspace [b][r][a][(i,g,j)] ------------------------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    PROF_DECL
    int t = 0, b = blockIdx.x;
    for (int c = 0; c < CQ; ++c) {
      const int slot = c & 1;
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const int u = c * 4 + s, r = t * 4 + s;
        PWAIT(3, bar(A_B4FULL + (u & 1)), (u >> 1) & 1);
        PROF_T0();
        tcgen05_fence_after();
        uint32_t vv[2][16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B4 + (u & 1) * 32, vv[0]);
        if (nt2 > 1) tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_B4 + (u & 1) * 32 + 16, vv[1]);
        tmem_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(A_B4EMPTY + (u & 1)));
          if (s == 3) {      // B4 of this quad has retired: Dt tile and Qc chunk are free of it
            mbar_arrive(bar(A_DTEMPTY + slot));
            mbar_arrive(bar(A_OPEMPTY + slot));
          }
        }
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const uint32_t (&v)[16] = vv[t2];
          const int rho = t2 * 128 + L;
          const int a = rho >> 5, g = (rho >> 4) & 1, i = rho & 15;
          if (t2 < nt2 && a < p.A) {
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
        PROF_ADD(4);
      }
      if (warp == 16) TRACE(6, c);
      if (++t == NQ) { t = 0; b += gridDim.x; }
    }
    if (warp == 16) PROF_FLUSH(8);
  } else if (warp < 22) {
    // ------------------------------ E3 (two warps, accumulator lanes 0-63): dQc_r[q, j] (diagonal blocks of B3) ->
    // ReLU mask -> dzq ---------------------------------------------------------------------------------------------
    const int qd = warp & 3, L = qd * 32 + lane;                  // qd = 0, 1
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int RD = p.R * 16;
    const int rr = L >> 4, q = L & 15;          // rank within the quad, question token
    const bool hi = lane & 16;
    PROF_DECL
    int t = 0, b = blockIdx.x;
    for (int c = 0; c < CQ; ++c) {
      const int slot = c & 1, ph = (c >> 1) & 1;
      PWAIT(0, bar(A_OPFULL + slot), ph);
      PWAIT(1, bar(A_B3FULL), c & 1);
      if (warp == 20) TRACE(5, c);
      PROF_T0();
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + TM_B3 + qd * 32, v);      // blocks r' = 2 qd (lanes 0-15), 2 qd + 1 (16-31)
      tmem_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(A_B3EMPTY));
        mbar_arrive(bar(A_DTEMPTY + slot));      // B3 has retired: Dt and N1 tiles are free of it
        mbar_arrive(bar(A_N1EMPTY + slot));
      }
      if (q < p.Q) {
        float gv[16];
        const uint32_t rb = sOp + slot * OP_BYTES + OP_Q + (q >> 3) * 1024u + (q & 7) * 128u;
        uint32_t w0[4], w1[4];
        ld_shared_v4(rb + (((rr * 2) ^ (q & 7)) << 4), w0);
        ld_shared_v4(rb + (((rr * 2 + 1) ^ (q & 7)) << 4), w1);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float2 f0 = unpack_bf16x2(w0[x]), f1 = unpack_bf16x2(w1[x]);
          gv[2 * x] = f0.x > 0.f ? __uint_as_float(hi ? v[16 + 2 * x] : v[2 * x]) : 0.f;
          gv[2 * x + 1] = f0.y > 0.f ? __uint_as_float(hi ? v[16 + 2 * x + 1] : v[2 * x + 1]) : 0.f;
          gv[8 + 2 * x] = f1.x > 0.f ? __uint_as_float(hi ? v[24 + 2 * x] : v[8 + 2 * x]) : 0.f;
          gv[8 + 2 * x + 1] = f1.y > 0.f ? __uint_as_float(hi ? v[24 + 2 * x + 1] : v[8 + 2 * x + 1]) : 0.f;
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(gv[0], gv[1]); o0.y = pack_bf16x2(gv[2], gv[3]); o0.z = pack_bf16x2(gv[4], gv[5]); o0.w = pack_bf16x2(gv[6], gv[7]);
        o1.x = pack_bf16x2(gv[8], gv[9]); o1.y = pack_bf16x2(gv[10], gv[11]); o1.z = pack_bf16x2(gv[12], gv[13]); o1.w = pack_bf16x2(gv[14], gv[15]);
        uint4* dst = reinterpret_cast<uint4*>(p.dzq + ((size_t)b * p.Q + q) * RD + (t * 4 + rr) * 16);
        dst[0] = o0;
        dst[1] = o1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(A_OPEMPTY + slot));      // this warp has read its Qc activations
      PROF_ADD(2);
      if (++t == NQ) { t = 0; b += gridDim.x; }
    }
    if (warp == 20) PROF_FLUSH(10);
  }

  tcgen05_fence_before();
  __syncthreads();
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
    if (elect_one_sync()) {
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
    if (elect_one_sync()) {
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

int debug_prof_read_bwd1(unsigned long long* host_dst, int n) {
#ifdef CTI_PROF
  if (n == 16 * 128) return (int)cudaMemcpyFromSymbol(host_dst, g_trace1, sizeof(unsigned long long) * 16 * 128);
  return (int)cudaMemcpyFromSymbol(host_dst, g_prof1, sizeof(unsigned long long) * (n < 148 * 128 ? n : 148 * 128));
#else
  (void)host_dst; (void)n;
  return -1;
#endif
}

size_t trilinear_bwd_tc_workspace(TriDims d) {        // dN1 [B][R][A][512] bf16
  return (size_t)d.B * d.R * d.A * 512 * sizeof(bf16);
}

// Returns -100 when the shape is outside the fast path (or the forward did not save its N1 tiles).
// dlm: [B][K][32 A] bf16 (gradient of the logits); n1: the tiles written by trilinear_fwd_tc (trilinear_n1_bytes).
int trilinear_bwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const bf16* dlm, const void* n1,
                     bf16* dn1, bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dtpack,
                     TriDims d, cudaStream_t stream) {
  if (n1 == nullptr) return -100;
  if (d.G != 2 || d.K > 64 || d.Q > 16 || d.A > 6 || (d.R & 3) != 0 || d.R > 32) return -100;
  const int RD = d.R * 16, N = 32 * d.A;
  CUtensorMap tt, tv, tq, tdl, tdn, ta8;
  if (int rc = make_tmap_3d(&tt, tpack, 512, (uint64_t)d.R * 16, 1, 512, (uint64_t)d.R * 16 * 512, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&tv, vc, RD, d.K, d.B / d.VR, RD, (uint64_t)d.K * RD, 64, 64)) return rc;
  if (int rc = make_tmap_3d(&tq, qc, RD, d.Q, d.B, RD, (uint64_t)d.Q * RD, 64, 16)) return rc;
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
  Bwd1Params p1{static_cast<const uint8_t*>(n1), dzv, dzq, dn1, d.B, d.K, d.Q, d.A, d.R, N, d.VR};
  const int grid1 = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(trilinear_bwd1_tc_kernel, dim3(grid1), dim3(kThreads1), kSmem1, stream, tv, tq, tdl, p1);
  if (int rc = check_launch("trilinear_bwd1_tc_kernel")) return rc;
  int n_chunks = kNumSMsB200 / d.R;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > (d.B + OCT - 1) / OCT) n_chunks = (d.B + OCT - 1) / OCT;
  if (n_chunks < 1) n_chunks = 1;
  Bwd2Params p2{dza, dba, dtpack, d.B, d.A, d.R, n_chunks};
  launch_pdl(trilinear_bwd2_tc_kernel, dim3(d.R * n_chunks), dim3(kThreads2), kSmem2, stream, tt, tdn, ta8, p2);
  if (int rc = check_launch("trilinear_bwd2_tc_kernel")) return rc;
  // bias gradients of the image / question side: column sums of dzv / dzq (kept out of kernel 1: the shuffle trees and
  // atomics cost more there, in instruction-cache footprint, than two streaming passes over 65 MB)
  if (int rc = act_bwd_bias(dzv, 1, nullptr, nullptr, dbv, (long)d.B * d.K, RD, stream)) return rc;
  return act_bwd_bias(dzq, 1, nullptr, nullptr, dbq, (long)d.B * d.Q, RD, stream);
}

}  // namespace cti
