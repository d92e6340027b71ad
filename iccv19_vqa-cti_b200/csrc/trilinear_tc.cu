// Rank-R trilinear logit map on tcgen05 (forward), the glimpse-2 fast path of TCNet.forward
// (reference src/tc.py:46-52 with the mode products of src/Tensor.py:6-19):
//
//   L[b,k,q,a,g] = sum_r sum_{i,j,l} T_eff[r,i,j,l,g] Vc[b,k,r,i] Qc[b,q,r,j] Ac[b,a,r,l]
//
// contracted per sample and per rank in the minimal-FLOP order a -> q -> v (T_min),
// every stage a tcgen05.mma with its accumulator in TMEM:
//   F1   N1^T[(i,g,j), a]  = T_r^T[(i,g,j), l] . Ac_r[a, l]^T        4 x (128 x 16 x 16)
//   F2   M[(a,g,i), q]     = N1[(a,g,i), j]    . Qc_r[q, j]^T        ceil(A/4) x (128 x 16 x 16)
//   III  L[k, (a,g,q)]    += Vc_r[k, i]        . M[i, (a,g,q)]       128 x 32A x 16, accumulated over r in TMEM
// Between the stages two warp groups move the fp32 accumulator to the bf16 operand tile of the next
// stage (TMEM -> registers -> 128B-swizzled shared memory, tc_tiles.cuh); N1, M and the K x Q x A x R
// intermediate never leave the SM and the (B,K,Q,A,G) accumulator is written to HBM exactly once.
// The stages of consecutive ranks are software pipelined: F1(u), F2(u-1), III(u-2).
//
// Operands arrive by TMA: the packed core T_r (16 KB per rank, L2 resident, ring of 3) and, per four
// ranks, one 64-column chunk of the sample's Vc / Qc / Ac rows (ring of 3).
//
// Roles (384 threads): warp 0 TMA | warps 1, 3, 2 one MMA issuer per stage (warp 2 also owns TMEM) |
// warps 4-7 N1 tiles | warps 8-11 M tiles + per-sample epilogue (mask, coalesced (B,G,K,Q,A) store).
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

namespace cti {

namespace {

using bf16 = __nv_bfloat16;

// Debug build (CTI_PROF=1 python build.py --force): per-role cycle accounting of the forward kernel, read back with
// cti_debug_prof_read().  Slot layout: [block][role 0..7][counter 0..7].
#ifdef CTI_PROF
__device__ unsigned long long g_prof[148 * 64];
#define PROF_DECL unsigned long long prof_t0 = 0, prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; (void)prof_t0;
#define PROF_T0() (prof_t0 = clock64())
#define PROF_ADD(i) (prof_acc[i] += clock64() - prof_t0)
#define PROF_FLUSH(role)                                                                         \
  do {                                                                                           \
    if ((threadIdx.x & 31) == 0)                                                                 \
      for (int i_ = 0; i_ < 8; ++i_) g_prof[(blockIdx.x * 8 + (role)) * 8 + i_] = prof_acc[i_]; \
  } while (0)
#else
#define PROF_DECL
#define PROF_T0()
#define PROF_ADD(i)
#define PROF_FLUSH(role)
#endif

constexpr int kThreads = 384;
constexpr int T_BYTES = 16 * 1024;            // T_r: [16 l][512 (i,g,j)] bf16 = 8 chunks x [16 rows][128 B]
constexpr int T_RING = 3;
constexpr int OP_V = 0, OP_Q = 8192, OP_A = 10240, OP_BYTES = 12288;   // Vc [64][64], Qc [16][64], Ac [16][64]
constexpr int OP_RING = 3;
constexpr int N1_BYTES = 32 * 1024;           // [(a8,g,i) 256 rows][64 cols]: columns (r % 4) * 16 + j -- one tile holds 4 ranks
constexpr int M_BYTES = 8 * 1024;             // [16 i][<= 256 (a,g,q16)] = 4 chunks x [16 rows][128 B]
// Ring depths.  The per-(sample, rank) chain F1 -> convert -> F2 -> convert -> III is several thousand cycles of
// latency (TMEM load, smem store, proxy fence, mbarrier hand-offs), so throughput = steps in flight / latency.
constexpr int F1_RING = 3, N1_RING = 8, F2_RING = 4, M_RING = 4;
constexpr uint32_t TM_F1 = 0, TM_F2 = F1_RING * 64, TM_ACC = TM_F2 + F2_RING * 32;      // 192 + 128 + N (<= 192) columns

enum { B_TFULL = 0, B_TEMPTY = 3, B_OPFULL = 6, B_OPEMPTY = 9, B_F1FULL = 12, B_F1EMPTY = B_F1FULL + F1_RING,
       B_N1FULL = B_F1EMPTY + F1_RING, B_N1EMPTY = B_N1FULL + N1_RING, B_F2FULL = B_N1EMPTY + N1_RING,
       B_F2EMPTY = B_F2FULL + F2_RING, B_MFULL = B_F2EMPTY + F2_RING, B_MEMPTY = B_MFULL + M_RING,
       B_ACCFULL = B_MEMPTY + M_RING, B_ACCEMPTY = B_ACCFULL + 1, B_COUNT = B_ACCEMPTY + 1 };

struct RingPos {                               // slot / phase of a D-deep mbarrier ring
  uint32_t slot = 0, ph = 0;
  __device__ __forceinline__ void next(uint32_t depth) {
    if (++slot == depth) { slot = 0; ph ^= 1u; }
  }
};

struct TriTcParams {
  const uint8_t* rowmask;
  float* logits;
  int B, K, Q, A, R, N;      // N = 32 * A columns (a, g, q16)
  int VR;                    // rows b share the v operand (and mask) of row b / VR
};

__host__ __device__ inline size_t tri_tc_smem(int K, int Q, int A) {
  return (size_t)T_RING * T_BYTES + OP_RING * OP_BYTES + 2 * N1_BYTES + M_RING * M_BYTES + (size_t)2 * K * Q * A * 4 + 16 +
         B_COUNT * 8 + 16 + 1024;
}

__global__ void __launch_bounds__(kThreads, 1)
trilinear_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_t, const __grid_constant__ CUtensorMap tmap_v,
                        const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_a,
                        const TriTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sT = base;
  const uint32_t sOp = sT + T_RING * T_BYTES;
  const uint32_t sN1 = sOp + OP_RING * OP_BYTES;
  const uint32_t sM = sN1 + 2 * N1_BYTES;
  const uint32_t sOut = sM + M_RING * M_BYTES;
  const int out_floats = 2 * p.K * p.Q * p.A;
  const uint32_t sBar = (sOut + out_floats * 4 + 15u) & ~15u;
  const uint32_t tmem_slot = sBar + B_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };
  float* out_stage = reinterpret_cast<float*>(smem_raw + (sOut - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_t);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_a);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(bar(B_TFULL + s), 1);
      mbar_init(bar(B_TEMPTY + s), 1);
      mbar_init(bar(B_OPFULL + s), 1);
      mbar_init(bar(B_OPEMPTY + s), 1);
    }
    for (int s = 0; s < F1_RING; ++s) {
      mbar_init(bar(B_F1FULL + s), 1);
      mbar_init(bar(B_F1EMPTY + s), 4);
    }
    for (int s = 0; s < N1_RING; ++s) {
      mbar_init(bar(B_N1FULL + s), 4);
      mbar_init(bar(B_N1EMPTY + s), 1);
    }
    for (int s = 0; s < F2_RING; ++s) {
      mbar_init(bar(B_F2FULL + s), 1);
      mbar_init(bar(B_F2EMPTY + s), 4);
    }
    for (int s = 0; s < M_RING; ++s) {
      mbar_init(bar(B_MFULL + s), 4);
      mbar_init(bar(B_MEMPTY + s), 1);
    }
    mbar_init(bar(B_ACCFULL), 1);
    mbar_init(bar(B_ACCEMPTY), 4);
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
  const int U = n_my * p.R;                       // (sample, rank) steps of this CTA
  const int nt2 = (p.A + 3) >> 2;                 // 128-row tiles of the F2 output (rows = (a,g,i))

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      PROF_DECL
      uint32_t tslot = 0, tph = 0, oslot = 0, oph = 0;
      int r = 0, b = blockIdx.x;
      for (int u = 0; u < U; ++u) {
        if ((r & 3) == 0) {
          PROF_T0();
          mbar_wait(bar(B_OPEMPTY + oslot), oph ^ 1u);
          PROF_ADD(0);
          mbar_arrive_expect_tx(bar(B_OPFULL + oslot), OP_BYTES);
          const uint32_t dst = sOp + oslot * OP_BYTES;
          tma_load_3d(&tmap_v, bar(B_OPFULL + oslot), dst + OP_V, r * 16, 0, b / p.VR);
          tma_load_3d(&tmap_q, bar(B_OPFULL + oslot), dst + OP_Q, r * 16, 0, b);
          tma_load_3d(&tmap_a, bar(B_OPFULL + oslot), dst + OP_A, r * 16, 0, b);
          if (++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        }
        PROF_T0();
        mbar_wait(bar(B_TEMPTY + tslot), tph ^ 1u);
        PROF_ADD(1);
        mbar_arrive_expect_tx(bar(B_TFULL + tslot), T_BYTES);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          tma_load_3d(&tmap_t, bar(B_TFULL + tslot), sT + tslot * T_BYTES + c * 2048, c * 64, r * 16, 0);
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) { r = 0; b += gridDim.x; }
      }
      PROF_FLUSH(0);
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer 1: F1(u)  N1^T = T_r^T . Ac_r^T ------------------------
    // (three issuing threads, one per stage: a single thread's serial descriptor / barrier code was the
    //  bottleneck of the first version; tcgen05.commit tracks the MMAs of the committing thread only)
    if (lane == 0) {
      const uint32_t id_f1 = make_idesc_rt(128, 16, 1, 0);
      RingPos f1;
      PROF_DECL
      uint32_t tslot = 0, tph = 0, oslot = 0, oph = 0;
      int r = 0;
      for (int u = 0; u < U; ++u) {
        const uint32_t tt = sT + tslot * T_BYTES;
        PROF_T0();
        mbar_wait(bar(B_TFULL + tslot), tph);
        PROF_ADD(0);
        PROF_T0();
        if ((r & 3) == 0) mbar_wait(bar(B_OPFULL + oslot), oph);
        PROF_ADD(1);
        PROF_T0();
        mbar_wait(bar(B_F1EMPTY + f1.slot), f1.ph ^ 1u);
        PROF_ADD(2);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t db = desc_kmajor(sOp + oslot * OP_BYTES + OP_A, r & 3);
        const uint64_t da = desc_mnmajor(tt, 0, 2048);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          umma_bf16_ss(tmem_base + TM_F1 + f1.slot * 64 + t * 16, da + (uint64_t)(2 * t * 2048 >> 4), db, id_f1, 0u);
        umma_commit(bar(B_F1FULL + f1.slot));
        umma_commit(bar(B_TEMPTY + tslot));
        f1.next(F1_RING);
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        PROF_ADD(3);
      }
      PROF_FLUSH(1);
    }
  } else if (warp == 3) {
    // ------------------------------ MMA issuer 2: F2(u)  M = N1 . Qc_r^T ---------------------------------
    if (lane == 0) {
      const uint32_t id_f2 = make_idesc_rt(128, 16, 0, 0);
      RingPos n1, f2;
      PROF_DECL
      uint32_t oslot = 0, oph = 0;
      int r = 0;
      for (int u = 0; u < U; ++u) {
        PROF_T0();
        if ((r & 3) == 0) mbar_wait(bar(B_OPFULL + oslot), oph);
        PROF_ADD(0);
        PROF_T0();
        mbar_wait(bar(B_N1FULL + n1.slot), n1.ph);
        PROF_ADD(1);
        PROF_T0();
        mbar_wait(bar(B_F2EMPTY + f2.slot), f2.ph ^ 1u);
        PROF_ADD(2);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t db = desc_kmajor(sOp + oslot * OP_BYTES + OP_Q, r & 3);
        const uint64_t da = desc_kmajor(sN1 + (n1.slot >> 2) * N1_BYTES, n1.slot & 3);     // K step = rank within the quad tile
        for (int t2 = 0; t2 < nt2; ++t2)
          umma_bf16_ss(tmem_base + TM_F2 + f2.slot * 32 + t2 * 16, da + (uint64_t)(t2 * 16384 >> 4), db, id_f2, 0u);
        umma_commit(bar(B_F2FULL + f2.slot));
        umma_commit(bar(B_N1EMPTY + n1.slot));
        n1.next(N1_RING);
        f2.next(F2_RING);
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        PROF_ADD(3);
      }
      PROF_FLUSH(2);
    }
  } else if (warp == 2) {
    // ------------------------------ MMA issuer 3: III(u)  L += Vc_r . M  (this warp also owns TMEM) --------
    if (lane == 0) {
      const uint32_t id_3 = make_idesc_rt(128, p.N, 0, 1);
      RingPos mr;
      PROF_DECL
      uint32_t oslot = 0, oph = 0;
      int r = 0, sl = 0;
      for (int u = 0; u < U; ++u) {
        PROF_T0();
        if ((r & 3) == 0) mbar_wait(bar(B_OPFULL + oslot), oph);
        PROF_ADD(0);
        PROF_T0();
        mbar_wait(bar(B_MFULL + mr.slot), mr.ph);
        PROF_ADD(1);
        PROF_T0();
        if (r == 0) mbar_wait(bar(B_ACCEMPTY), (sl & 1) ^ 1);
        PROF_ADD(2);
        PROF_T0();
        tcgen05_fence_after();
        umma_bf16_ss(tmem_base + TM_ACC, desc_kmajor(sOp + oslot * OP_BYTES + OP_V, r & 3),
                     desc_mnmajor(sM + mr.slot * M_BYTES, 0, 2048), id_3, r > 0 ? 1u : 0u);
        umma_commit(bar(B_MEMPTY + mr.slot));
        mr.next(M_RING);
        if ((r & 3) == 3) umma_commit(bar(B_OPEMPTY + oslot));
        if (r == p.R - 1) umma_commit(bar(B_ACCFULL));
        if (++r == p.R) { r = 0; ++sl; }
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        PROF_ADD(3);
      }
      PROF_FLUSH(3);
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------ C1: N1^T (TMEM) -> N1 tile rows (a,g,i), columns j --------------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int j = L & 15, g = (L >> 4) & 1;
    RingPos f1, n1r;
    PROF_DECL
    for (int u = 0; u < U; ++u) {
      PROF_T0();
      mbar_wait(bar(B_F1FULL + f1.slot), f1.ph);
      PROF_ADD(0);
      tcgen05_fence_after();
      PROF_T0();
      mbar_wait(bar(B_N1EMPTY + n1r.slot), n1r.ph ^ 1u);
      PROF_ADD(1);
      PROF_T0();
      const uint32_t n1 = sN1 + (n1r.slot >> 2) * N1_BYTES;
      const uint32_t sub = n1r.slot & 3;                       // rank within the quad tile: columns sub * 16 + j
      uint32_t v[4][8];
#pragma unroll
      for (int t = 0; t < 4; ++t) tmem_ld_32x32b_x8(tmem_base + lane_addr + TM_F1 + f1.slot * 64 + t * 16, v[t]);
      tmem_wait_ld();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = 4 * t + (L >> 5);
        // row = a*32 + g*16 + i  ->  row>>3 = a*4 + g*2 + (i>>3),  row&7 = i&7
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
        mbar_arrive(bar(B_N1FULL + n1r.slot));
        mbar_arrive(bar(B_F1EMPTY + f1.slot));
      }
      f1.next(F1_RING);
      n1r.next(N1_RING);
      PROF_ADD(2);
    }
    if (warp == 4) PROF_FLUSH(4);
  } else if (warp >= 8) {
    // ------------------------------ C2: M (TMEM) -> M tile [i][(a,g,q16)];  per-sample epilogue ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int et = threadIdx.x - 8 * 32;                  // 0..127 within the group
    const int per_g = p.K * p.Q * p.A;
    int r = -1, sl = 0;
    RingPos f2, mr;
    PROF_DECL
    for (int u = 0; u < U; ++u) {
      if (++r == p.R) { r = 0; ++sl; }
      PROF_T0();
      mbar_wait(bar(B_F2FULL + f2.slot), f2.ph);
      PROF_ADD(0);
      tcgen05_fence_after();
      PROF_T0();
      mbar_wait(bar(B_MEMPTY + mr.slot), mr.ph ^ 1u);
      PROF_ADD(1);
      PROF_T0();
      const uint32_t mt = sM + mr.slot * M_BYTES;
      for (int t2 = 0; t2 < nt2; ++t2) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + f2.slot * 32 + t2 * 16, v);
        tmem_wait_ld();
        const int rho = t2 * 128 + L;
        const int a = rho >> 5, i = rho & 15;
        if (a < p.A) {
          const int ag = rho >> 4;                        // a*2 + g
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
        mbar_arrive(bar(B_MFULL + mr.slot));
        mbar_arrive(bar(B_F2EMPTY + f2.slot));
      }
      f2.next(F2_RING);
      mr.next(M_RING);
      PROF_ADD(2);
      if (r == p.R - 1) {
        // ---- sample epilogue: TMEM lane = region k, column (a,g,q16) -> (G,K,Q,A) order, mask, coalesced store
        const int b = blockIdx.x + sl * gridDim.x;
        PROF_T0();
        mbar_wait(bar(B_ACCFULL), sl & 1);
        PROF_ADD(3);
        PROF_T0();
        tcgen05_fence_after();
        const int k = L;
        const bool masked = (k < p.K) && p.rowmask != nullptr && p.rowmask[(size_t)(b / p.VR) * p.K + k] != 0;
        for (int ag = 0; ag < 2 * p.A; ++ag) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_ACC + ag * 16, v);
          tmem_wait_ld();
          if (k < p.K) {
            const int a = ag >> 1, gg = ag & 1;
            float* dst = out_stage + (size_t)gg * per_g + (size_t)k * p.Q * p.A + a;
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (q < p.Q) dst[q * p.A] = masked ? -INFINITY : __uint_as_float(v[q]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_ACCEMPTY));
        named_bar_sync(1, 128);
        float* gdst = p.logits + (size_t)b * 2 * per_g;
        for (int e = et; e < 2 * per_g; e += 128) gdst[e] = out_stage[e];
        named_bar_sync(1, 128);
        PROF_ADD(4);
      }
    }
    if (warp == 8) PROF_FLUSH(5);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int debug_prof_read(unsigned long long* host_dst, int n) {
#ifdef CTI_PROF
  return (int)cudaMemcpyFromSymbol(host_dst, g_prof, sizeof(unsigned long long) * (n < 148 * 64 ? n : 148 * 64));
#else
  (void)host_dst; (void)n;
  return -1;
#endif
}

// Returns -100 when the shape is outside the fast path (caller falls back to the generic kernel).
int trilinear_fwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack, const uint8_t* rowmask,
                     float* logits, TriDims d, cudaStream_t stream) {
  if (d.G != 2 || d.K > 64 || d.Q > 16 || d.A > 6 || (d.R & 3) != 0) return -100;      // TMEM: 320 + 32 A columns
  const size_t smem = tri_tc_smem(d.K, d.Q, d.A);
  if (smem > 227 * 1024) return -100;
  const int RD = d.R * 16;
  CUtensorMap tt, tv, tq, ta;
  if (int rc = make_tmap_3d(&tt, tpack, 512, (uint64_t)d.R * 16, 1, 512, (uint64_t)d.R * 16 * 512, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&tv, vc, RD, d.K, d.B / d.VR, RD, (uint64_t)d.K * RD, 64, 64)) return rc;
  if (int rc = make_tmap_3d(&tq, qc, RD, d.Q, d.B, RD, (uint64_t)d.Q * RD, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&ta, ac, RD, d.A, d.B, RD, (uint64_t)d.A * RD, 64, 16)) return rc;
  TriTcParams p{rowmask, logits, d.B, d.K, d.Q, d.A, d.R, 32 * d.A, d.VR};
  static size_t smem_set = 0;          // raise the dynamic shared memory limit once per size (not a stream operation)
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(trilinear_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("trilinear_fwd_tc smem attr: %s", cudaGetErrorString(e));
      return (int)e;
    }
    smem_set = smem;
  }
  const int grid = d.B < kNumSMsB200 ? d.B : kNumSMsB200;
  launch_pdl(trilinear_fwd_tc_kernel, dim3(grid), dim3(kThreads), smem, stream, tt, tv, tq, ta, p);
  return check_launch("trilinear_fwd_tc_kernel");
}

}  // namespace cti
