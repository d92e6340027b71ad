// Rank-R trilinear logit map on tcgen05 (forward), the glimpse-2 fast path of TCNet.forward
// (reference src/tc.py:46-52 with the mode products of src/Tensor.py:6-19):
//
//   L[b,k,q,a,g] = sum_r sum_{i,j,l} T_eff[r,i,j,l,g] Vc[b,k,r,i] Qc[b,q,r,j] Ac[b,a,r,l]
//
// contracted per sample and per rank in the minimal-FLOP order a -> q -> v (T_min),
// every stage a tcgen05.mma with its accumulator in TMEM:
//   F1   N1^T[x, a]        = T_r^T[x, l] . Ac_r[a, l]^T              4 x (128 x 16 x 16),  x = (j % 4, g, i, j / 4)
//   F2   M[(a,g,i), q]     = N1[(a,g,i), j]    . Qc_r[q, j]^T        ceil(A/4) x (128 x 16 x 16)
//   III  L[k, (a,g,q)]    += Vc_r[k, i]        . M[i, (a,g,q)]       128 x 32A x 16, accumulated over r in TMEM
// Between the stages converter warps move the fp32 accumulator to the bf16 operand tile of the next
// stage (TMEM -> registers -> 128B-swizzled shared memory, tc_tiles.cuh); N1, M and the K x Q x A x R
// intermediate never leave the SM and the (B,K,Q,A,G) accumulator is written to HBM exactly once.
//
// What shapes this kernel (round 2, measured with tools/ubench/tc_ubench.cu and the CTI_PROF build):
//   * one thread issues a tcgen05.mma every ~75 cycles and a tcgen05.commit every ~46 cycles whatever the tile size,
//     while a 128 x 16 x 16 MMA occupies the tensor pipe for 8 cycles: with 7 small MMAs and 5-7 commits per
//     (sample, rank) the ISSUING THREADS were the bottleneck of the first version (433 / 469 / 367 cycles of issue
//     work per unit on its three issuer threads).  Issue is spread over more threads here (F1 over two), commits are
//     batched per rank quad where the hand-off allows, and the core slot is released by a converter arrive instead of
//     a second commit;
//   * the N1 conversion used 24 two-byte shared-memory stores per thread and unit (2-way conflicts): the packed core now
//     arrives in a lane order (cti_trilinear_logits_fwd's `tpack_perm`) that puts four consecutive j of one (a,g,i) row
//     into one thread -- 6 conflict-free 8-byte stores -- and two converter groups alternate units;
//   * the per-sample epilogue (8-way bank conflicts, 6200 cycles, on the warps that also convert M) moved to its own
//     warps with a conflict-free staging layout: the contraction of the next sample overlaps it.
//
// Operands arrive by TMA: the packed core T_r (16 KB per rank, L2 resident, ring of 3) and, per four
// ranks, one 64-column chunk of the sample's Vc / Qc / Ac rows (ring of 3).
//
// Training: with `n1_save` the bf16 N1 quad tiles (the shared-memory image, 4 KB per answer token and rank quad) are
// also bulk-copied to HBM; the backward (trilinear_bwd_tc.cu) loads them back instead of recomputing F1 -- that takes
// the core re-load and a quarter of the tcgen05 instructions out of the backward for 0.2 GB of traffic per 1024 rows.
//
// Roles (672 threads): warp 0 TMA | warp 1 F1 issuer | warp 3 F2 issuer | warp 2 III issuer
// (+ TMEM owner) | warps 4-7, 8-11 N1 converters (even / odd units) | warps 12-15 M converters | warps 16-19 epilogue |
// warp 20 N1 store.
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

namespace cti {

namespace {

using bf16 = __nv_bfloat16;

// Debug build (CTI_PROF=1 python build.py): per-role cycle accounting, read back with cti_debug_prof_read().
// Slot layout: [block][role 0..7][counter 0..7].
#ifdef CTI_PROF
__device__ unsigned long long g_prof[148 * 64];
__device__ unsigned long long g_trace[8 * 256];          // block 0: time at which role x starts its work on unit u
#define TRACE(role, u) do { if (blockIdx.x == 0 && (u) < 256 && (threadIdx.x & 31) == 0) g_trace[(role) * 256 + (u)] = clock64(); } while (0)
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
#define TRACE(role, u)
#endif

constexpr int kThreads = 672;                 // 21 warps
constexpr int T_BYTES = 16 * 1024;            // T_r: [16 l][512 x] bf16 = 8 chunks x [16 rows][128 B]
constexpr int T_RING = 3;
constexpr int OP_V = 0, OP_Q = 8192, OP_A = 10240, OP_BYTES = 12288;   // Vc [64][64], Qc [16][64], Ac [16][64]
constexpr int OP_RING = 3;
constexpr int N1_BYTES = 24 * 1024;           // [(a6,g,i) 192 rows][64 cols]: columns (r % 4) * 16 + j -- one tile holds 4 ranks
                                              // (the second 128-row F2 operand over-reads 64 rows of the next buffer:
                                              //  accumulator rows a >= A are never read back)
constexpr int M_BYTES = 6 * 1024;             // [16 i][192 (a,g,q16)] = 3 chunks x [16 rows][128 B]
constexpr int F1_RING = 3, N1_RING = 8, F2_RING = 4, M_RING = 8;      // M_RING = 2 quad slots of 4 rank tiles
constexpr uint32_t TM_F1 = 0, TM_F2 = F1_RING * 64, TM_ACC = TM_F2 + F2_RING * 32;      // 192 + 128 + N (<= 192) columns

enum { B_TFULL = 0, B_TEMPTY = B_TFULL + T_RING, B_OPFULL = B_TEMPTY + T_RING, B_OPEMPTY = B_OPFULL + OP_RING,
       B_F1FULL = B_OPEMPTY + OP_RING, B_F1EMPTY = B_F1FULL + F1_RING, B_N1FULL = B_F1EMPTY + F1_RING,
       B_N1EMPTY = B_N1FULL + N1_RING /* one per quad tile */, B_F2FULL = B_N1EMPTY + 2, B_F2EMPTY = B_F2FULL + F2_RING,
       B_MFULL = B_F2EMPTY + F2_RING /* per quad slot */, B_MEMPTY = B_MFULL + 2, B_ACCFULL = B_MEMPTY + 2,
       B_ACCEMPTY = B_ACCFULL + 1, B_N1QFULL = B_ACCEMPTY + 1, B_COUNT = B_N1QFULL + 2 };

struct TriTcParams {
  const uint8_t* rowmask;
  float* logits;
  uint8_t* n1_save;          // optional: [b][R / 4][A * 4096 bytes], the N1 quad tiles for the backward
  int B, K, Q, A, R, N;      // N = 32 * A columns (a, g, q16)
  int VR;                    // rows b share the v operand (and mask) of row b / VR
};

__host__ __device__ inline int stage_pad(int Q, int A) { return (Q * A) | 1; }      // odd row pitch: conflict-free lanes = k
__host__ __device__ inline size_t tri_tc_smem(int K, int Q, int A) {
  return (size_t)T_RING * T_BYTES + OP_RING * OP_BYTES + 2 * N1_BYTES + M_RING * M_BYTES + (size_t)2 * K * stage_pad(Q, A) * 4 +
         16 + B_COUNT * 8 + 16 + 1024;
}

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
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
  const int pad = stage_pad(p.Q, p.A);
  const uint32_t sBar = (sOut + 2 * p.K * pad * 4 + 15u) & ~15u;
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
    for (int s = 0; s < T_RING; ++s) {
      mbar_init(bar(B_TFULL + s), 1);
      mbar_init(bar(B_TEMPTY + s), 1);          // one N1-converter thread, after F1FULL (= both issuers' MMAs retired)
    }
    for (int s = 0; s < OP_RING; ++s) {
      mbar_init(bar(B_OPFULL + s), 1);
      mbar_init(bar(B_OPEMPTY + s), 1);
    }
    for (int s = 0; s < F1_RING; ++s) {
      mbar_init(bar(B_F1FULL + s), 1);
      mbar_init(bar(B_F1EMPTY + s), 4);
    }
    for (int s = 0; s < 2; ++s) mbar_init(bar(B_N1EMPTY + s), p.n1_save != nullptr ? 2 : 1);   // F2 commit (+ the store)
    for (int s = 0; s < F2_RING; ++s) {
      mbar_init(bar(B_F2FULL + s), 1);
      mbar_init(bar(B_F2EMPTY + s), 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_MFULL + s), 4);           // the M converters, once per quad
      mbar_init(bar(B_MEMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) mbar_init(bar(B_N1QFULL + s), 8);      // both N1 converter groups, once per quad
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
    if (elect_one_sync()) {
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
        TRACE(0, u);
        mbar_arrive_expect_tx(bar(B_TFULL + tslot), T_BYTES);
        // one TMA operation per core slice: box (64 x, 16 l, 8 chunks) of the view [chunk][l][64] lands as the 8
        // consecutive [16 rows][128 B] swizzled chunks the MN-major descriptor expects
        tma_load_3d(&tmap_t, bar(B_TFULL + tslot), sT + tslot * T_BYTES, 0, r * 16, 0);
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) { r = 0; b += gridDim.x; }
      }
      PROF_FLUSH(0);
    }
  } else if (warp == 1) {
    // ------------------------------ F1 issuer: N1^T = T_r^T . Ac_r^T.  (Splitting the four tiles over several issuing
    // threads does not help: every tcgen05.mma / tcgen05.commit is an entry of ONE in-order queue per SM that drains at
    // ~75 cycles per entry in these kernels -- measured: the kernel time follows (MMAs + commits) per unit x 75 cycles
    // whatever the number of issuers; more issuers only add commits.) ----
    if (elect_one_sync()) {
      const uint32_t id_f1 = make_idesc_rt(128, 16, 1, 0);
      PROF_DECL
      uint32_t tslot = 0, tph = 0, oslot = 0, oph = 0, fslot = 0, fph = 0;
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
        mbar_wait(bar(B_F1EMPTY + fslot), fph ^ 1u);
        PROF_ADD(2);
        PROF_T0();
        tcgen05_fence_after();
        TRACE(1, u);
        const uint64_t db = desc_kmajor(sOp + oslot * OP_BYTES + OP_A, r & 3);
        const uint64_t da = desc_mnmajor(tt, 0, 2048);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          umma_bf16_ss(tmem_base + TM_F1 + fslot * 64 + t * 16, da + (uint64_t)(2 * t * 2048 >> 4), db, id_f1, 0u);
        umma_commit(bar(B_F1FULL + fslot));
        if (++fslot == F1_RING) { fslot = 0; fph ^= 1u; }
        if (++tslot == T_RING) { tslot = 0; tph ^= 1u; }
        if (++r == p.R) r = 0;
        if ((r & 3) == 0 && ++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        PROF_ADD(3);
      }
      PROF_FLUSH(1);
    }
  } else if (warp == 3) {
    // ------------------------------ F2 issuer: M = N1 . Qc_r^T.  One wait for the whole N1 quad tile, then the 8 MMAs of
    // its four ranks back to back (a single thread pays ~50 cycles per barrier wait and ~75 per MMA: per-rank hand-offs
    // made this loop 730 cycles per rank in the first version) ---------------------------------
    if (elect_one_sync()) {
      const uint32_t id_f2 = make_idesc_rt(128, 16, 0, 0);
      PROF_DECL
      uint32_t oslot = 0, oph = 0, fslot = 0, fph = 0;
      for (int uq = 0; uq < U; uq += 4) {
        const int qslot = (uq >> 2) & 1, qph = (uq >> 3) & 1;
        PROF_T0();
        mbar_wait(bar(B_OPFULL + oslot), oph);
        PROF_ADD(0);
        PROF_T0();
        mbar_wait(bar(B_N1QFULL + qslot), qph);
        PROF_ADD(1);
        tcgen05_fence_after();
        const uint64_t dq0 = desc_kmajor(sOp + oslot * OP_BYTES + OP_Q, 0);
        const uint64_t dn0 = desc_kmajor(sN1 + qslot * N1_BYTES, 0);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          PROF_T0();
          mbar_wait(bar(B_F2EMPTY + fslot), fph ^ 1u);
          PROF_ADD(2);
          PROF_T0();
          tcgen05_fence_after();
          TRACE(2, uq + sub);
          for (int t2 = 0; t2 < nt2; ++t2)
            umma_bf16_ss(tmem_base + TM_F2 + fslot * 32 + t2 * 16, dn0 + (uint64_t)((sub * 32 + t2 * 16384) >> 4),
                         dq0 + (uint64_t)((sub * 32) >> 4), id_f2, 0u);
          umma_commit(bar(B_F2FULL + fslot));
          if (++fslot == F2_RING) { fslot = 0; fph ^= 1u; }
          PROF_ADD(3);
        }
        umma_commit(bar(B_N1EMPTY + qslot));     // the quad tile is free once its 4th rank is read
        if (++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
      }
      PROF_FLUSH(2);
    }
  } else if (warp == 2) {
    // ------------------------------ III issuer: L += Vc_r . M  (this warp also owns TMEM).  Per quad: one wait, four
    // MMAs, one commit that frees the M quad slot and one that frees the operand slot --------
    if (elect_one_sync()) {
      const uint32_t id_3 = make_idesc_rt(128, p.N, 0, 1);
      PROF_DECL
      uint32_t oslot = 0, oph = 0;
      const int qpr = p.R >> 2;                   // quads per sample
      int qi = 0, sl = 0;
      for (int uq = 0; uq < U; uq += 4) {
        const int qslot = (uq >> 2) & 1, qph = (uq >> 3) & 1;
        PROF_T0();
        mbar_wait(bar(B_OPFULL + oslot), oph);
        PROF_ADD(0);
        PROF_T0();
        mbar_wait(bar(B_MFULL + qslot), qph);
        PROF_ADD(1);
        PROF_T0();
        if (qi == 0) mbar_wait(bar(B_ACCEMPTY), (sl & 1) ^ 1);
        PROF_ADD(2);
        PROF_T0();
        tcgen05_fence_after();
        const uint64_t dv0 = desc_kmajor(sOp + oslot * OP_BYTES + OP_V, 0);
        const uint64_t dm0 = desc_mnmajor(sM + qslot * 4 * M_BYTES, 0, 2048);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          TRACE(3, uq + sub);
          umma_bf16_ss(tmem_base + TM_ACC, dv0 + (uint64_t)((sub * 32) >> 4), dm0 + (uint64_t)((sub * M_BYTES) >> 4), id_3,
                       (qi > 0 || sub > 0) ? 1u : 0u);
        }
        umma_commit(bar(B_MEMPTY + qslot));
        umma_commit(bar(B_OPEMPTY + oslot));
        if (qi == qpr - 1) umma_commit(bar(B_ACCFULL));
        if (++qi == qpr) { qi = 0; ++sl; }
        if (++oslot == OP_RING) { oslot = 0; oph ^= 1u; }
        PROF_ADD(3);
      }
      PROF_FLUSH(3);
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------ C1: N1^T (TMEM) -> N1 tile rows (a,g,i), columns j.  Group 0 (warps 4-7) takes
    // the even units, group 1 (warps 8-11) the odd ones.  Lane order of the core (tpack_perm): accumulator lane L of
    // tile t holds x = (j = 4 (L & 3) + t, i = (L >> 2) & 15, g = L >> 6): the four tiles give one thread four
    // consecutive j of row (a,g,i) -- one 8-byte store per a; a warp's 32 stores cover all 32 banks exactly twice.
    const int grp = (warp - 4) >> 2;
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int jj = L & 3, i = (L >> 2) & 15, g = L >> 6;
    const uint32_t row_off = (g * 2 + (i >> 3)) * 1024u + (i & 7) * 128u;
    PROF_DECL
    for (int u = grp; u < U; u += 2) {
      const int fslot = u % F1_RING, fph = (u / F1_RING) & 1;
      const int nslot = u & (N1_RING - 1), nph = (u >> 3) & 1;
      PROF_T0();
      mbar_wait(bar(B_F1FULL + fslot), fph);
      PROF_ADD(0);
      tcgen05_fence_after();
      // the core slot of this unit: every F1 MMA that read it has retired (F1FULL counts both issuers' commits)
      if (qd == 0 && lane == 0) mbar_arrive(bar(B_TEMPTY + u % T_RING));
      PROF_T0();
      if ((nslot & 3) < 2) mbar_wait(bar(B_N1EMPTY + (nslot >> 2)), nph ^ 1u);   // first unit of this group in the quad tile
      PROF_ADD(1);
      PROF_T0();
      if (qd == 0) TRACE(4, u);
      const uint32_t n1 = sN1 + (nslot >> 2) * N1_BYTES;
      const uint32_t sub = nslot & 3;                          // rank within the quad tile: columns sub * 16 + j
      uint32_t v[4][8];
#pragma unroll
      for (int t = 0; t < 4; ++t) tmem_ld_32x32b_x8(tmem_base + lane_addr + TM_F1 + fslot * 64 + t * 16, v[t]);
      tmem_wait_ld();
      tcgen05_fence_before();
      const uint32_t off = row_off + (((sub * 2 + (jj >> 1)) ^ (i & 7)) << 4) + (jj & 1) * 8u;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        if (a < p.A)
          st_shared_v2(n1 + a * 4096u + off, pack_bf16x2(__uint_as_float(v[0][a]), __uint_as_float(v[1][a])),
                       pack_bf16x2(__uint_as_float(v[2][a]), __uint_as_float(v[3][a])));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if ((nslot & 3) >= 2) mbar_arrive(bar(B_N1QFULL + (nslot >> 2)));    // this group's last unit of the quad tile
        mbar_arrive(bar(B_F1EMPTY + fslot));
      }
      if (qd == 0) TRACE(6, u);
      PROF_ADD(2);
    }
    if (warp == 4) PROF_FLUSH(4);
  } else if (warp >= 12 && warp < 16) {
    // ------------------------------ C2: M (TMEM) -> M tile [i][(a,g,q16)] ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    uint32_t fslot = 0, fph = 0;
    PROF_DECL
    for (int u = 0; u < U; ++u) {
      const int mslot = u & 7, qslot = (u >> 2) & 1, qph = (u >> 3) & 1;
      PROF_T0();
      mbar_wait(bar(B_F2FULL + fslot), fph);
      PROF_ADD(0);
      tcgen05_fence_after();
      PROF_T0();
      if ((u & 3) == 0) mbar_wait(bar(B_MEMPTY + qslot), qph ^ 1u);
      PROF_ADD(1);
      PROF_T0();
      if (qd == 0) TRACE(5, u);
      const uint32_t mt = sM + mslot * M_BYTES;
      uint32_t v[2][16];
      tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + fslot * 32, v[0]);
      if (nt2 > 1) tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_F2 + fslot * 32 + 16, v[1]);
      tmem_wait_ld();
      tcgen05_fence_before();
#pragma unroll
      for (int t2 = 0; t2 < 2; ++t2) {
        const int rho = t2 * 128 + L;
        const int a = rho >> 5, i = rho & 15;
        if (t2 < nt2 && a < p.A) {
          const int ag = rho >> 4;                        // a*2 + g
          uint32_t pk[8];
#pragma unroll
          for (int x = 0; x < 8; ++x) pk[x] = pack_bf16x2(__uint_as_float(v[t2][2 * x]), __uint_as_float(v[t2][2 * x + 1]));
          const uint32_t tile = mt + (ag >> 2) * 2048u;
          const uint32_t c0 = (ag & 3) * 16;
          st_shared_v4(tile + sw128_off(i, c0), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(tile + sw128_off(i, c0 + 8), pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if ((u & 3) == 3) mbar_arrive(bar(B_MFULL + qslot));
        mbar_arrive(bar(B_F2EMPTY + fslot));
      }
      if (qd == 0) TRACE(7, u);
      if (++fslot == F2_RING) { fslot = 0; fph ^= 1u; }
      PROF_ADD(2);
    }
    if (warp == 12) PROF_FLUSH(5);
  } else if (warp >= 16 && warp < 20) {
    // ------------------------------ per-sample epilogue: TMEM lane = region k, column (a,g,q16) -> (G,K,Q,A) order in a
    // staging buffer whose row pitch is odd (the 32 lanes of a warp are 32 regions: 32 distinct banks), mask,
    // coalesced (B,G,K,Q,A) store.  The accumulator is released as soon as it sits in shared memory. ------
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    const int QA = p.Q * p.A;
    PROF_DECL
    for (int sl = 0; sl < n_my; ++sl) {
      const int b = blockIdx.x + sl * gridDim.x;
      PROF_T0();
      mbar_wait(bar(B_ACCFULL), sl & 1);
      PROF_ADD(0);
      PROF_T0();
      tcgen05_fence_after();
      const int k = L;
      const bool masked = (k < p.K) && p.rowmask != nullptr && p.rowmask[(size_t)(b / p.VR) * p.K + k] != 0;
      for (int a = 0; a < p.A; ++a) {
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_ACC + (2 * a) * 16, v0);
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_ACC + (2 * a + 1) * 16, v1);
        tmem_wait_ld();
        if (k < p.K) {
          float* d0 = out_stage + (size_t)k * pad + a;                       // g = 0
          float* d1 = d0 + (size_t)p.K * pad;                                // g = 1
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < p.Q) {
              d0[q * p.A] = masked ? -INFINITY : __uint_as_float(v0[q]);
              d1[q * p.A] = masked ? -INFINITY : __uint_as_float(v1[q]);
            }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ACCEMPTY));
      named_bar_sync(1, 128);
      // rows (g,k) of Q*A floats: warp w of the group copies rows w, w+4, ...
      float* gdst = p.logits + (size_t)b * 2 * p.K * QA;
      for (int row = qd; row < 2 * p.K; row += 4) {
        const float* src = out_stage + (size_t)row * pad;
        float* dst = gdst + (size_t)row * QA;
        for (int e = lane; e < QA; e += 32) dst[e] = src[e];
      }
      named_bar_sync(1, 128);
      PROF_ADD(1);
    }
    if (warp == 16) PROF_FLUSH(6);
  } else if (warp == 20) {
    // ------------------------------ N1 store (training): quad tile -> HBM as one bulk copy; the tile is released to the
    // converters once the copy has READ it (and F2 has, through its own commit) ------
    if (p.n1_save != nullptr && elect_one_sync()) {
      const uint32_t bytes = static_cast<uint32_t>(p.A) * 4096u;
      const int nq = p.R >> 2;
      int qi = 0, b = blockIdx.x;
      for (int uq = 0; uq < U; uq += 4) {
        const int qslot = (uq >> 2) & 1, qph = (uq >> 3) & 1;
        mbar_wait(bar(B_N1QFULL + qslot), qph);
        bulk_store_1d(p.n1_save + ((size_t)b * nq + qi) * bytes, sN1 + qslot * N1_BYTES, bytes);
        bulk_commit_group();
        bulk_wait_group_read<0>();
        mbar_arrive(bar(B_N1EMPTY + qslot));
        if (++qi == nq) { qi = 0; b += gridDim.x; }
      }
      bulk_wait_group_all();
    }
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
  if (n == 8 * 256) return (int)cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(unsigned long long) * 8 * 256);
  return (int)cudaMemcpyFromSymbol(host_dst, g_prof, sizeof(unsigned long long) * (n < 148 * 64 ? n : 148 * 64));
#else
  (void)host_dst; (void)n;
  return -1;
#endif
}

// Returns -100 when the shape is outside the fast path (caller falls back to the generic kernel).
// tpack_perm: the packed core with the (i,g,j) axis of every [r][l] row in the lane order x = (j % 4) * 128 + g * 64 +
// i * 4 + j / 4 (functions.tpack_perm_index).
size_t trilinear_n1_bytes(TriDims d) {        // 0: shape outside the fast path (nothing to save)
  if (d.G != 2 || d.K > 64 || d.Q > 16 || d.A > 6 || (d.R & 3) != 0 || d.R > 32) return 0;
  return (size_t)d.B * (d.R >> 2) * d.A * 4096;
}

int trilinear_fwd_tc(const bf16* vc, const bf16* qc, const bf16* ac, const bf16* tpack_perm, const uint8_t* rowmask,
                     float* logits, void* n1_save, TriDims d, cudaStream_t stream) {
  if (tpack_perm == nullptr) return -100;
  if (d.G != 2 || d.K > 64 || d.Q > 16 || d.A > 6 || (d.R & 3) != 0) return -100;      // TMEM: 320 + 32 A columns
  const size_t smem = tri_tc_smem(d.K, d.Q, d.A);
  if (smem > 227 * 1024) return -100;
  const int RD = d.R * 16;
  CUtensorMap tt, tv, tq, ta;
  if (int rc = make_tmap_3d(&tt, tpack_perm, 64, (uint64_t)d.R * 16, 8, 512, 64, 64, 16, true, 8)) return rc;
  if (int rc = make_tmap_3d(&tv, vc, RD, d.K, d.B / d.VR, RD, (uint64_t)d.K * RD, 64, 64)) return rc;
  if (int rc = make_tmap_3d(&tq, qc, RD, d.Q, d.B, RD, (uint64_t)d.Q * RD, 64, 16)) return rc;
  if (int rc = make_tmap_3d(&ta, ac, RD, d.A, d.B, RD, (uint64_t)d.A * RD, 64, 16)) return rc;
  TriTcParams p{rowmask, logits, static_cast<uint8_t*>(n1_save), d.B, d.K, d.Q, d.A, d.R, 32 * d.A, d.VR};
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
