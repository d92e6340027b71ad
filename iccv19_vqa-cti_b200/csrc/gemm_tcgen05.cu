// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared memory (128B swizzle)
// -> tcgen05.mma (UMMA 128 x BLOCK_N x 16, cta_group::1) -> fp32 accumulators in TMEM
// -> epilogue warps (tcgen05.ld) -> bias / ReLU / ReLU-mask / split-K atomics -> global.
//
// One kernel covers every dense projection on the CTI hot path:
//   forward   Y  = act(X W_eff^T + b)          A = X   (K-major), B = W_eff (K-major)
//   dgrad     dX = dZ W_eff                    A = dZ  (K-major), B = W_eff (MN-major)
//   wgrad     dW = dZ^T X   (split over rows)  A = dZ  (MN-major), B = X    (MN-major)
// which is the work of reference src/fc.py:33-34 (FCNet.forward) and its autograd.
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter).
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>

namespace cti {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 384;      // warps 0-2: TMA / MMA / TMEM alloc, warp 3 idle, warps 4-11: epilogue
constexpr int kEpiWarps = 8;
constexpr int kEpiWarp0 = 4;

// CTAS = 2: a CTA PAIR (cluster of two, cta_group::2) works on one 256 x BLOCK_N tile: each CTA stages its own 128 rows
// of A and HALF of the B tile, the leader issues tcgen05.mma M = 256 for both, each CTA's TMEM gets its 128 rows.
// Half the B-operand shared-memory traffic per SM and 6 instead of 4 pipeline stages in the same 192 KB.
template <int BLOCK_N, int CTAS = 1>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = BLOCK_N / CTAS;                  // B rows (N) staged by one CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256 && CTAS == 1) ? 4 : 6;
  static constexpr int ACC_STAGES = 2;
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;            // 512 or 256 (power of two)
  static constexpr int OUT_BYTES = kEpiWarps * 4096;  // per epilogue warp: one 32-row x 128-byte staging box
  static constexpr int BAR_BYTES = (2 * STAGES + 2 * ACC_STAGES) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + 1024;   // +1024: manual alignment
};

struct GemmDevParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks, k_splits, k_blocks_per_split;
  const float* bias;            // [N] or null
  const __nv_bfloat16* relu_aux;   // [M, ld_aux] or null: out *= (aux > 0)
  __nv_bfloat16* out_bf16;      // [M, ldc] or null
  float* out_f32;               // [M, ldc] or null
  int ldc, ld_aux;
  int relu;                     // apply max(0, .) after bias
  int atomic_f32;               // out_f32 += acc (split-K reduction)
  float alpha;                  // acc scale applied before bias
  int tma_out;                  // 0: direct stores; 1: bf16 via TMA store; 2: fp32 via TMA store; 3: fp32 TMA reduce-add
};

enum { kOutDirect = 0, kOutTmaBf16 = 1, kOutTmaF32 = 2, kOutTmaAddF32 = 3 };

// ---- CTA-pair (cta_group::2) forms of the PTX wrappers ------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's CTA 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (executed by both CTAs of the pair)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once the pair's MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {      // arrive on CTA 0's copy of the barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tmem_alloc_n(uint32_t smem_result_addr, uint32_t ncols) {      // whole warp (of each CTA)
  if (CTAS == 1) {
    tmem_alloc(smem_result_addr, ncols);
    tmem_relinquish();
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void tmem_dealloc_n(uint32_t taddr, uint32_t ncols) {
  if (CTAS == 1) tmem_dealloc(taddr, ncols);
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// One launch runs ONE problem or a PAIR of independent problems that share the operand layouts and the tile width
// (gemm_bf16_pair: the question-side and answer-side projections of a step -- same weights shape, different row
// counts -- which on their own leave SMs idle and pay the launch / pipeline-fill floor twice).  The tiles of problem 0
// come first in the persistent tile loop, then those of problem 1.
struct GemmMaps {
  CUtensorMap a, b, c;
};

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ GemmMaps maps0, const __grid_constant__ GemmMaps maps1,
                 const __grid_constant__ GemmDevParams prm0, const __grid_constant__ GemmDevParams prm1, const int tiles0,
                 const int total_tiles) {
  using Cfg = GemmCfg<BLOCK_N, CTAS>;
  // CTA pair: `unit` = the pair, `rank` = this CTA's half (rows rank * 128 .. of the 256-row tile, columns
  // rank * BLOCK_N / 2 .. of the B tile); a single CTA is a unit of its own
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int unit = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = BLOCK_M * CTAS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_out = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;          // 1024-byte aligned
  const uint32_t bar_base = smem_out + Cfg::OUT_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + Cfg::ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 2 * Cfg::ACC_STAGES);
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps0.a);
    tma_prefetch_desc(&maps0.b);
    if (prm0.tma_out != kOutDirect) tma_prefetch_desc(&maps0.c);
    if (total_tiles > tiles0) {
      tma_prefetch_desc(&maps1.a);
      tma_prefetch_desc(&maps1.b);
      if (prm1.tma_out != kOutDirect) tma_prefetch_desc(&maps1.c);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < Cfg::ACC_STAGES; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps * 32 * CTAS);      // the leader's copy hears from both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_n<CTAS>(tmem_slot, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();      // the peer's barriers exist before anything arrives on them
  else __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_prologue_done();      // everything above touched only this CTA's shared memory / TMEM

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      auto load = [&](const CUtensorMap* m, uint32_t dst, int32_t c0, int32_t c1) {
        if (CTAS == 2) tma_load_2d_pair(m, full_bar(stage), dst, c0, c1);      // bytes counted on the leader's barrier
        else tma_load_2d(m, full_bar(stage), dst, c0, c1);
      };
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const bool second = tile >= tiles0;
        const GemmDevParams& p = second ? prm1 : prm0;
        const CUtensorMap* tmap_a = second ? &maps1.a : &maps0.a;
        const CUtensorMap* tmap_b = second ? &maps1.b : &maps0.b;
        const int t = second ? tile - tiles0 : tile;
        const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
        const int ks = t / tiles_mn;
        const int mn = t - ks * tiles_mn;
        const int m_blk = mn / p.num_n_blocks;
        const int n_blk = mn - m_blk * p.num_n_blocks;
        const int kb0 = ks * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
        const int m0 = m_blk * TILE_M + (int)rank * BLOCK_M;               // this CTA's rows of the tile
        const int n0 = n_blk * BLOCK_N + (int)rank * Cfg::B_ROWS;          // this CTA's share of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CTAS);
          if (!A_MN) {
            load(tmap_a, smem_a(stage), kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) load(tmap_a, smem_a(stage) + i * 8192, m0 + i * 64, kb * BLOCK_K);
          }
          if (!B_MN) {
            load(tmap_b, smem_b(stage), kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::B_ROWS / 64; ++i) load(tmap_b, smem_b(stage) + i * 8192, n0 + i * 64, kb * BLOCK_K);
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (rank == 0 && elect_one_sync()) {      // the leader issues for the pair
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BLOCK_N, A_MN, B_MN);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accum) {
        if (CTAS == 2) umma_bf16_ss_pair(d, da, db, idesc, accum);
        else umma_bf16_ss(d, da, db, idesc, accum);
      };
      auto commit = [&](uint32_t b) {
        if (CTAS == 2) umma_commit_pair(b);
        else umma_commit(b);
      };
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const bool second = tile >= tiles0;
        const GemmDevParams& p = second ? prm1 : prm0;
        const int ks = (second ? tile - tiles0 : tile) / (p.num_m_blocks * p.num_n_blocks);
        const int kb0 = ks * p.k_blocks_per_split;
        const int kb1 = min(kb0 + p.k_blocks_per_split, p.num_k_blocks);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: advance 16 elements = 32 B inside the swizzle row.
            // MN-major: advance 16 K-rows = 2 swizzle atoms = 2048 B.
            const uint64_t da = A_MN ? make_smem_desc_sw128(smem_a(stage) + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(smem_a(stage) + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(smem_b(stage) + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(smem_b(stage) + k * 32, 16, 1024);
            mma(d_tmem, da, db, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          commit(empty_bar(stage));               // frees the smem slot (in both CTAs) when the MMAs retire
          if (kb == kb1 - 1) commit(tfull_bar(acc));   // accumulator complete
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == Cfg::ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------ epilogue ----------------------------------
    // Thread = one accumulator row (TMEM lane), 32 columns per tcgen05.ld.  The next chunk's TMEM load is in flight
    // while the current one is processed; results are staged in this warp's swizzled 32-row x 128-byte boxes and
    // leave through TMA (store, or reduce-add for split-K), so global writes are full lines and asynchronous.
    // Two warps per TMEM lane quarter, each taking half of the tile's columns: the per-thread work is a long dependent
    // chain (tcgen05.ld -> shuffle -> fma -> pack -> st.shared), so two warps per scheduler hide each other's latency.
    constexpr int NCH = BLOCK_N / 64;             // 32-column chunks per warp
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const int chalf = (warp - kEpiWarp0) >> 2;    // which half of the tile's columns
    const uint32_t stage0 = smem_out + (warp - kEpiWarp0) * 4096;
    const uint32_t my_row = stage0 + lane * 128;  // this thread's 128-byte row inside the staging box
    const uint32_t sw = lane & 7;
    constexpr uint32_t obuf = 0;
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units) {
      const bool second = tile >= tiles0;
      const GemmDevParams& p = second ? prm1 : prm0;
      const CUtensorMap* tmap_c = second ? &maps1.c : &maps0.c;
      const int t = second ? tile - tiles0 : tile;
      const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
      const int ks = t / tiles_mn;
      const int mn = t - ks * tiles_mn;
      const int m_blk = mn / p.num_n_blocks;
      const int n_blk = mn - m_blk * p.num_n_blocks;
      const bool empty_split = (ks * p.k_blocks_per_split >= p.num_k_blocks);
      const int row0 = m_blk * TILE_M + (int)rank * BLOCK_M + quarter * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int n0 = n_blk * BLOCK_N + chalf * (BLOCK_N / 2);
      // bias: lane l holds column (chunk base + l); the first two chunks' loads are issued before the accumulator is
      // ready, the following ones one iteration ahead (the chunk loop stays rolled: unrolled it overflows the i-cache)
      auto load_bias = [&](int c) {
        const int col = n0 + c * 32 + lane;
        return (p.bias != nullptr && c < NCH && col < p.N) ? __ldg(p.bias + col) : 0.f;
      };
      float b0 = load_bias(0), b1 = load_bias(1);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + chalf * (BLOCK_N / 2);

      auto process = [&](const int c, uint32_t (&r)[32], const float bias_l) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N || empty_split) return;               // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), p.alpha, __shfl_sync(0xffffffffu, bias_l, j));
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const bool full = (col0 + 32 <= p.N);
        if (p.relu_aux != nullptr && row_ok) {
          const __nv_bfloat16* aux = p.relu_aux + static_cast<size_t>(row) * p.ld_aux + col0;
          if (full && (p.ld_aux % 8 == 0)) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const uint4 a = __ldg(reinterpret_cast<const uint4*>(aux) + j8);
              const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = unpack_bf16x2(w[t]);
                if (!(f.x > 0.f)) v[j8 * 8 + 2 * t] = 0.f;
                if (!(f.y > 0.f)) v[j8 * 8 + 2 * t + 1] = 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N && !(__bfloat162float(aux[j]) > 0.f)) v[j] = 0.f;
          }
        }
        const size_t off = static_cast<size_t>(row) * p.ldc + col0;
        // ---- bf16 output
        if (p.out_bf16 != nullptr) {
          if (p.tma_out == kOutTmaBf16) {
            // 64-column slab = two chunks = one 128-byte row per thread; 16-byte piece q sits at q ^ (row & 7)
            const uint32_t half = static_cast<uint32_t>(c & 1) * 4u;
            if (half == 0) {
              if (lane == 0) bulk_wait_group_read<0>();        // the previous box has been read out
              __syncwarp();
            }
            const uint32_t dst = my_row + obuf * 4096;
#pragma unroll
            for (uint32_t q4 = 0; q4 < 4; ++q4)
              st_shared_v4(dst + (((half + q4) ^ sw) << 4), pack_bf16x2(v[q4 * 8 + 0], v[q4 * 8 + 1]),
                           pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]), pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]),
                           pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]));
            if (half != 0 || col0 + 32 >= p.N) {              // slab complete (or the row of tiles ends here)
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(tmap_c, stage0 + obuf * 4096, n0 + (c >> 1) * 64, row0);
                bulk_commit_group();
              }
            }
          } else if (row_ok) {
            __nv_bfloat16* o = p.out_bf16 + off;
            if (full && (p.ldc % 8 == 0)) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint4 u;
                u.x = pack_bf16x2(v[j8 * 8 + 0], v[j8 * 8 + 1]);
                u.y = pack_bf16x2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
                u.z = pack_bf16x2(v[j8 * 8 + 4], v[j8 * 8 + 5]);
                u.w = pack_bf16x2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
                reinterpret_cast<uint4*>(o)[j8] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16(v[j]);
            }
          }
        }
        // ---- fp32 output (plain, or accumulated over the K splits)
        if (p.out_f32 != nullptr) {
          if (p.tma_out >= kOutTmaF32) {
            if (lane == 0) bulk_wait_group_read<0>();
            __syncwarp();
            const uint32_t dst = my_row + obuf * 4096;
#pragma unroll
            for (uint32_t q8 = 0; q8 < 8; ++q8)
              st_shared_v4(dst + ((q8 ^ sw) << 4), __float_as_uint(v[q8 * 4]), __float_as_uint(v[q8 * 4 + 1]),
                           __float_as_uint(v[q8 * 4 + 2]), __float_as_uint(v[q8 * 4 + 3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.tma_out == kOutTmaAddF32) tma_reduce_add_2d(tmap_c, stage0 + obuf * 4096, col0, row0);
              else                            tma_store_2d(tmap_c, stage0 + obuf * 4096, col0, row0);
              bulk_commit_group();
            }
          } else if (row_ok) {
            float* o = p.out_f32 + off;
            if (p.atomic_f32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(o + j, v[j]);
            } else if (full && (p.ldc % 4 == 0)) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                reinterpret_cast<float4*>(o)[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = v[j];
            }
          }
        }
      };

      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(t_addr, ra);
#pragma unroll 1
      for (int c = 0; c < NCH; c += 2) {
        const float nb0 = load_bias(c + 2), nb1 = load_bias(c + 3);
        tmem_wait_ld_regs(ra);
        tmem_ld_32x32b_x32(t_addr + (c + 1) * 32, rb);
        process(c, ra, b0);
        tmem_wait_ld_regs(rb);
        if (c + 2 < NCH) {
          tmem_ld_32x32b_x32(t_addr + (c + 2) * 32, ra);
        } else {
          tcgen05_fence_before();                 // every TMEM read of this tile has completed
          if (CTAS == 2) mbar_arrive_leader(tempty_bar(acc));      // the issuer lives in CTA 0
          else mbar_arrive(tempty_bar(acc));
        }
        process(c + 1, rb, b1);
        b0 = nb0;
        b1 = nb1;
      }
      if (++acc == Cfg::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_group_all();         // staged boxes must be drained before the CTA's smem goes away
  }

  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all();      // neither CTA leaves (or frees TMEM) while the peer may still touch it
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_n<CTAS>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ----------------------------- host side ---------------------------------- //
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch ld elements.
int make_tmap(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
              uint32_t box_outer, bool f32 = false) {
  auto fn = get_encode_fn();
  CTI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CTI_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  const uint64_t esz = f32 ? 4 : 2;
  CTI_REQUIRE((ld_elems * esz) % 16 == 0, "TMA operand row pitch must be a multiple of 16 bytes (ld=%llu)",
              (unsigned long long)ld_elems);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld_elems * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// Tensor maps + device parameters of one problem for the given tile width / layouts.
template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS>
int build_problem(const GemmArgs& g, GemmMaps* maps, GemmDevParams* out) {
  int rc;
  if (!A_MN) rc = make_tmap(&maps->a, g.a, g.K, g.M, g.lda, BLOCK_K, BLOCK_M);
  else       rc = make_tmap(&maps->a, g.a, g.M, g.K, g.lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&maps->b, g.b, g.K, g.N, g.ldb, BLOCK_K, BLOCK_N / CTAS);      // a pair's CTA stages half of B
  else       rc = make_tmap(&maps->b, g.b, g.N, g.K, g.ldb, 64, BLOCK_K);
  if (rc) return rc;

  GemmDevParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.num_m_blocks = (g.M + BLOCK_M * CTAS - 1) / (BLOCK_M * CTAS);
  p.num_n_blocks = (g.N + BLOCK_N - 1) / BLOCK_N;
  p.num_k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  int splits = g.k_splits < 1 ? 1 : g.k_splits;
  if (splits > p.num_k_blocks) splits = p.num_k_blocks;
  p.k_blocks_per_split = (p.num_k_blocks + splits - 1) / splits;
  p.k_splits = (p.num_k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;   // no empty splits
  p.bias = g.bias; p.relu_aux = g.relu_aux; p.out_bf16 = g.out_bf16; p.out_f32 = g.out_f32;
  p.ldc = g.ldc; p.ld_aux = g.ld_aux; p.relu = g.relu; p.atomic_f32 = g.atomic_f32; p.alpha = g.alpha;
  // Output through TMA when the buffer allows it (16-byte aligned base and row pitch); with both outputs requested
  // the fp32 one takes the TMA path.  Each epilogue warp stores 32-row boxes of 128 bytes (64 bf16 / 32 fp32).
  auto tma_ok = [&](const void* ptr, int esz) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (static_cast<long>(g.ldc) * esz) % 16 == 0;
  };
  maps->c = maps->a;
  p.tma_out = kOutDirect;
  if (g.out_f32 != nullptr && tma_ok(g.out_f32, 4)) {
    p.tma_out = g.atomic_f32 ? kOutTmaAddF32 : kOutTmaF32;
    if ((rc = make_tmap(&maps->c, g.out_f32, g.N, g.M, g.ldc, 32, 32, true))) return rc;
  } else if (g.out_bf16 != nullptr && g.out_f32 == nullptr && tma_ok(g.out_bf16, 2)) {
    p.tma_out = kOutTmaBf16;
    if ((rc = make_tmap(&maps->c, g.out_bf16, g.N, g.M, g.ldc, 64, 32))) return rc;
  }
  *out = p;
  return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int CTAS = 1>
int launch_gemm(const GemmArgs& g0, const GemmArgs* g1, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, CTAS>;
  GemmMaps m0, m1;
  GemmDevParams p0, p1;
  if (int rc = build_problem<BLOCK_N, A_MN, B_MN, CTAS>(g0, &m0, &p0)) return rc;
  const long t0 = (long)p0.num_m_blocks * p0.num_n_blocks * p0.k_splits;
  long t1 = 0;
  if (g1 != nullptr) {
    if (int rc = build_problem<BLOCK_N, A_MN, B_MN, CTAS>(*g1, &m1, &p1)) return rc;
    t1 = (long)p1.num_m_blocks * p1.num_n_blocks * p1.k_splits;
  } else {
    m1 = m0;
    p1 = p0;
  }
  CTI_REQUIRE(t0 + t1 < (1L << 30), "gemm: too many tiles");

  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, CTAS>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const long total = t0 + t1;
  int sms = g0.max_ctas > 0 ? g0.max_ctas : kNumSMsB200;
  if (CTAS == 2) {
    // one CTA pair per TPC: cluster of 2 + programmatic dependent launch
    const int pairs = (int)(total < sms / 2 ? total : sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, kern, m0, m1, p0, p1, (int)t0, (int)total);
    return check_launch("gemm_bf16_kernel (CTA pairs)");
  }
  const int grid = (int)(total < sms ? total : sms);
  launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::SMEM_BYTES, stream, m0, m1, p0, p1, (int)t0, (int)total);
  return check_launch("gemm_bf16_kernel");
}

int check_gemm_args(const GemmArgs& g) {
  CTI_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  CTI_REQUIRE(g.out_bf16 != nullptr || g.out_f32 != nullptr, "gemm: no output buffer");
  CTI_REQUIRE(!g.atomic_f32 || g.out_f32 != nullptr, "gemm: split-K accumulation needs an fp32 output");
  CTI_REQUIRE(g.atomic_f32 || g.k_splits <= 1, "gemm: k_splits > 1 needs atomic_f32");
  return 0;
}

long tiles256_of(const GemmArgs& g) {
  return (long)((g.M + BLOCK_M - 1) / BLOCK_M) * ((g.N + 255) / 256) * (g.k_splits < 1 ? 1 : g.k_splits);
}

int dispatch_gemm(const GemmArgs& g0, const GemmArgs* g1, cudaStream_t stream) {
  // Small-N problems and very small grids use the 128-wide tile so more CTAs get work.  Measured (tools/tile_sweep.py):
  // the 256-wide tile wins from ~64 tiles on (6144 x 512 x 1024: 11.3 vs 15.0 us; the GRU step 1024 x 3072 x 1024:
  // 11.4 vs 15.4 us) even though the grid no longer fills 148 SMs -- half as many B-operand bytes per FLOP; below
  // that (1024 x 1024 x 1024, 32 tiles) the narrow tile is ahead by ~1 us.  A pair is sized as one problem.
  const long tiles256 = tiles256_of(g0) + (g1 ? tiles256_of(*g1) : 0);
  const int minN = g1 ? (g0.N < g1->N ? g0.N : g1->N) : g0.N;
  const int forced = g0.tile_n != 0 ? g0.tile_n : (g1 ? g1->tile_n : 0);
  const bool use256 = (forced == 256) || (forced == 0 && minN >= 256 && tiles256 >= 64);
  // CTA pairs (256 x 256 tiles, cta_group::2) once there are enough 256-row tiles for every pair; tile_n = 512 forces
  // them, CTI_GEMM_PAIRS=0 in the environment switches them off (A/B measurements)
  static const bool pairs_on = [] { const char* e = getenv("CTI_GEMM_PAIRS"); return !(e && e[0] == '0'); }();
  auto tiles512 = [](const GemmArgs& g) {
    return (long)((g.M + 255) / 256) * ((g.N + 255) / 256) * (g.k_splits < 1 ? 1 : g.k_splits);
  };
  const long tp = tiles512(g0) + (g1 ? tiles512(*g1) : 0);
  // measured (tools/gemm_pairs_bench.py): pairs win once a tile's K loop is >= 1024 deep (51200 x 1024 x 2048: 159 -> 146 us,
  // 0.83 -> 0.90 of the burst peak; 18432 x 1024 x 1024: 31.7 -> 29.4 us) and lose a little below that and on the
  // K = 1024 dgrad layout (shorter loops do not amortise the pair's cluster-wide hand-offs)
  const int k_per_split = g0.K / (g0.k_splits < 1 ? 1 : g0.k_splits);
  const bool dgrad_layout = !g0.a_mn_major && g0.b_mn_major;
  const bool use_pairs = forced == 512 || (pairs_on && forced == 0 && use256 && tp >= 56 &&
                                           k_per_split >= (dgrad_layout ? 2048 : 1024));
  if (use_pairs) {
    if (!g0.a_mn_major && !g0.b_mn_major) return launch_gemm<256, false, false, 2>(g0, g1, stream);
    if (!g0.a_mn_major && g0.b_mn_major) return launch_gemm<256, false, true, 2>(g0, g1, stream);
    if (g0.a_mn_major && g0.b_mn_major) return launch_gemm<256, true, true, 2>(g0, g1, stream);
  }
  if (use256) {
    if (!g0.a_mn_major && !g0.b_mn_major) return launch_gemm<256, false, false>(g0, g1, stream);
    if (!g0.a_mn_major && g0.b_mn_major) return launch_gemm<256, false, true>(g0, g1, stream);
    if (g0.a_mn_major && g0.b_mn_major) return launch_gemm<256, true, true>(g0, g1, stream);
  } else {
    if (!g0.a_mn_major && !g0.b_mn_major) return launch_gemm<128, false, false>(g0, g1, stream);
    if (!g0.a_mn_major && g0.b_mn_major) return launch_gemm<128, false, true>(g0, g1, stream);
    if (g0.a_mn_major && g0.b_mn_major) return launch_gemm<128, true, true>(g0, g1, stream);
  }
  set_error("gemm: A MN-major with B K-major is not instantiated");
  return -1;
}

}  // namespace

// 3-D bf16 tensor map (d0 contiguous), 128-byte swizzle or none, box {box0, box1, 1}; out-of-range rows read as zero.
// Used by the per-sample kernels to load one sample's (tokens x channels) tile with the token rows padded.
int make_tmap_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                 uint64_t stride2_elems, uint32_t box0, uint32_t box1, bool swizzle128, uint32_t box2) {
  auto fn = get_encode_fn();
  CTI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CTI_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  CTI_REQUIRE((stride1_elems * 2) % 16 == 0 && (stride2_elems * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return 0;
}

// 4-D bf16 tensor map (128-byte swizzle): strides in elements for dims 1..3, one box extent per dim.
int make_tmap_4d(CUtensorMap* map, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&strides_elems)[3],
                 const uint32_t (&box)[4]) {
  auto fn = get_encode_fn();
  CTI_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CTI_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstride[3];
  for (int i = 0; i < 3; ++i) {
    CTI_REQUIRE((strides_elems[i] * 2) % 16 == 0, "TMA strides must be multiples of 16 bytes");
    gstride[i] = strides_elems[i] * 2;
  }
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTI_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
  return 0;
}

int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  if (int rc = check_gemm_args(g)) return rc;
  return dispatch_gemm(g, nullptr, stream);
}

// Two independent problems with the same operand layouts in one launch (see gemm_bf16_kernel).
int gemm_bf16_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t stream) {
  if (int rc = check_gemm_args(g0)) return rc;
  if (int rc = check_gemm_args(g1)) return rc;
  CTI_REQUIRE(g0.a_mn_major == g1.a_mn_major && g0.b_mn_major == g1.b_mn_major,
              "gemm pair: both problems must use the same operand layouts");
  CTI_REQUIRE(g0.tile_n == g1.tile_n || g0.tile_n == 0 || g1.tile_n == 0, "gemm pair: conflicting forced tile widths");
  return dispatch_gemm(g0, &g1, stream);
}

}  // namespace cti
