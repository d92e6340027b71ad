// Bilinear attention logits of BCNet.forward (`h_out <= 32` branch, reference src/bc.py:52-58, wrapped by
// BiAttention, src/attention.py:19-20,33) on tcgen05, forward and backward:
//
//   logits[b,g,k,q] = sum_c Vb[b,k,c] h[g,c] Qb[b,q,c] + h_bias[g]
//
// Per sample and per 128-channel chunk a builder warp group folds h into the question operand,
//   HQ[(g,q), c] = h[g,c] Qb[q,c]      (bf16 tile in the TMA swizzle layout, rows (g,q16), 128 channels)
// so the reference's (B,G,K,3072) fp32 tensor `v_ * h_mat` never exists, and
//   forward   L[k,(g,q)]  += Vb[k,c] . HQ[(g,q),c]^T          accumulated over all chunks in TMEM
//   backward  dV[k,c]      = dL[k,(g,q)] . HQ[(g,q),c]         HQ read MN-major from the same tile
//             PT[c,(g,q)]  = Vb[k,c]^T . dL[k,(g,q)]           TMEM lane = channel:
//                            dQb[q,c] = sum_g h[g,c] PT[c,(g,q)],  dh[g,c] += sum_q Qb[q,c] PT[c,(g,q)]
//                            are per-thread dot products (no cross-lane reduction)
// Vb and Qb chunks arrive by TMA (one pipeline stage = V tile + Q tile); outputs are pre-activation gradients
// (ReLU masks of Vb / Qb applied).  The bias gradient of the v projection is a column sum of dzv taken afterwards
// by act_bwd_bias (HBM bound, 25 us at B = 256).
//
// Roles: warp 0 TMA | warp 1 MMA issuer | warp 2 TMEM | warps 4-7 epilogue (forward: logits; backward: dV) |
// warps 8-11 HQ builder (+ dL tile, backward) | warps 12-15 PT epilogue (backward).
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

namespace cti {

namespace {

using bf16 = __nv_bfloat16;

constexpr int CCH = 128, KP = 64;
constexpr int ST_V = 0, ST_Q = 2 * KP * 128, ST_BYTES = ST_Q + 16 * CCH * 2;        // 16 KB V (swizzled) + 4 KB Qb (plain)
constexpr int STAGES = 4;
constexpr int HQ_BYTES = 2 * 64 * 128;          // two 64-channel tiles of up to 64 (g,q16) rows
constexpr int DL_BYTES = KP * 128;              // dL tile [64 k][64 (g,q16)]
constexpr int MAX_C = 3072, MAX_G = 4;

enum { E_SFULL = 0, E_SEMPTY = 4, E_HQFULL = 8, E_HQEMPTY = 10, E_ACCFULL = 12, E_ACCEMPTY = 14, E_DLFULL = 16, E_DLEMPTY = 18,
       E_DVFULL = 20, E_DVEMPTY = 22, E_PTFULL = 24, E_PTEMPTY = 26, E_COUNT = 28 };

struct BiTcParams {
  const float* hmat;      // (G, C)
  const float* hbias;     // (G)
  const uint8_t* rowmask;
  float* logits;          // forward (B,G,K,Q)
  const float* dlogits;   // backward
  bf16 *dzv, *dzq;
  float *dbq, *dhmat, *dhbias;
  int B, K, Q, G, C, NR, nchunks;
};

template <bool BWD>
__host__ __device__ constexpr size_t bi_tc_smem() {
  return (size_t)STAGES * ST_BYTES + 2 * HQ_BYTES + (BWD ? 2 * DL_BYTES + (MAX_G + 1) * MAX_C * 4 : 0) + E_COUNT * 8 + 16 + 1024;
}

template <bool BWD>
__global__ void __launch_bounds__(BWD ? 512 : 384, 1)
bilinear_tc_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q,
                   const BiTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sSt = base;
  const uint32_t sHQ = sSt + STAGES * ST_BYTES;
  const uint32_t sDL = sHQ + 2 * HQ_BYTES;
  const uint32_t sAcc = sDL + (BWD ? 2 * DL_BYTES : 0);
  const uint32_t sBar = sAcc + (BWD ? (MAX_G + 1) * MAX_C * 4 : 0);
  const uint32_t tmem_slot = sBar + E_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };
  float* acc_sm = reinterpret_cast<float*>(smem_raw + (sAcc - smem_u32(smem_raw)));     // [G][MAX_C] dh, then [MAX_C] dbq
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NR = p.NR;                          // 16 * G rows of HQ / columns of L

  if (BWD) {
    for (int i = threadIdx.x; i < (MAX_G + 1) * MAX_C; i += blockDim.x) acc_sm[i] = 0.f;
    for (uint32_t i = threadIdx.x; i < 2 * DL_BYTES / 16; i += blockDim.x) st_shared_v4(sDL + i * 16, 0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar(E_SFULL + s), 1);
      mbar_init(bar(E_SEMPTY + s), BWD ? 13 : 5);        // MMA commit + every warp that reads the stage with ld.shared
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(E_HQFULL + s), 4);
      mbar_init(bar(E_HQEMPTY + s), 1);
      mbar_init(bar(E_ACCFULL + s), 1);
      mbar_init(bar(E_ACCEMPTY + s), 4);
      mbar_init(bar(E_DLFULL + s), 4);
      mbar_init(bar(E_DLEMPTY + s), 1);
      mbar_init(bar(E_DVFULL + s), 1);
      mbar_init(bar(E_DVEMPTY + s), 4);
      mbar_init(bar(E_PTFULL + s), 1);
      mbar_init(bar(E_PTEMPTY + s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_prologue_done();      // everything above touched only this CTA's shared memory / TMEM

  const int n_my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_my * p.nchunks;
  // TMEM: forward L accumulators 2 x 64 columns at 0; backward dV 2 x 128 at 0, PT 2 x 64 at 256
  constexpr uint32_t TM_L = 0, TM_DV = 0, TM_PT = 256;

  if (warp == 0) {
    if (elect_one_sync()) {
      for (int g = 0; g < total; ++g) {
        const int b = blockIdx.x + (g / p.nchunks) * gridDim.x, ch = g % p.nchunks;
        const int st = g % STAGES;
        mbar_wait(bar(E_SEMPTY + st), ((g / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(E_SFULL + st), ST_BYTES);
        const uint32_t dst = sSt + st * ST_BYTES;
        tma_load_3d(&tmap_v, bar(E_SFULL + st), dst + ST_V, ch * CCH, 0, b);
        tma_load_3d(&tmap_v, bar(E_SFULL + st), dst + ST_V + KP * 128, ch * CCH + 64, 0, b);
        tma_load_3d(&tmap_q, bar(E_SFULL + st), dst + ST_Q, ch * CCH, 0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t id_l = make_idesc_rt(128, NR, 0, 0);
      const uint32_t id_dv = make_idesc_rt(128, 128, 0, 1);
      const uint32_t id_pt = make_idesc_rt(128, NR, 1, 1);
      const int kq = NR >> 4;                                 // K steps over the (g,q16) index
      const int ktok = (p.K + 15) >> 4;
      for (int g = 0; g < total; ++g) {
        const int sl = g / p.nchunks, ch = g % p.nchunks;
        const uint32_t st = sSt + (g % STAGES) * ST_BYTES, hq = sHQ + (g & 1) * HQ_BYTES;
        mbar_wait(bar(E_SFULL + g % STAGES), (g / STAGES) & 1);
        mbar_wait(bar(E_HQFULL + (g & 1)), (g >> 1) & 1);
        if (!BWD) {
          if (ch == 0) mbar_wait(bar(E_ACCEMPTY + (sl & 1)), ((sl >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ss(tmem_base + TM_L + (sl & 1) * 64, desc_kmajor(st + ST_V + half * (KP * 128), kk),
                           desc_kmajor(hq + half * (64 * 128), kk), id_l, (ch | half | kk) ? 1u : 0u);
          umma_commit(bar(E_SEMPTY + g % STAGES));
          umma_commit(bar(E_HQEMPTY + (g & 1)));
          if (ch == p.nchunks - 1) umma_commit(bar(E_ACCFULL + (sl & 1)));
        } else {
          const uint32_t dl = sDL + (sl & 1) * DL_BYTES;
          if (ch == 0) mbar_wait(bar(E_DLFULL + (sl & 1)), (sl >> 1) & 1);
          mbar_wait(bar(E_DVEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
          mbar_wait(bar(E_PTEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
          tcgen05_fence_after();
          // dV[k, c] = dL[k, (g,q)] . HQ[(g,q), c]        A K-major (K = columns of the dL tile), B MN-major
          for (int ks = 0; ks < kq; ++ks)
            umma_bf16_ss(tmem_base + TM_DV + (g & 1) * 128, desc_kmajor(dl, ks), desc_mnmajor(hq, ks, 64 * 128), id_dv,
                         ks > 0 ? 1u : 0u);
          umma_commit(bar(E_DVFULL + (g & 1)));
          // PT[c, (g,q)] = Vb[k, c]^T . dL[k, (g,q)]       A MN-major (M = channels), B MN-major (N = columns of dL)
          for (int ks = 0; ks < ktok; ++ks)
            umma_bf16_ss(tmem_base + TM_PT + (g & 1) * 64, desc_mnmajor(st + ST_V, ks, KP * 128), desc_mnmajor(dl, ks, 0),
                         id_pt, ks > 0 ? 1u : 0u);
          umma_commit(bar(E_PTFULL + (g & 1)));
          umma_commit(bar(E_HQEMPTY + (g & 1)));
          umma_commit(bar(E_SEMPTY + g % STAGES));
          if (ch == p.nchunks - 1) umma_commit(bar(E_DLEMPTY + (sl & 1)));
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ------------------------------ builder: HQ tile per chunk (+ dL tile per sample, backward) --------------
    const int t = threadIdx.x - 8 * 32;
    const int n_tasks = NR * 16;                     // (row (g,q16), group of 8 channels)
    for (int g = 0; g < total; ++g) {
      const int sl = g / p.nchunks, ch = g % p.nchunks;
      const int b = blockIdx.x + sl * gridDim.x;
      if (BWD && ch == 0) {
        // dL tile: row k, column g*16 + q, bf16; bias gradient = plain sums of dL
        mbar_wait(bar(E_DLEMPTY + (sl & 1)), ((sl >> 1) & 1) ^ 1);
        const uint32_t dl = sDL + (sl & 1) * DL_BYTES;
        const float* src = p.dlogits + (size_t)b * p.G * p.K * p.Q;
        const int per_g = p.K * p.Q;
        for (int gi = 0; gi < p.G; ++gi) {
          float part = 0.f;
          for (int e = t; e < per_g; e += 128) {
            const float val = __ldg(src + (size_t)gi * per_g + e);
            const int k = e / p.Q, q = e - k * p.Q;
            const __nv_bfloat16 hv = __float2bfloat16(val);
            st_shared_u16(dl + sw128_off(k, gi * 16 + q), *reinterpret_cast<const uint16_t*>(&hv));
            part += val;
          }
          part = warp_sum(part);
          if (lane == 0) atomicAdd(p.dhbias + gi, part);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(E_DLFULL + (sl & 1)));
      }
      const uint32_t st = sSt + (g % STAGES) * ST_BYTES, hq = sHQ + (g & 1) * HQ_BYTES;
      mbar_wait(bar(E_SFULL + g % STAGES), (g / STAGES) & 1);
      mbar_wait(bar(E_HQEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
      for (int task = t; task < n_tasks; task += 128) {
        const int row = task >> 4, cg = task & 15;            // row = g*16 + q, 8 channels starting at cg*8
        const int gi = row >> 4, q = row & 15;
        uint32_t qw[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(qw[0]), "=r"(qw[1]), "=r"(qw[2]), "=r"(qw[3])
                     : "r"(st + ST_Q + (q * CCH + cg * 8) * 2));
        const float4* hp = reinterpret_cast<const float4*>(p.hmat + (size_t)gi * p.C + ch * CCH + cg * 8);
        const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        const float2 a0 = unpack_bf16x2(qw[0]), a1 = unpack_bf16x2(qw[1]), a2 = unpack_bf16x2(qw[2]), a3 = unpack_bf16x2(qw[3]);
        const uint32_t tile = hq + (cg >> 3) * (64 * 128);
        st_shared_v4(tile + sw128_off(row, (cg & 7) * 8), pack_bf16x2(a0.x * h0.x, a0.y * h0.y),
                     pack_bf16x2(a1.x * h0.z, a1.y * h0.w), pack_bf16x2(a2.x * h1.x, a2.y * h1.y),
                     pack_bf16x2(a3.x * h1.z, a3.y * h1.w));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(E_HQFULL + (g & 1)));
        mbar_arrive(bar(E_SEMPTY + g % STAGES));
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int qd = warp & 3, L = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    if (!BWD) {
      // ------------------------------ forward epilogue: one (G,K,Q) tile per sample ----------------------
      for (int sl = 0; sl < n_my; ++sl) {
        const int b = blockIdx.x + sl * gridDim.x;
        mbar_wait(bar(E_ACCFULL + (sl & 1)), (sl >> 1) & 1);
        tcgen05_fence_after();
        const bool masked = (L < p.K) && p.rowmask != nullptr && p.rowmask[(size_t)b * p.K + L] != 0;
        for (int gi = 0; gi < p.G; ++gi) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_L + (sl & 1) * 64 + gi * 16, v);
          tmem_wait_ld();
          if (L < p.K) {
            const float hb = __ldg(p.hbias + gi);
            float* dst = p.logits + (((size_t)b * p.G + gi) * p.K + L) * p.Q;
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (q < p.Q) dst[q] = masked ? -INFINITY : __uint_as_float(v[q]) + hb;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(E_ACCEMPTY + (sl & 1)));
      }
    } else {
      // ------------------------------ backward epilogue 1: dV[k, 128 channels] -> ReLU mask, dzv -----------
      for (int g = 0; g < total; ++g) {
        const int b = blockIdx.x + (g / p.nchunks) * gridDim.x, ch = g % p.nchunks;
        const uint32_t st = sSt + (g % STAGES) * ST_BYTES;
        mbar_wait(bar(E_SFULL + g % STAGES), (g / STAGES) & 1);
        mbar_wait(bar(E_DVFULL + (g & 1)), (g >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {                       // 16 channels per step
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_DV + (g & 1) * 128 + j * 16, v);
          tmem_wait_ld();
          if (L < p.K) {
            const uint32_t rb = st + ST_V + (j >> 2) * (KP * 128) + (L >> 3) * 1024u + (L & 7) * 128u;
            uint32_t w0[4], w1[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0[0]), "=r"(w0[1]), "=r"(w0[2]), "=r"(w0[3])
                         : "r"(rb + (((((j & 3) * 2)) ^ (L & 7)) << 4)));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w1[0]), "=r"(w1[1]), "=r"(w1[2]), "=r"(w1[3])
                         : "r"(rb + (((((j & 3) * 2 + 1)) ^ (L & 7)) << 4)));
            uint32_t o[8];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const float2 a = unpack_bf16x2(w0[x]), c2 = unpack_bf16x2(w1[x]);
              o[x] = pack_bf16x2(a.x > 0.f ? __uint_as_float(v[2 * x]) : 0.f, a.y > 0.f ? __uint_as_float(v[2 * x + 1]) : 0.f);
              o[4 + x] = pack_bf16x2(c2.x > 0.f ? __uint_as_float(v[8 + 2 * x]) : 0.f,
                                     c2.y > 0.f ? __uint_as_float(v[8 + 2 * x + 1]) : 0.f);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.dzv + ((size_t)b * p.K + L) * p.C + ch * CCH + j * 16);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(E_DVEMPTY + (g & 1)));
          mbar_arrive(bar(E_SEMPTY + g % STAGES));
        }
      }
    }
  } else if (BWD && warp >= 12) {
    // ------------------------------ backward epilogue 2: PT[c, (g,q)] -> dQb, dh, dbq (lane = channel) ---------
    const int qd = warp & 3, cl = qd * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
    for (int g = 0; g < total; ++g) {
      const int b = blockIdx.x + (g / p.nchunks) * gridDim.x, ch = g % p.nchunks;
      const int c = ch * CCH + cl;
      const uint32_t st = sSt + (g % STAGES) * ST_BYTES;
      mbar_wait(bar(E_SFULL + g % STAGES), (g / STAGES) & 1);
      float qb[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) qb[q] = bf16_bits_to_float(ld_shared_u16(st + ST_Q + (q * CCH + cl) * 2));
      mbar_wait(bar(E_PTFULL + (g & 1)), (g >> 1) & 1);
      tcgen05_fence_after();
      float dq[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) dq[q] = 0.f;
      for (int gi = 0; gi < p.G; ++gi) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_PT + (g & 1) * 64 + gi * 16, v);
        tmem_wait_ld();
        const float hg = __ldg(p.hmat + (size_t)gi * p.C + c);
        float dh = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float pv = __uint_as_float(v[q]);
          dh = fmaf(qb[q], pv, dh);
          dq[q] = fmaf(hg, pv, dq[q]);
        }
        acc_sm[gi * MAX_C + c] += dh;                       // channel c is owned by this thread
      }
      tcgen05_fence_before();
      float sq = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (q < p.Q) {
          const float gq = qb[q] > 0.f ? dq[q] : 0.f;
          p.dzq[((size_t)b * p.Q + q) * p.C + c] = __float2bfloat16(gq);
          sq += gq;
        }
      }
      acc_sm[MAX_G * MAX_C + c] += sq;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(E_PTEMPTY + (g & 1)));
        mbar_arrive(bar(E_SEMPTY + g % STAGES));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (BWD) {
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      atomicAdd(p.dbq + c, acc_sm[MAX_G * MAX_C + c]);
      for (int gi = 0; gi < p.G; ++gi) atomicAdd(p.dhmat + (size_t)gi * p.C + c, acc_sm[gi * MAX_C + c]);
    }
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BWD>
int launch_bi(const bf16* vb, const bf16* qb, BiTcParams p, cudaStream_t stream, const char* who) {
  CUtensorMap tv, tq;
  if (int rc = make_tmap_3d(&tv, vb, p.C, p.K, p.B, p.C, (uint64_t)p.K * p.C, 64, KP)) return rc;
  if (int rc = make_tmap_3d(&tq, qb, p.C, p.Q, p.B, p.C, (uint64_t)p.Q * p.C, CCH, 16, false)) return rc;
  constexpr size_t smem = bi_tc_smem<BWD>();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bilinear_tc_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s smem attr: %s", who, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int grid = p.B < kNumSMsB200 ? p.B : kNumSMsB200;
  launch_pdl(bilinear_tc_kernel<BWD>, dim3(grid), dim3(BWD ? 512 : 384), smem, stream, tv, tq, p);
  return check_launch(who);
}

bool bi_fast_path(const BiDims& d) {
  return d.G >= 1 && d.G <= MAX_G && d.K <= KP && d.Q <= 16 && d.C % CCH == 0 && d.C <= MAX_C;
}

}  // namespace

// Both return -100 when the shape is outside the fast path (caller falls back to the generic kernels).
int bilinear_fwd_tc(const bf16* vb, const bf16* qb, const float* hmat, const float* hbias, const uint8_t* rowmask,
                    float* logits, BiDims d, cudaStream_t stream) {
  if (!bi_fast_path(d)) return -100;
  BiTcParams p{};
  p.hmat = hmat; p.hbias = hbias; p.rowmask = rowmask; p.logits = logits;
  p.B = d.B; p.K = d.K; p.Q = d.Q; p.G = d.G; p.C = d.C; p.NR = 16 * d.G; p.nchunks = d.C / CCH;
  return launch_bi<false>(vb, qb, p, stream, "bilinear_fwd_tc");
}

int bilinear_bwd_tc(const bf16* vb, const bf16* qb, const float* hmat, const float* dlogits, bf16* dzv, bf16* dzq,
                    float* dbv, float* dbq, float* dhmat, float* dhbias, BiDims d, cudaStream_t stream) {
  if (!bi_fast_path(d)) return -100;
  BiTcParams p{};
  p.hmat = hmat; p.dlogits = dlogits; p.dzv = dzv; p.dzq = dzq; p.dbq = dbq; p.dhmat = dhmat; p.dhbias = dhbias;
  p.B = d.B; p.K = d.K; p.Q = d.Q; p.G = d.G; p.C = d.C; p.NR = 16 * d.G; p.nchunks = d.C / CCH;
  if (int rc = launch_bi<true>(vb, qb, p, stream, "bilinear_bwd_tc")) return rc;
  // bias gradient of the v projection: column sums of dzv
  return act_bwd_bias(dzv, 1, nullptr, nullptr, dbv, (long)d.B * d.K, d.C, stream);
}

}  // namespace cti
