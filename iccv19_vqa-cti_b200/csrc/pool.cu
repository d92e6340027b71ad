// Attention-weighted trilinear / bilinear pooling to the joint embedding (sm_100a, tcgen05):
//   TCNet.forward_with_weights  (reference src/tc.py:54-61)
//       out[b,c] = sum_{k,q,a} V[b,k,c] w[b,k,q,a] Qp[b,q,c] Ap[b,a,c]
//   BCNet.forward_with_weights  (reference src/bc.py:70-74)  -- the A == 0 case, Ap == 1
//
// Per sample and per 128-channel chunk (TMEM lane = channel c, column n = a*16 + q):
//   forward   Z[c,n]   = sum_k V[k,c] w[k,n]                      tcgen05.mma 128 x NC x 16, K = tokens
//             out[c]   = sum_a Ap[a,c] sum_q Qp[q,c] Z[c,(a,q)]   epilogue: a per-thread dot product
//   backward  dQp, dAp from Z (same epilogue shape);  KRd[c,n] = do[c] Qp[q,c] Ap[a,c] written to
//             shared memory ONCE and used by two more MMAs:
//             dV[k,c]  = sum_n w[k,n] KRd[c,n]                     128 x 128 x 16, K = columns n, TMEM lane = token k:
//                        every thread then owns 64 consecutive channels of one dzv row (16-byte stores; the first
//                        version had the channel on the lane and wrote 50 two-byte values per thread and chunk)
//             dw[k,n] += sum_c V[k,c] KRd[c,n]                     128 x NC x 16, K = channels (all chunks)
// The (B,C,K,Q,A) broadcast product of the reference's einsum never exists; V is read from HBM
// once per pass (TMA, 128B swizzle).  Every operand tile lives in shared memory in the TMA
// swizzle layout, so the SAME buffer is consumed K-major by one MMA and MN-major by another
// (tc_tiles.cuh): V as A(MN) for Z and A(K) for dw; w as B(MN) for Z and B(K) for dV; KRd as A(K)
// for dV and B(MN) for dw.
//
// The ReLU masks of the producing projections are applied here (V, Qp, Ap are post-ReLU), so the
// backward outputs are pre-activation gradients dz* plus their bias gradients.
//
// Roles (384 threads): warp 0 TMA producer (V tiles) | warp 1 MMA issuer | warp 2 TMEM allocator |
// warps 4-7 epilogue (TMEM lane quarter = warp % 4) | warps 8-11 build the bf16 w tile of the next sample.
#include "cti_common.cuh"
#include "cti_kernels.h"
#include "tc_tiles.cuh"

namespace cti {

namespace {

using bf16 = __nv_bfloat16;

constexpr int kThreads = 384;                 // forward
constexpr int kThreadsBwd = 768;              // backward: warps 12, 13 / 20, 21 take the two channel halves of the dV epilogue (12, 13 also dw), warps 16-19 share the Z epilogue (odd chunks)
constexpr int kEpiWarp0 = 4;
constexpr int kBuildWarp0 = 8;
constexpr int CCH = 128;                       // channels per chunk = TMEM lanes
constexpr int KP = 64;                         // token rows per tile
constexpr int V_STAGES = 4;
constexpr int V_TILE_BYTES = 2 * KP * 128;     // V: two 64-channel halves (128B swizzle)
constexpr int ST_Q = V_TILE_BYTES;             // Qp chunk [16 q][128 c] bf16, rows >= Q zero (TMA OOB fill)
constexpr int ST_A = ST_Q + 16 * CCH * 2;      // Ap chunk [8 a][128 c] bf16
constexpr int ST_DO = ST_A + 8 * CCH * 2;      // dout chunk [128] fp32 (backward)
constexpr int V_STAGE_BYTES = 23 * 1024;       // one pipeline stage: V + Qp + Ap + dout (+ pad to keep 1024B alignment)
constexpr int W_BYTES = 2 * KP * 128;          // two 64-column chunks
constexpr int KRD_BYTES = 2 * CCH * 128;       // two 64-column chunks x 128 channel rows
constexpr int MAX_C = 1024;

// TMEM column map
constexpr uint32_t TM_Z = 0, TM_DV = 256, TM_DW = 384;     // Z: 2 x 128, dV: 128 (token lanes x channels), dw: 128

struct PoolParams {
  const bf16* q;
  const bf16* a;
  const float* w;
  long w_stride_b;
  float* out;            // forward
  const float* dout;     // backward
  bf16 *dzv, *dzq, *dza;
  float *dbv, *dbq, *dba, *dw;
  long dw_stride_b;          // elements between the dw tiles of consecutive rows (K * Q * An when contiguous)
  int B, K, Q, A, An, C, NC, nchunks;
  int VR;                // rows b share the v tile of row b / VR (dzv stays per row b)
};

struct Smem {
  uint32_t v, w, krd, db, bars;
};
// barrier indices
enum { B_VFULL = 0, B_VEMPTY = 4, B_WFULL = 8, B_WEMPTY = 10, B_ZFULL = 12, B_ZEMPTY = 14, B_KFULL = 16, B_KEMPTY = 18,
       B_DFULL = 20, B_DEMPTY = 22, B_DWFULL = 24, B_DWEMPTY = 25, B_COUNT = 26 };

__host__ __device__ inline size_t pool_smem_bytes(bool bwd) {
  size_t o = V_STAGES * V_STAGE_BYTES + 2 * W_BYTES;
  if (bwd) o += 2 * KRD_BYTES + 3 * MAX_C * 4;
  return o + B_COUNT * 8 + 16 + 1024;
}

template <bool BWD>
__global__ void __launch_bounds__(BWD ? kThreadsBwd : kThreads, 1)
pool_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q,
            const __grid_constant__ CUtensorMap tmap_a, const PoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sV = base;
  const uint32_t sW = sV + V_STAGES * V_STAGE_BYTES;
  const uint32_t sK = sW + 2 * W_BYTES;
  const uint32_t sDb = sK + (BWD ? 2 * KRD_BYTES : 0);
  const uint32_t sBar = sDb + (BWD ? 3 * MAX_C * 4 : 0);
  const uint32_t tmem_slot = sBar + B_COUNT * 8;
  auto bar = [&](int i) { return sBar + 8u * i; };
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));     // generic pointer to the aligned base
  float* db_acc = reinterpret_cast<float*>(gen + (sDb - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the w tiles (pad rows / columns stay zero for the whole kernel) and the bias accumulators
  for (uint32_t i = threadIdx.x; i < 2 * W_BYTES / 16; i += blockDim.x) st_shared_v4(sW + i * 16, 0, 0, 0, 0);
  if (BWD)
    for (int i = threadIdx.x; i < 3 * MAX_C; i += blockDim.x) db_acc[i] = 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_q);
    if (p.A > 0) tma_prefetch_desc(&tmap_a);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(bar(B_VFULL + s), 1);
      mbar_init(bar(B_VEMPTY + s), BWD ? 9 : 5);      // MMA commit + every epilogue warp that reads the stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(B_WFULL + s), 128);
      mbar_init(bar(B_WEMPTY + s), 1);
      mbar_init(bar(B_ZFULL + s), 1);
      mbar_init(bar(B_ZEMPTY + s), 4);
      mbar_init(bar(B_KFULL + s), 4);
      mbar_init(bar(B_KEMPTY + s), 1);
      mbar_init(bar(B_DFULL + s), 1);
      mbar_init(bar(B_DEMPTY + s), 4);
    }
    mbar_init(bar(B_DWFULL), 1);
    mbar_init(bar(B_DWEMPTY), 2);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();          // zero fill of sW visible to the async proxy
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_prologue_done();      // everything above touched only this CTA's shared memory / TMEM

  const int n_my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_my * p.nchunks;
  const int ktok = (p.K + 15) >> 4;         // K steps over tokens
  const int kcol = p.NC >> 4;               // K steps over columns

  if (warp == 0) {
    // ------------------------------ TMA producer: V tiles ------------------------------
    if (elect_one_sync()) {
      for (int g = 0; g < total; ++g) {
        const int b = blockIdx.x + (g / p.nchunks) * gridDim.x, ch = g % p.nchunks;
        const int st = g % V_STAGES;
        mbar_wait(bar(B_VEMPTY + st), ((g / V_STAGES) & 1) ^ 1);
        const uint32_t dst = sV + st * V_STAGE_BYTES;
        mbar_arrive_expect_tx(bar(B_VFULL + st), V_TILE_BYTES + 16 * CCH * 2 + (p.A > 0 ? 8 * CCH * 2 : 0) + (BWD ? CCH * 4 : 0));
        tma_load_3d(&tmap_v, bar(B_VFULL + st), dst, ch * CCH, 0, b / p.VR);
        tma_load_3d(&tmap_v, bar(B_VFULL + st), dst + KP * 128, ch * CCH + 64, 0, b / p.VR);
        tma_load_3d(&tmap_q, bar(B_VFULL + st), dst + ST_Q, ch * CCH, 0, b);
        if (p.A > 0) tma_load_3d(&tmap_a, bar(B_VFULL + st), dst + ST_A, ch * CCH, 0, b);
        if (BWD) bulk_load_1d(dst + ST_DO, p.dout + (size_t)b * p.C + ch * CCH, CCH * 4, bar(B_VFULL + st));
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------------------
    if (elect_one_sync()) {
      const uint32_t id_z = make_idesc_rt(128, p.NC, 1, 1);
      const uint32_t id_dv = make_idesc_rt(128, CCH, 0, 0);
      const uint32_t id_dw = make_idesc_rt(128, p.NC, 0, 1);
      auto issue_z = [&](int g) {
        const int sl = g / p.nchunks, ch = g % p.nchunks;
        const uint32_t wbuf = sW + (sl & 1) * W_BYTES;
        if (ch == 0) mbar_wait(bar(B_WFULL + (sl & 1)), (sl >> 1) & 1);
        mbar_wait(bar(B_VFULL + g % V_STAGES), (g / V_STAGES) & 1);
        mbar_wait(bar(B_ZEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t vst = sV + (g % V_STAGES) * V_STAGE_BYTES;
        for (int ks = 0; ks < ktok; ++ks)
          umma_bf16_ss(tmem_base + TM_Z + (g & 1) * 128, desc_mnmajor(vst, ks, KP * 128), desc_mnmajor(wbuf, ks, KP * 128),
                       id_z, ks > 0 ? 1u : 0u);
        umma_commit(bar(B_ZFULL + (g & 1)));
        if (!BWD) {
          umma_commit(bar(B_VEMPTY + g % V_STAGES));
          if (ch == p.nchunks - 1) umma_commit(bar(B_WEMPTY + (sl & 1)));
        }
      };
      if (total > 0) issue_z(0);
      for (int g = 0; g < total; ++g) {
        if (g + 1 < total) issue_z(g + 1);
        if (!BWD) continue;
        const int sl = g / p.nchunks, ch = g % p.nchunks;
        const uint32_t wbuf = sW + (sl & 1) * W_BYTES;
        const uint32_t krd = sK + (g & 1) * KRD_BYTES;
        const uint32_t vst = sV + (g % V_STAGES) * V_STAGE_BYTES;
        mbar_wait(bar(B_KFULL + (g & 1)), (g >> 1) & 1);
        mbar_wait(bar(B_DEMPTY), (g & 1) ^ 1);
        tcgen05_fence_after();
        // A = w tile (rows = tokens; 64 rows per chunk, the upper 64 accumulator lanes read whatever follows and are
        // never read back), B = KRd (rows = channels)
        for (int ks = 0; ks < kcol; ++ks)
          umma_bf16_ss(tmem_base + TM_DV, desc_kmajor(wbuf + (ks >> 2) * (KP * 128), ks & 3),
                       desc_kmajor(krd + (ks >> 2) * (CCH * 128), ks & 3), id_dv, ks > 0 ? 1u : 0u);
        umma_commit(bar(B_DFULL));
        if (ch == 0) {
          mbar_wait(bar(B_DWEMPTY), (sl & 1) ^ 1);
          tcgen05_fence_after();
        }
        for (int ks = 0; ks < CCH / 16; ++ks)
          umma_bf16_ss(tmem_base + TM_DW, desc_kmajor(vst + (ks >> 2) * (KP * 128), ks & 3),
                       desc_mnmajor(krd, ks, CCH * 128), id_dw, (ch > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar(B_KEMPTY + (g & 1)));
        umma_commit(bar(B_VEMPTY + g % V_STAGES));
        if (ch == p.nchunks - 1) {
          umma_commit(bar(B_DWFULL));
          umma_commit(bar(B_WEMPTY + (sl & 1)));
        }
      }
    }
  } else if (warp >= kBuildWarp0 && warp < kBuildWarp0 + 4) {
    // ------------------------------ w tile builder ----------------------------------------
    // w[k][q][a] fp32 (contiguous per sample) -> bf16 tile element (row k, column a*16 + q).  All global
    // loads of a batch are issued before the first shared-memory store so their latencies overlap.
    const int t = threadIdx.x - kBuildWarp0 * 32;
    const int per_k = p.Q * p.An;
    const int n_el = p.K * per_k;
    // floor(e / d) = umulhi(e, ceil(2^32 / d)) for e < 2^16, d >= 2 (d == 1 handled apart: the magic overflows)
    const uint32_t m_perk = 0xFFFFFFFFu / per_k + 1, m_an = 0xFFFFFFFFu / p.An + 1;
    auto put = [&](uint32_t wbuf, int e, float val) {
      const int k = per_k == 1 ? e : (int)__umulhi((uint32_t)e, m_perk), rem = e - k * per_k;
      const int qi = p.An == 1 ? rem : (int)__umulhi((uint32_t)rem, m_an), ai = rem - qi * p.An;
      const int col = ai * 16 + qi;
      const __nv_bfloat16 hv = __float2bfloat16(val);
      st_shared_u16(wbuf + (col >> 6) * (KP * 128) + sw128_off(k, col & 63), *reinterpret_cast<const uint16_t*>(&hv));
    };
    for (int sl = 0; sl < n_my; ++sl) {
      const int b = blockIdx.x + sl * gridDim.x;
      const float* src = p.w + (size_t)b * p.w_stride_b;
      const uint32_t wbuf = sW + (sl & 1) * W_BYTES;
      const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n_el & 3) == 0);
      constexpr int U = 8;
      if (vec) {
        const int n4 = n_el >> 2;
        for (int i0 = t; i0 < n4; i0 += 128 * U) {
          float4 r[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * 128;
            r[u] = (i < n4) ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (i0 == t) mbar_wait(bar(B_WEMPTY + (sl & 1)), ((sl >> 1) & 1) ^ 1);   // buffer free (loads already in flight)
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * 128;
            if (i < n4) {
              put(wbuf, 4 * i, r[u].x);
              put(wbuf, 4 * i + 1, r[u].y);
              put(wbuf, 4 * i + 2, r[u].z);
              put(wbuf, 4 * i + 3, r[u].w);
            }
          }
        }
        if (t >= n4) mbar_wait(bar(B_WEMPTY + (sl & 1)), ((sl >> 1) & 1) ^ 1);
      } else {
        mbar_wait(bar(B_WEMPTY + (sl & 1)), ((sl >> 1) & 1) ^ 1);
        for (int e0 = t; e0 < n_el; e0 += 128 * U) {
          float r[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int e = e0 + u * 128;
            r[u] = (e < n_el) ? __ldg(src + e) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int e = e0 + u * 128;
            if (e < n_el) put(wbuf, e, r[u]);
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar(B_WFULL + (sl & 1)));
    }
  } else if (warp >= kEpiWarp0) {      // warps 4-7: Z epilogue (forward output / dQp, dAp, KRd); warps 12-15 (backward): dV, dw
    // ------------------------------ epilogue ------------------------------------------------
    const int qd = warp & 3;
    const int cl = qd * 32 + lane;                       // channel within the chunk = TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;

    struct Side { float qp[16]; float ap[8]; float dout; };
    auto load_side = [&](int g, Side& s) {      // this thread's channel of the Qp / Ap / dout chunk that came with the V tile
      const uint32_t st = sV + (g % V_STAGES) * V_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 16; ++i) s.qp[i] = bf16_bits_to_float(ld_shared_u16(st + ST_Q + (i * CCH + cl) * 2));
#pragma unroll
      for (int i = 0; i < 8; ++i)
        s.ap[i] = (p.A == 0) ? 1.f : bf16_bits_to_float(ld_shared_u16(st + ST_A + (i * CCH + cl) * 2));
      s.dout = BWD ? ld_shared_f32(st + ST_DO + cl * 4) : 0.f;
    };

    // dV epilogue: TMEM lane = token k (warps on lanes 0-63), `half` = which 64 channels of the chunk.  ReLU mask from
    // the V tile row of the same token, 128 contiguous bytes of dzv per thread; group 0 also drains the sample's dw tile.
    // (The image-side bias gradient is the column sum of dzv: a separate streaming pass, see tri_pool_bwd.)
    auto epi_b = [&](int h, int half) {
      const int sl = h / p.nchunks, ch = h % p.nchunks;
      const int b = blockIdx.x + sl * gridDim.x;
      const int k = cl;                                   // token = TMEM lane
      mbar_wait(bar(B_DFULL), h & 1);
      tcgen05_fence_after();
      const uint32_t vrow = sV + (h % V_STAGES) * V_STAGE_BYTES + half * (KP * 128) + (k >> 3) * 1024u + (k & 7) * 128u;
      bf16* dst = p.dzv + ((size_t)b * p.K + k) * p.C + ch * CCH + half * 64;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + TM_DV + half * 64 + t * 32, r);
        tmem_wait_ld();
        if (t == 1) {                                     // every TMEM read of this tile is done
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(B_DEMPTY));
        }
        if (k < p.K) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t vv[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3])
                         : "r"(vrow + ((((uint32_t)(t * 4 + j)) ^ (k & 7)) << 4)));
            uint32_t o[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const float2 f = unpack_bf16x2(vv[x]);
              o[x] = pack_bf16x2(f.x > 0.f ? __uint_as_float(r[j * 8 + 2 * x]) : 0.f,
                                 f.y > 0.f ? __uint_as_float(r[j * 8 + 2 * x + 1]) : 0.f);
            }
            reinterpret_cast<uint4*>(dst + t * 32)[j] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_VEMPTY + h % V_STAGES));
      if (half == 0 && ch == p.nchunks - 1) {       // the sample's dw tile is complete: lane = token
        mbar_wait(bar(B_DWFULL), sl & 1);
        tcgen05_fence_after();
        for (int ai = 0; ai < p.An; ++ai) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_DW + ai * 16, r);
          tmem_wait_ld();
          if (k < p.K) {
#pragma unroll
            for (int qi = 0; qi < 16; ++qi)
              if (qi < p.Q) p.dw[(size_t)b * p.dw_stride_b + ((size_t)k * p.Q + qi) * p.An + ai] = __uint_as_float(r[qi]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_DWEMPTY));
      }
    };

    if (BWD && ((warp >= 12 && warp < 16) || warp >= 20)) {
      const int half = warp >= 20 ? 1 : 0;
      if (qd < 2) {                          // token lanes 0-63; the other warps of the two groups have nothing to do
        for (int h = 0; h < total; ++h) {
          mbar_wait(bar(B_VFULL + h % V_STAGES), (h / V_STAGES) & 1);      // TMA data of the stage visible to this thread
          epi_b(h, half);
        }
      }
    } else {
    Side cur;
    // backward: two warp groups alternate over the chunks (Z / KRd are double buffered on the chunk parity)
    for (int g = (BWD && warp >= 16) ? 1 : 0; g < total; g += BWD ? 2 : 1) {
      const int sl = g / p.nchunks, ch = g % p.nchunks;
      const int b = blockIdx.x + sl * gridDim.x, c = ch * CCH + cl;
      mbar_wait(bar(B_VFULL + g % V_STAGES), (g / V_STAGES) & 1);      // TMA data of this stage visible to this thread
      load_side(g, cur);
      __syncwarp();                                                  // this group does not read the stage again
      if (lane == 0) mbar_arrive(bar(B_VEMPTY + g % V_STAGES));
      mbar_wait(bar(B_ZFULL + (g & 1)), (g >> 1) & 1);
      tcgen05_fence_after();
      if (BWD) mbar_wait(bar(B_KEMPTY + (g & 1)), ((g >> 1) & 1) ^ 1);
      const uint32_t krd = sK + (g & 1) * KRD_BYTES;
      // packed fp32x2 math (FFMA2 / FMUL2): pairs (q, q + 1) of this thread's channel
      float2 qp2[8], dq2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        qp2[j] = make_float2(cur.qp[2 * j], cur.qp[2 * j + 1]);
        dq2[j] = make_float2(0.f, 0.f);
      }
      float acc = 0.f, sum_a = 0.f;
      const uint32_t krow = (cl >> 3) * 1024u + (cl & 7u) * 128u, kx = cl & 7u;      // this thread's KRd row (SW128)
      bf16* dza_c = BWD && p.A > 0 ? p.dza + (size_t)b * p.A * p.C + c : nullptr;
#pragma unroll
      for (int ai = 0; ai < 8; ++ai) {
        if (ai < p.An) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_base + lane_addr + TM_Z + (g & 1) * 128 + ai * 16, r);
          tmem_wait_ld();
          const float apv = cur.ap[ai];
          float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            s2 = __ffma2_rn(qp2[j], make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), s2);
          const float s = s2.x + s2.y;
          if (!BWD) {
            acc = fmaf(apv, s, acc);
          } else {
            const float2 ap2 = make_float2(apv, apv);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dq2[j] = __ffma2_rn(ap2, make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), dq2[j]);
            if (p.A > 0) {
              const float ga = apv > 0.f ? cur.dout * s : 0.f;
              dza_c[(size_t)ai * p.C] = __float2bfloat16(ga);
              sum_a += ga;
            }
            // KRd[c, (ai, q)] = do * Qp[q] * Ap[ai]: one 32-byte piece of this thread's operand row
            const float da = cur.dout * apv;
            const float2 da2 = make_float2(da, da);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 pr = __fmul2_rn(da2, qp2[j]);
              pk[j] = pack_bf16x2(pr.x, pr.y);
            }
            const uint32_t tile = krd + (ai >> 2) * (CCH * 128) + krow;
            const uint32_t ch0 = (ai & 3) * 2;
            st_shared_v4(tile + (((ch0) ^ kx) << 4), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(tile + (((ch0 + 1) ^ kx) << 4), pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      if (!BWD) {
        p.out[(size_t)b * p.C + c] = acc;
      } else {
        float sum_q = 0.f;
        bf16* dzq_c = p.dzq + (size_t)b * p.Q * p.C + c;
        const float2 do2 = make_float2(cur.dout, cur.dout);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 gq2 = __fmul2_rn(do2, dq2[j]);
          if (2 * j < p.Q) {
            const float gq = qp2[j].x > 0.f ? gq2.x : 0.f;
            dzq_c[(size_t)(2 * j) * p.C] = __float2bfloat16(gq);
            sum_q += gq;
          }
          if (2 * j + 1 < p.Q) {
            const float gq = qp2[j].y > 0.f ? gq2.y : 0.f;
            dzq_c[(size_t)(2 * j + 1) * p.C] = __float2bfloat16(gq);
            sum_q += gq;
          }
        }
        atomicAdd(db_acc + MAX_C + c, sum_q);        // the two Z-epilogue groups may own the same channel in turn
        atomicAdd(db_acc + 2 * MAX_C + c, sum_a);
        fence_proxy_async_smem();        // KRd rows written with st.shared are read by tcgen05.mma
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (BWD) mbar_arrive(bar(B_KFULL + (g & 1)));
        mbar_arrive(bar(B_ZEMPTY + (g & 1)));
      }
    }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (BWD) {
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      atomicAdd(p.dbq + c, db_acc[MAX_C + c]);
      if (p.A > 0) atomicAdd(p.dba + c, db_acc[2 * MAX_C + c]);
    }
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int check_pool(const PoolDims& d, const char* who) {
  CTI_REQUIRE(d.B >= 0 && d.K > 0 && d.Q > 0 && d.A >= 0 && d.C > 0, "%s: bad dims", who);
  CTI_REQUIRE(d.VR >= 1 && d.B % d.VR == 0, "%s: B=%d rows do not divide into groups of v_rep=%d", who, d.B, d.VR);
  CTI_REQUIRE(d.C % CCH == 0 && d.C <= MAX_C, "%s: channel count %d must be a multiple of %d and <= %d", who, d.C, CCH,
              MAX_C);
  CTI_REQUIRE(d.K <= KP, "%s: at most %d regions per sample (K=%d)", who, KP, d.K);
  CTI_REQUIRE(d.Q <= 16 && d.A <= 8, "%s: Q <= 16 and A <= 8 are supported (Q=%d A=%d)", who, d.Q, d.A);
  return 0;
}

template <bool BWD>
int launch_pool(const bf16* v, PoolParams p, cudaStream_t stream, const char* who) {
  CUtensorMap tv, tq, ta;
  if (int rc = make_tmap_3d(&tv, v, p.C, p.K, p.B / p.VR, p.C, (uint64_t)p.K * p.C, 64, KP)) return rc;
  if (int rc = make_tmap_3d(&tq, p.q, p.C, p.Q, p.B, p.C, (uint64_t)p.Q * p.C, CCH, 16, false)) return rc;
  if (p.A > 0) {
    if (int rc = make_tmap_3d(&ta, p.a, p.C, p.A, p.B, p.C, (uint64_t)p.A * p.C, CCH, 8, false)) return rc;
  } else {
    ta = tq;
  }
  if (BWD) CTI_REQUIRE((reinterpret_cast<uintptr_t>(p.dout) & 15) == 0, "%s: dout must be 16-byte aligned", who);
  const size_t smem = pool_smem_bytes(BWD);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pool_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s smem attr: %s", who, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int grid = p.B < kNumSMsB200 ? p.B : kNumSMsB200;
  launch_pdl(pool_kernel<BWD>, dim3(grid), dim3(BWD ? kThreadsBwd : kThreads), smem, stream, tv, tq, ta, p);
  return check_launch(who);
}

PoolParams make_params(const bf16* q, const bf16* a, const float* w, long w_stride_b, PoolDims d) {
  PoolParams p{};
  p.q = q; p.a = a; p.w = w; p.w_stride_b = w_stride_b;
  p.B = d.B; p.K = d.K; p.Q = d.Q; p.A = d.A; p.An = d.A > 0 ? d.A : 1; p.C = d.C;
  p.NC = 16 * p.An;
  p.nchunks = d.C / CCH;
  p.VR = d.VR;
  return p;
}

}  // namespace

int tri_pool_fwd(const bf16* v, const bf16* q, const bf16* a, const float* w, long w_stride_b, float* out, PoolDims d,
                 cudaStream_t stream) {
  if (int rc = check_pool(d, "tri_pool_fwd")) return rc;
  if (d.B == 0) return 0;
  PoolParams p = make_params(q, a, w, w_stride_b, d);
  p.out = out;
  return launch_pool<false>(v, p, stream, "tri_pool_fwd");
}

int tri_pool_bwd(const bf16* v, const bf16* q, const bf16* a, const float* w, long w_stride_b, const float* dout,
                 bf16* dzv, bf16* dzq, bf16* dza, float* dbv, float* dbq, float* dba, float* dw, long dw_stride_b,
                 PoolDims d, cudaStream_t stream) {
  if (int rc = check_pool(d, "tri_pool_bwd")) return rc;
  if (d.B == 0) return 0;
  PoolParams p = make_params(q, a, w, w_stride_b, d);
  p.dout = dout; p.dzv = dzv; p.dzq = dzq; p.dza = dza; p.dbv = dbv; p.dbq = dbq; p.dba = dba; p.dw = dw;
  p.dw_stride_b = dw_stride_b > 0 ? dw_stride_b : (long)d.K * d.Q * (d.A > 0 ? d.A : 1);
  if (int rc = launch_pool<true>(v, p, stream, "tri_pool_bwd")) return rc;
  // image-side bias gradient: column sums of dzv (token-on-lane epilogue: no cheap in-kernel reduction over tokens)
  return act_bwd_bias(dzv, 1, nullptr, nullptr, dbv, (long)d.B * d.K, d.C, stream);
}

}  // namespace cti
