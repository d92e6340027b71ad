// The R per-rank projections of a TCNet modality in TRAINING mode, with the reference's semantics: every per-rank FCNet
// owns an nn.Dropout on its input (reference src/tc.py:29-31, src/fc.py:25-26), so rank r sees the tucker output y masked
// with ITS OWN Bernoulli mask:
//
//     z[row, r*16 + j] = relu( s * sum_k keep_r[row, k] * y[row, k] * W_r[j, k] + b_r[j] ),       s = 1 / (1 - p)
//
// As a GEMM this has the FLOPs of the eval-mode grouped projection (M x 512 x 512) but 32 different A operands.  The first
// version materialised the masked copies in HBM (cti_dropout_expand, 4 ranks at a time, block-diagonal weights: 4x redundant
// FLOPs and ~2 GB of traffic per call; 5.8 ms of a 9 ms training step).  Here the mask is applied to the A FRAGMENT IN
// REGISTERS: y sits in shared memory once per row tile, every (16-row slab, rank, 16-column step) regenerates its 8 keep
// decisions per lane from Philox and ANDs them onto the bf16x2 fragment registers before mma.sync.  The work is bound by
// the integer pipe (mask generation), not by the tensor pipe -- warp-level mma.sync (operands in registers) is the right
// instruction here; tcgen05.mma reads A from shared memory / TMEM and would need the 32 masked copies written there.
//
//   fwd  : CTA = 128 rows, 16 warps = 8 slabs x 2 rank sub-groups; W streamed through smem 16 ranks x 64 k at a time.
//   dgrad: dy[row, k] = s * sum_r keep_r[row, k] * (dz_r[row, :] W_r)[k], ReLU mask of y applied, bf16 out.  The mask sits
//          on the OUTPUT fragment of each per-rank MMA, so ranks cannot share an accumulator pass: 64-row tiles, W streamed
//          one rank pair at a time.
//   wgrad: dW_r[j, k] += s * sum_row dz_r[row, j] * keep_r[row, k] * y[row, k]: CTA = (rank pair, row split); the masked y
//          fragment is transposed in registers (movmatrix) into the B operand; fp32 red.add into the zeroed accumulator.
//
// Mask definition (shared by the three kernels and by rank_proj_mask, which the tests use).  A lane's 8 decisions for one
// (16-row slab, 16-column step ks, rank) are ordered like the mma A fragment: e <-> (column half e/4, row half (e/2)%2,
// column parity e%2).
//   general p: one Philox4x32-7 call per (slab, ks, rank pair, lane) yields 16 bytes = 8 decisions for each rank of the pair;
//     keep <=> byte >= round(256 p): the drop probability is quantised to 1/256 (0.2 -> 0.19922) and
//     s = 256 / (256 - round(256 p)) keeps the expectation exact.
//   p == 0.5 (the image side: 3/4 of the rows): one random BIT per decision -- one call per (slab, 4 steps, rank quad, lane)
//     yields 16 bytes = 8 decisions for each (step, rank); 8x fewer Philox calls and no threshold compare.
// Up to 4 problems (modalities) share a launch: the question and answer sides fill one wave together.
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {
namespace {

using bf16 = __nv_bfloat16;

constexpr int H = 512;                 // width of the tucker output (h_mm of the reference's TriAttention)
constexpr int kThreads = 512;
constexpr int kMaxProb = 4;

struct RankRng {
  uint2 key;
  uint32_t site;
  uint32_t t7x4;       // low 7 bits of the threshold byte, replicated 4x
  int thr_hi;          // threshold >= 128
  float scale;
};

struct Prob {
  const bf16* y;       // (M, H)
  const bf16* W;       // (R*16, H) effective weights
  const float* bias;   // fwd
  bf16* out;           // fwd: (M, R*16)
  const bf16* dz;      // bwd: (M, R*16)
  bf16* dzt;           // dgrad: (M, H)
  float* dW;           // wgrad: (R*16, H) accumulator
  long M;
  RankRng rng;
};

struct Probs {
  Prob p[kMaxProb];
  int first[kMaxProb + 1];     // first block (fwd / dgrad: x, wgrad: y) of each problem
  int per_split[kMaxProb];     // wgrad: row tiles per split
  int n;
};

__device__ __forceinline__ int find_prob(const Probs& ps, int block) {
  int i = 0;
#pragma unroll
  for (int k = 1; k < kMaxProb; ++k)
    if (k < ps.n && block >= ps.first[k]) i = k;
  return i;
}

__device__ __forceinline__ uint4 philox4x32_7(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// general p: 16 random bytes of (slab, ks, rank pair, lane): .x/.y rank 2*pair, .z/.w rank 2*pair + 1
__device__ __forceinline__ uint4 rank_bytes(const RankRng& r, uint32_t slab, int ks, int pair, int lane) {
  return philox4x32_7(make_uint4(slab, (uint32_t)ks | ((uint32_t)pair << 16), (uint32_t)lane, r.site), r.key);
}
// p == 0.5: 128 random bits of (slab, 4 steps kc, rank quad, lane): word k4 (.x .y .z .w) = step kc*4 + k4, byte i = rank 4*quad + i
__device__ __forceinline__ uint4 rank_bits(const RankRng& r, uint32_t slab, int kc, int quad, int lane) {
  return philox4x32_7(make_uint4(slab, (uint32_t)kc | ((uint32_t)quad << 16) | 0x80000000u, (uint32_t)lane, r.site), r.key);
}
__device__ __forceinline__ uint32_t word_of(const uint4& w, int k4) { return k4 == 0 ? w.x : (k4 == 1 ? w.y : (k4 == 2 ? w.z : w.w)); }

// per byte: MSB set <=> byte >= threshold (the other bits are garbage)
__device__ __forceinline__ uint32_t ge_msb(const RankRng& r, uint32_t x) {
  const uint32_t d = (x | 0x80808080u) - r.t7x4;        // no borrow crosses a byte: MSB <=> low 7 bits >= low 7 bits of thr
  return r.thr_hi ? (x & d) : (x | d);
}
// 4 keep bits (a nibble) -> the MSBs of the 4 bytes (products of distinct powers of two: no carries)
__device__ __forceinline__ uint32_t nib_msb(uint32_t nib) { return nib * 0x10204080u; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}

// 4 AND masks (one per A-fragment register: 0xFFFF per kept bf16) from the two MSB words of one rank
__device__ __forceinline__ void frag_masks(uint32_t glo, uint32_t ghi, uint32_t (&m)[4]) {
  m[0] = prmt(glo, 0x9988u);      // bytes 0, 1 -> sign-replicated halves
  m[1] = prmt(glo, 0xBBAAu);      // bytes 2, 3
  m[2] = prmt(ghi, 0x9988u);
  m[3] = prmt(ghi, 0xBBAAu);
}

// The two MSB words (decisions 0-3, 4-7) of rank `i` (0..3 within its quad) at step k4 of the 4-step group.
// HALF: from the quad's bit words; general: from the pair's byte words (rb of pair i / 2).
template <bool HALF>
__device__ __forceinline__ void keep_words(const RankRng& rng, const uint4& src, int k4, int i, uint32_t& glo, uint32_t& ghi) {
  if constexpr (HALF) {
    const uint32_t byte = (word_of(src, k4) >> (8 * i)) & 0xFFu;
    glo = nib_msb(byte & 0xFu);
    ghi = nib_msb(byte >> 4);
  } else {
    glo = ge_msb(rng, (i & 1) ? src.z : src.x);
    ghi = ge_msb(rng, (i & 1) ? src.w : src.y);
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                           // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory tiles of 16-byte chunks, chunk index XOR-swizzled with the row so that the 8 rows of an ldmatrix hit
// 8 different bank groups: byte offset of (row, chunk) in a tile with `pitch` bytes per row
__device__ __forceinline__ uint32_t sw(int row, int chunk, int pitch) { return row * pitch + ((chunk ^ (row & 7)) << 4); }

// ------------------------------------------------------------------------------------------------------------------- //
// forward
// ------------------------------------------------------------------------------------------------------------------- //
constexpr int F_TILE = 128;                          // rows per CTA
constexpr int F_WST = 256 * 128;                     // W stage: 16 ranks x 16 outputs rows of 64 k (128 bytes)
constexpr size_t F_SMEM = (size_t)F_TILE * H * 2 + 2 * F_WST;

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
rank_proj_fwd_kernel(const __grid_constant__ Probs ps, int R) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sY = smem_u32(smem), sW = sY + F_TILE * H * 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp & 7, sub = warp >> 3;        // 16 rows of the tile; ranks sub*8 .. sub*8+7 of the 16 staged ones
  const int pi = find_prob(ps, blockIdx.x);
  const Prob& pb = ps.p[pi];
  const RankRng& rng = pb.rng;
  const bf16* __restrict__ y = pb.y;
  const bf16* __restrict__ W = pb.W;
  const long M = pb.M;
  const long row0 = (long)(blockIdx.x - ps.first[pi]) * F_TILE;
  const int N = R * 16;
  const int halves = R / 16, n_stage = halves * (H / 64);
  pdl_prologue_done();

  auto load_w = [&](int st) {                        // stage st = (half, 64-wide k chunk)
    const int h = st / (H / 64), kc = st % (H / 64);
    const uint32_t dst = sW + (st & 1) * F_WST;
    for (int i = tid; i < 256 * 8; i += kThreads) {
      const int row = i >> 3, c = i & 7;
      cp16(dst + sw(row, c, 128), W + ((size_t)h * 256 + row) * H + kc * 64 + c * 8, true);
    }
  };
  for (int i = tid; i < F_TILE * (H / 8); i += kThreads) {
    const int row = i / (H / 8), c = i % (H / 8);
    const bool ok = row0 + row < M;
    cp16(sY + sw(row, c, H * 2), y + (ok ? (size_t)(row0 + row) * H + c * 8 : 0), ok);
  }
  load_w(0);
  cp_commit();

  const uint32_t gslab = (uint32_t)(row0 / 16) + slab;
  const int a_row = slab * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_csel = lane >> 4;      // ldmatrix.x4 address roles (A)
  const int b_row = (lane & 7) + (lane >> 4) * 8, b_csel = (lane >> 3) & 1;                  // (B: matrices n0k0, n0k1, n1k0, n1k1)
  float acc[8][2][4];
  for (int st = 0; st < n_stage; ++st) {
    const int h = st / (H / 64), kc = st % (H / 64);
    if (kc == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;
    }
    if (st + 1 < n_stage) {
      load_w(st + 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const uint32_t wst = sW + (st & 1) * F_WST;
    const int quad0 = h * 4 + sub * 2;               // this warp's two rank quads
    uint4 qb[2];
    if constexpr (HALF) {
      qb[0] = rank_bits(rng, gslab, kc, quad0, lane);
      qb[1] = rank_bits(rng, gslab, kc, quad0 + 1, lane);
    }
    auto step = [&](int k4) {
      const int ks = kc * 4 + k4;
      uint32_t a[4];
      ldsm_x4(sY + sw(a_row, ks * 2 + a_csel, H * 2), a);
#pragma unroll
      for (int qd = 0; qd < 2; ++qd) {
        uint4 rb[2];
        if constexpr (!HALF) {
          rb[0] = rank_bytes(rng, gslab, ks, (quad0 + qd) * 2, lane);
          rb[1] = rank_bytes(rng, gslab, ks, (quad0 + qd) * 2 + 1, lane);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t glo, ghi, m[4], am[4], b[4];
          keep_words<HALF>(rng, HALF ? qb[qd] : rb[i >> 1], k4, i, glo, ghi);
          frag_masks(glo, ghi, m);
#pragma unroll
          for (int e = 0; e < 4; ++e) am[e] = a[e] & m[e];
          const int rr = sub * 8 + qd * 4 + i;
          ldsm_x4(wst + sw(rr * 16 + b_row, k4 * 2 + b_csel, 128), b);
          mma_bf16(acc[qd * 4 + i][0], am, b[0], b[1]);
          mma_bf16(acc[qd * 4 + i][1], am, b[2], b[3]);
        }
      }
    };
    if constexpr (HALF) {                            // k4 must be a compile-time constant to pick the bit word
      step(0); step(1); step(2); step(3);
    } else {
#pragma unroll 1
      for (int k4 = 0; k4 < 4; ++k4) step(k4);
    }
    __syncthreads();                                 // the buffer of stage st is refilled by the prefetch of iteration st + 1
    if (kc == H / 64 - 1) {                          // the half is complete: bias, ReLU, bf16
      const int g = lane >> 2, t = lane & 3;
      const long r_lo = row0 + slab * 16 + g, r_hi = r_lo + 8;
      const float* __restrict__ bias = pb.bias;
      bf16* __restrict__ out = pb.out;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rank = h * 16 + sub * 8 + i;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int col = rank * 16 + n * 8 + 2 * t;
          const float2 bb = *reinterpret_cast<const float2*>(bias + col);
          if (r_lo < M)
            *reinterpret_cast<uint32_t*>(out + (size_t)r_lo * N + col) =
                pack_bf16x2(fmaxf(fmaf(acc[i][n][0], rng.scale, bb.x), 0.f), fmaxf(fmaf(acc[i][n][1], rng.scale, bb.y), 0.f));
          if (r_hi < M)
            *reinterpret_cast<uint32_t*>(out + (size_t)r_hi * N + col) =
                pack_bf16x2(fmaxf(fmaf(acc[i][n][2], rng.scale, bb.x), 0.f), fmaxf(fmaf(acc[i][n][3], rng.scale, bb.y), 0.f));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------- //
// dgrad: dzt[row, k] = (y[row, k] > 0) * s * sum_r keep_r[row, k] * sum_j dz[row, r*16 + j] W[r*16 + j, k]     (bf16)
// ------------------------------------------------------------------------------------------------------------------- //
constexpr int D_TILE = 64;
constexpr int D_WST = 64 * H * 2;                    // W stage: one rank quad, 64 rows of H
// smem: dz tile [64][R*16] + 2 W stages

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
rank_proj_dgrad_kernel(const __grid_constant__ Probs ps, int R) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int N = R * 16;
  const uint32_t sZ = smem_u32(smem), sW = sZ + D_TILE * N * 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp & 3, cq = warp >> 2;         // 16 rows of the tile; column quarter (128 columns = 8 steps of 16)
  const int pi = find_prob(ps, blockIdx.x);
  const Prob& pb = ps.p[pi];
  const RankRng& rng = pb.rng;
  const bf16* __restrict__ dz = pb.dz;
  const bf16* __restrict__ W = pb.W;
  const long M = pb.M;
  const long row0 = (long)(blockIdx.x - ps.first[pi]) * D_TILE;
  pdl_prologue_done();

  auto load_w = [&](int quad) {
    const uint32_t dst = sW + (quad & 1) * D_WST;
    for (int i = tid; i < 64 * (H / 8); i += kThreads) {
      const int row = i / (H / 8), c = i % (H / 8);
      cp16(dst + sw(row, c, H * 2), W + ((size_t)quad * 64 + row) * H + c * 8, true);
    }
  };
  for (int i = tid; i < D_TILE * (N / 8); i += kThreads) {
    const int row = i / (N / 8), c = i % (N / 8);
    const bool ok = row0 + row < M;
    cp16(sZ + sw(row, c, N * 2), dz + (ok ? (size_t)(row0 + row) * N + c * 8 : 0), ok);
  }
  load_w(0);
  cp_commit();

  const uint32_t gslab = (uint32_t)(row0 / 16) + slab;
  const int a_row = slab * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, a_csel = lane >> 4;
  const int bt_row = (lane & 7) + ((lane >> 3) & 1) * 8, bt_csel = lane >> 4;     // .trans: matrices (j lo, n0), (j hi, n0), (j lo, n1), (j hi, n1)
  float acc[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
  const int quads = R / 4;
  for (int qd = 0; qd < quads; ++qd) {
    if (qd + 1 < quads) {
      load_w(qd + 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const uint32_t wst = sW + (qd & 1) * D_WST;
    uint4 qb[2];
    if constexpr (HALF) {
      qb[0] = rank_bits(rng, gslab, cq * 2, qd, lane);
      qb[1] = rank_bits(rng, gslab, cq * 2 + 1, qd, lane);
    }
#pragma unroll 1
    for (int pr = 0; pr < 2; ++pr) {                 // the quad's two rank pairs
      uint32_t a[2][4];
      ldsm_x4(sZ + sw(a_row, (qd * 4 + pr * 2) * 2 + a_csel, N * 2), a[0]);
      ldsm_x4(sZ + sw(a_row, (qd * 4 + pr * 2 + 1) * 2 + a_csel, N * 2), a[1]);
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const int ks = cq * 8 + k8;
        uint4 rb;
        if constexpr (!HALF) rb = rank_bytes(rng, gslab, ks, qd * 2 + pr, lane);
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          uint32_t b[4];
          ldsm_x4_trans(wst + sw((pr * 2 + s2) * 16 + bt_row, ks * 2 + bt_csel, H * 2), b);
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16(c0, a[s2], b[0], b[1]);           // columns ks*16 .. +7
          mma_bf16(c1, a[s2], b[2], b[3]);           // columns ks*16 + 8 .. +15
          // keep decisions of the OUTPUT fragment: c[0], c[1] = (row g, cols 2t, 2t+1) <-> 0, 1; c[2], c[3] = row g+8 <-> 2, 3
          if constexpr (HALF) {
            // pr is a run-time value here: shift the byte of rank pr*2 + s2 out of the step's word
            const uint32_t byte = word_of(qb[k8 >> 2], k8 & 3) >> (16 * pr + 8 * s2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[2 * k8][e] += (byte >> e) & 1u ? c0[e] : 0.f;
              acc[2 * k8 + 1][e] += (byte >> (4 + e)) & 1u ? c1[e] : 0.f;
            }
          } else {
            const uint32_t glo = ge_msb(rng, s2 ? rb.z : rb.x), ghi = ge_msb(rng, s2 ? rb.w : rb.y);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[2 * k8][e] += (glo >> (8 * e + 7)) & 1u ? c0[e] : 0.f;
              acc[2 * k8 + 1][e] += (ghi >> (8 * e + 7)) & 1u ? c1[e] : 0.f;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  const int g = lane >> 2, t = lane & 3;
  const long r_lo = row0 + slab * 16 + g, r_hi = r_lo + 8;
  const bf16* __restrict__ y = pb.y;
  bf16* __restrict__ dzt = pb.dzt;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    const int col = cq * 128 + n * 8 + 2 * t;
    if (r_lo < M) {
      const float2 yy = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(y + (size_t)r_lo * H + col));
      *reinterpret_cast<uint32_t*>(dzt + (size_t)r_lo * H + col) =
          pack_bf16x2(yy.x > 0.f ? acc[n][0] * rng.scale : 0.f, yy.y > 0.f ? acc[n][1] * rng.scale : 0.f);
    }
    if (r_hi < M) {
      const float2 yy = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(y + (size_t)r_hi * H + col));
      *reinterpret_cast<uint32_t*>(dzt + (size_t)r_hi * H + col) =
          pack_bf16x2(yy.x > 0.f ? acc[n][2] * rng.scale : 0.f, yy.y > 0.f ? acc[n][3] * rng.scale : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------- //
// wgrad: dW[r*16 + j, k] += s * sum_row dz[row, r*16 + j] * keep_r[row, k] * y[row, k]        (fp32, red.add)
// CTA = (rank pair, row split); warp = (64-column block, half of the tile's slabs), both ranks of the pair.
// ------------------------------------------------------------------------------------------------------------------- //
constexpr int W_TILE = 64;
constexpr int W_YST = W_TILE * H * 2;                // y stage
constexpr int W_ZST = W_TILE * 64;                   // dz stage: 64 rows x 32 columns (the pair's 2 x 16 outputs) = 64 bytes per row
constexpr size_t W_SMEM = 2 * (size_t)(W_YST + W_ZST);

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
rank_proj_wgrad_kernel(const __grid_constant__ Probs ps, int R) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sY = smem_u32(smem), sZ = sY + 2 * W_YST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cb = warp & 7, sh = warp >> 3;           // 64-column block (4 steps); slabs sh*2, sh*2+1 of each 64-row tile
  const int pair = blockIdx.x, N = R * 16;
  const int pi = find_prob(ps, blockIdx.y);
  const Prob& pb = ps.p[pi];
  const RankRng& rng = pb.rng;
  const bf16* __restrict__ y = pb.y;
  const bf16* __restrict__ dz = pb.dz;
  const long M = pb.M;
  const long n_tiles = (M + W_TILE - 1) / W_TILE;
  const long t_lo = (long)(blockIdx.y - ps.first[pi]) * ps.per_split[pi], t_hi = min(t_lo + ps.per_split[pi], n_tiles);
  pdl_prologue_done();
  if (t_lo >= t_hi) return;

  auto load_tile = [&](long tile) {
    const int buf = (int)((tile - t_lo) & 1);
    const long row0 = tile * W_TILE;
    for (int i = tid; i < W_TILE * (H / 8); i += kThreads) {
      const int row = i / (H / 8), c = i % (H / 8);
      const bool ok = row0 + row < M;
      cp16(sY + buf * W_YST + sw(row, c, H * 2), y + (ok ? (size_t)(row0 + row) * H + c * 8 : 0), ok);
    }
    for (int i = tid; i < W_TILE * 4; i += kThreads) {
      const int row = i >> 2, c = i & 3;
      const bool ok = row0 + row < M;
      cp16(sZ + buf * W_ZST + row * 64 + ((c ^ ((row >> 1) & 3)) << 4), dz + (ok ? (size_t)(row0 + row) * N + pair * 32 + c * 8 : 0), ok);
    }
  };
  load_tile(t_lo);
  cp_commit();

  float acc[2][8][4];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[s2][n][e] = 0.f;
  const int y_row = (lane & 7) + ((lane >> 3) & 1) * 8, y_csel = lane >> 4;           // A-layout fragment of y (as the forward)
  const int z_row = (lane & 7) + (lane >> 4) * 8, z_csel = (lane >> 3) & 1;           // .trans: (rows lo, j lo), (rows lo, j hi), (rows hi, j lo), (rows hi, j hi)
  for (long tile = t_lo; tile < t_hi; ++tile) {
    const int buf = (int)((tile - t_lo) & 1);
    if (tile + 1 < t_hi) {
      load_tile(tile + 1);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
#pragma unroll 1
    for (int sl = 0; sl < 2; ++sl) {
      const int slab = sh * 2 + sl;
      const uint32_t gslab = (uint32_t)(tile * (W_TILE / 16)) + slab;
      uint32_t za[2][4];                             // dz^T fragments of the two ranks: A[m = j][k = row]
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int row = slab * 16 + z_row, c = s2 * 2 + z_csel;
        ldsm_x4_trans(sZ + buf * W_ZST + row * 64 + ((c ^ ((row >> 1) & 3)) << 4), za[s2]);
      }
      uint4 qb;
      if constexpr (HALF) qb = rank_bits(rng, gslab, cb, pair >> 1, lane);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int ks = cb * 4 + k4;
        uint32_t a[4];
        ldsm_x4(sY + buf * W_YST + sw(slab * 16 + y_row, ks * 2 + y_csel, H * 2), a);
        uint4 rb;
        if constexpr (!HALF) rb = rank_bytes(rng, gslab, ks, pair, lane);
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          uint32_t glo, ghi, m[4], b[4];
          if constexpr (HALF) {                      // byte of rank (pair & 1) * 2 + s2 of the quad
            const uint32_t byte = (word_of(qb, k4) >> (16 * (pair & 1) + 8 * s2)) & 0xFFu;
            glo = nib_msb(byte & 0xFu);
            ghi = nib_msb(byte >> 4);
          } else {
            keep_words<false>(rng, rb, k4, s2, glo, ghi);
          }
          frag_masks(glo, ghi, m);
#pragma unroll
          for (int e = 0; e < 4; ++e) b[e] = movm_trans(a[e] & m[e]);     // (rows x cols) blocks -> B[k = row][n = col]
          mma_bf16(acc[s2][2 * k4], za[s2], b[0], b[1]);                   // columns ks*16 .. +7:  rows 0-7 (a0), rows 8-15 (a1)
          mma_bf16(acc[s2][2 * k4 + 1], za[s2], b[2], b[3]);               // columns ks*16 + 8 .. +15
        }
      }
    }
    __syncthreads();
  }
  const int g = lane >> 2, t = lane & 3;
  float* __restrict__ dW = pb.dW;
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = cb * 64 + n * 8 + 2 * t;
      float* p_lo = dW + ((size_t)(pair * 2 + s2) * 16 + g) * H + col;
      float* p_hi = p_lo + 8 * H;
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p_lo), "f"(acc[s2][n][0] * rng.scale), "f"(acc[s2][n][1] * rng.scale) : "memory");
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p_hi), "f"(acc[s2][n][2] * rng.scale), "f"(acc[s2][n][3] * rng.scale) : "memory");
    }
}

// keep[r, row, k] in {0, 1} (uint8): the mask the three kernels above apply, element by element (tests)
__global__ void __launch_bounds__(256) rank_proj_mask_kernel(uint8_t* __restrict__ keep, long M, int R, const RankRng rng, int half) {
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= (long)R * M * H) return;
  const int k = (int)(i % H);
  const long row = (i / H) % M;
  const int r = (int)(i / ((long)H * M));
  const int r16 = (int)(row & 15), c16 = k & 15;
  const int lane = (r16 & 7) * 4 + ((c16 & 7) >> 1);
  const int e = (c16 >> 3) * 4 + (r16 >> 3) * 2 + (c16 & 1);
  const int ks = k >> 4;
  if (half) {
    const uint4 qb = rank_bits(rng, (uint32_t)(row >> 4), ks >> 2, r >> 2, lane);
    keep[i] = (word_of(qb, ks & 3) >> (8 * (r & 3) + e)) & 1u;
  } else {
    const uint4 rb = rank_bytes(rng, (uint32_t)(row >> 4), ks, r >> 1, lane);
    const uint32_t w = (r & 1) ? (e < 4 ? rb.z : rb.w) : (e < 4 ? rb.x : rb.y);
    keep[i] = (ge_msb(rng, w) >> (8 * (e & 3) + 7)) & 1u;
  }
}

int make_rank_rng(float p, uint64_t seed, uint64_t site, RankRng& r, bool& half, const char* who) {
  CTI_REQUIRE(p > 0.f && p < 1.f, "%s: dropout p=%f must be in (0, 1)", who, p);
  int thr = (int)lrintf(p * 256.f);
  thr = thr < 1 ? 1 : (thr > 255 ? 255 : thr);
  half = p == 0.5f;
  r.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  r.site = (uint32_t)site;
  r.t7x4 = (uint32_t)(thr & 0x7F) * 0x01010101u;
  r.thr_hi = thr >= 128;
  r.scale = 256.f / (256.f - (float)thr);
  return 0;
}

int check_shape(int Hin, int R, int n, const char* who) {
  CTI_REQUIRE(Hin == H && (R == 16 || R == 32),
              "%s: built for a %d-wide input and 16 or 32 ranks (got width %d, %d ranks)", who, H, Hin, R);
  CTI_REQUIRE(n >= 1 && n <= kMaxProb, "%s: 1 to %d problems per call (got %d)", who, kMaxProb, n);
  return 0;
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes, const char* who) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s smem attr: %s", who, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

enum class Pass { Fwd, Dgrad, Wgrad };

// The problems of one call, split by mask kind (p == 0.5 or not): one launch per kind.
int run(Pass pass, const RankProjProblem* probs, int n, int Hin, int R, cudaStream_t s, const char* who) {
  if (int rc = check_shape(Hin, R, n, who)) return rc;
  for (int kind = 0; kind < 2; ++kind) {
    Probs ps{};
    int blocks = 0;
    for (int i = 0; i < n; ++i) {
      const RankProjProblem& q = probs[i];
      CTI_REQUIRE(q.M >= 0, "%s: negative row count", who);
      if (q.M == 0) continue;
      RankRng rng;
      bool half;
      if (int rc = make_rank_rng(q.p, q.seed, q.site, rng, half, who)) return rc;
      if ((int)half != kind) continue;
      Prob& d = ps.p[ps.n];
      d.y = q.y; d.W = q.w_eff; d.bias = q.bias; d.out = q.out; d.dz = q.dz; d.dzt = q.dzt; d.dW = q.dw_accum;
      d.M = q.M;
      d.rng = rng;
      CTI_REQUIRE(q.y != nullptr && ((uintptr_t)q.y & 15) == 0, "%s: y must be a 16-byte aligned pointer", who);
      const long n64 = (q.M + 63) / 64;
      ps.first[ps.n] = blocks;
      if (pass == Pass::Fwd) {
        CTI_REQUIRE(q.w_eff && q.bias && q.out && ((uintptr_t)q.w_eff & 15) == 0 && ((uintptr_t)q.out & 3) == 0 && ((uintptr_t)q.bias & 7) == 0,
                    "%s: forward needs aligned w_eff, bias, out", who);
        blocks += (int)((q.M + F_TILE - 1) / F_TILE);
      } else if (pass == Pass::Dgrad) {
        CTI_REQUIRE(q.w_eff && q.dz && q.dzt && ((uintptr_t)q.w_eff & 15) == 0 && ((uintptr_t)q.dz & 15) == 0 && ((uintptr_t)q.dzt & 3) == 0,
                    "%s: dgrad needs aligned w_eff, dz, dzt", who);
        blocks += (int)n64;
      } else {
        CTI_REQUIRE(q.dz && q.dw_accum && ((uintptr_t)q.dz & 15) == 0 && ((uintptr_t)q.dw_accum & 7) == 0,
                    "%s: wgrad needs aligned dz, dw_accum", who);
        const int pairs = R / 2;
        long splits = (2 * kNumSMsB200 + pairs - 1) / pairs;     // ~2 waves of CTAs when one problem has the launch to itself
        if (splits > n64) splits = n64;
        const int per = (int)((n64 + splits - 1) / splits);
        ps.per_split[ps.n] = per;
        blocks += (int)((n64 + per - 1) / per);
      }
      ++ps.n;
    }
    if (ps.n == 0) continue;
    ps.first[ps.n] = blocks;
    const size_t d_smem = (size_t)D_TILE * R * 16 * 2 + 2 * D_WST;
    if (pass == Pass::Fwd) {
      auto kern = kind ? rank_proj_fwd_kernel<true> : rank_proj_fwd_kernel<false>;
      if (int rc = set_smem(kern, F_SMEM, who)) return rc;
      launch_pdl(kern, dim3(blocks), dim3(kThreads), F_SMEM, s, ps, R);
    } else if (pass == Pass::Dgrad) {
      auto kern = kind ? rank_proj_dgrad_kernel<true> : rank_proj_dgrad_kernel<false>;
      if (int rc = set_smem(kern, d_smem, who)) return rc;
      launch_pdl(kern, dim3(blocks), dim3(kThreads), d_smem, s, ps, R);
    } else {
      auto kern = kind ? rank_proj_wgrad_kernel<true> : rank_proj_wgrad_kernel<false>;
      if (int rc = set_smem(kern, W_SMEM, who)) return rc;
      launch_pdl(kern, dim3(R / 2, blocks), dim3(kThreads), W_SMEM, s, ps, R);
    }
    if (int rc = check_launch(who)) return rc;
  }
  return 0;
}

}  // namespace

float rank_proj_scale(float p) {
  RankRng r;
  bool half;
  if (make_rank_rng(p, 0, 0, r, half, "rank_proj_scale")) return 0.f;
  return r.scale;
}

int rank_proj_dropout_fwd(const RankProjProblem* probs, int n, int Hin, int R, cudaStream_t s) {
  return run(Pass::Fwd, probs, n, Hin, R, s, "rank_proj_dropout_fwd");
}
int rank_proj_dropout_dgrad(const RankProjProblem* probs, int n, int Hin, int R, cudaStream_t s) {
  return run(Pass::Dgrad, probs, n, Hin, R, s, "rank_proj_dropout_dgrad");
}
int rank_proj_dropout_wgrad(const RankProjProblem* probs, int n, int Hin, int R, cudaStream_t s) {
  return run(Pass::Wgrad, probs, n, Hin, R, s, "rank_proj_dropout_wgrad");
}

int rank_proj_dropout_mask(uint8_t* keep, long M, int Hin, int R, float p, uint64_t seed, uint64_t site, cudaStream_t s) {
  if (int rc = check_shape(Hin, R, 1, "rank_proj_dropout_mask")) return rc;
  RankRng rng;
  bool half;
  if (int rc = make_rank_rng(p, seed, site, rng, half, "rank_proj_dropout_mask")) return rc;
  if (M <= 0) return 0;
  const long n = (long)R * M * H;
  rank_proj_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keep, M, R, rng, (int)half);
  return check_launch("rank_proj_mask_kernel");
}

}  // namespace cti
