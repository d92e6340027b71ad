// Building blocks of the per-sample tcgen05 kernels (pooling, contraction): operand tiles that
// threads write into shared memory in the SAME 128-byte-swizzled layout TMA produces, so one
// buffer can be consumed by tcgen05.mma as a K-major or as an MN-major operand; descriptor
// helpers; TMEM loads; 3-D TMA loads.
//
// SW128 tile = rows of 128 bytes (64 bf16), 8-row atoms of 1024 bytes; inside an atom the 16-byte
// chunk c of row r is stored at chunk position c ^ (r % 8).  Read as
//   K-major  operand: tile row = M/N index, the 64 elements of a row run along K
//                     (descriptor: start + 32 B per 16-element K step, SBO = 1024);
//   MN-major operand: tile row = K index, the 64 elements of a row run along M/N
//                     (descriptor: start + 2048 B per 16-row K step, SBO = 1024, LBO = distance to
//                     the tile holding the next 64 M/N elements).
#pragma once

#include "cti_common.cuh"

namespace cti {

// Byte offset of element (row, col) inside a [rows][64] bf16 SW128 tile.
__device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t col) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((col >> 3) & 7u) ^ (row & 7u)) << 4) + ((col & 7u) << 1);
}

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, uint32_t kstep) {      // kstep: 16 elements
  return make_smem_desc_sw128(tile_addr + kstep * 32u, 16u, 1024u);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, uint32_t kstep, uint32_t lbo_bytes) {
  return make_smem_desc_sw128(tile_addr + kstep * 2048u, lbo_bytes, 1024u);
}

__device__ __forceinline__ uint32_t make_idesc_rt(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (quadrant base + t).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 1-D bulk copy global -> shared (bytes % 16 == 0, both 16-byte aligned), completes on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D bulk copy shared -> global (bulk-group completion; bytes % 16 == 0, both 16-byte aligned).
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t ld_shared_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }

// Pipeline stage / phase bookkeeping for an N-deep mbarrier ring.
struct Ring {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t depth) {
    if (++stage == depth) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// host side (defined in gemm_tcgen05.cu)
int make_tmap_3d(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                 uint64_t stride2_elems, uint32_t box0, uint32_t box1, bool swizzle128 = true, uint32_t box2 = 1);

int make_tmap_4d(CUtensorMap* map, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&strides_elems)[3],
                 const uint32_t (&box)[4]);

}  // namespace cti
