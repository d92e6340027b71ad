// Common device helpers for the sm_100a CTI kernels: mbarrier, TMA, tcgen05/TMEM
// PTX wrappers, small vector/bf16 utilities and the host-side error plumbing.
// Everything here is hand-written inline PTX for Blackwell (sm_100a); there is
// no fallback path for other architectures.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef CTI_WATCHDOG
// 0 (default): plain mbarrier wait loops.  1: trap instead of hanging forever if a pipeline barrier never flips;
// 2: also printf which one.  Debug builds only (CTI_WATCHDOG=1 python build.py --force): the spin counter in every wait
// loop costs ~5 % of a training step, and the inlined printf call sites another ~5 % (i-cache and issue slots of the
// role warps that share a scheduler with the waiting ones).
#define CTI_WATCHDOG 0
#endif

namespace cti {

// --------------------------------------------------------------------------- //
// host-side error state (thread local); see cti_capi.cu
// --------------------------------------------------------------------------- //
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // returns 0 or the positive cudaError_t, recording the message

#define CTI_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::cti::set_error(__VA_ARGS__);           \
      return -1;                               \
    }                                          \
  } while (0)

constexpr int kNumSMsB200 = 148;

// --------------------------------------------------------------------------- //
// small device utilities
// --------------------------------------------------------------------------- //
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// --------------------------------------------------------------------------- //
// Programmatic dependent launch.  Every kernel of the library is launched with launch_pdl(): it may begin (block
// scheduling, and in the TMA / tcgen05 kernels the prologue: barrier init, TMEM allocation, descriptor prefetch)
// while the previous kernel of the stream is still draining.  pdl_wait() returns once that kernel has completed and
// its writes are visible -- it precedes the first global-memory access of every kernel; pdl_launch_dependents() lets
// the NEXT kernel start early in the same way.  Both are no-ops when the neighbour is not a PDL launch.
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue_done() {
  pdl_launch_dependents();
  pdl_wait();
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface in check_launch()
}

// One lane of a CONVERGED warp.  The single-thread roles (TMA producer, tcgen05.mma issuers) enter their loop through
// this instead of `lane == 0`: with a threadIdx-derived condition ptxas cannot prove that exactly one lane is active
// and wraps every UTCHMMA / UTMALDG / UTCBAR in an ELECT ... BRA.U.ANY waterfall loop -- 95 instead of 50 cycles per
// tcgen05.mma from one thread (tools/ubench/tc_ubench3.cu).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// --------------------------------------------------------------------------- //
// mbarrier
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or ~the hint elapses,
// instead of re-issuing the probe.  In the warp-specialised kernels several mostly-idle role warps share an SM
// sub-partition with the warps doing the conversions; un-hinted probes from the idle ones took a large share of
// the issue slots (ncu: ~30 M of 53 M executed warp instructions in the first trilinear kernel were wait loops).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if CTI_WATCHDOG
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) {   // many seconds: a broken pipeline, not a slow one
#if CTI_WATCHDOG > 1              // bring-up only: every inlined printf call site costs ~30 instructions of i-cache
      printf("cti: mbarrier watchdog fired (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
#endif
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// --------------------------------------------------------------------------- //
// TMA (cp.async.bulk.tensor), 2-D tiled loads that complete on an mbarrier
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// c0 = coordinate along the innermost (contiguous) global dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t smem_dst, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// TMA stores: shared -> global, bulk-group completion.  c0 = innermost coordinate, c1 = row; boxes that stick out of
// the tensor are clipped.  The reduce form adds the box into global memory (fp32 add performed at L2).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {      // <= N groups may still be READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// --------------------------------------------------------------------------- //
// tcgen05 / TMEM
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {   // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Same wait, but the loaded registers pass through the asm: no use of r[] can be scheduled above it.
__device__ __forceinline__ void tmem_wait_ld_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major : rows of 128 B (64 bf16 along K), 8-row groups SBO bytes apart; LBO unused.
//   MN-major: rows of 128 B (64 bf16 along M/N), 8 K-rows per swizzle atom, next 8 K-rows
//             SBO bytes apart, next 64-element M/N chunk LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;     // leading byte offset, bits [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;     // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version = 1 (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                              // layout type 2 = SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                                   // D format: F32
         | (1u << 7)                                 // A format: BF16
         | (1u << 10)                                // B format: BF16
         | ((a_mn_major ? 1u : 0u) << 15)            // A major-ness
         | ((b_mn_major ? 1u : 0u) << 16)            // B major-ness
         | (static_cast<uint32_t>(N >> 3) << 17)     // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);    // M / 16
}

}  // namespace cti
