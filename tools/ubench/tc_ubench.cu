// Micro-benchmarks of the sm_100a primitives the per-sample kernels are built from (one CTA, cycle counts):
// tcgen05.mma issue / completion for small N, tcgen05.commit, mbarrier waits and hand-offs, tcgen05.ld,
// fence.proxy.async, shared-memory store patterns.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I../../iccv19_vqa-cti_b200/csrc tc_ubench.cu -o tc_ubench ; run on a B200.  Measurement tool, not product code.
#include <cstdio>
#include <cstring>
#include <vector>
#include "cti_common.cuh"
#include "tc_tiles.cuh"

namespace cti { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace cti;

__device__ __forceinline__ unsigned long long clk() { return clock64(); }

struct Res { unsigned long long v[64]; };

// mode 0: n MMAs (M=128, N, K=16) back to back by one thread, then commit, then wait.
__global__ void __launch_bounds__(192, 1) k_mma(Res* res, int N, int n, int same_d, int a_mn, int b_mn, int dual = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, sBar = base + 16384 + 32768, slot = sBar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(sBar + 8 * i, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(slot));
  if (dual && warp == 2 && lane == 0) {          // second issuer: same work into the upper TMEM columns, own barrier
    const uint32_t idesc = make_idesc_rt(128, N, a_mn, b_mn);
    const uint64_t da = desc_kmajor(sA, 1), db = desc_kmajor(sB, 1);
    for (int rep = 0; rep < 3; ++rep) {
      unsigned long long t0 = clk();
      for (int i = 0; i < n; ++i) umma_bf16_ss(tm + 256 + (same_d ? 0 : ((i * N) & 255)), da, db, idesc, 0u);
      unsigned long long t1 = clk();
      umma_commit(sBar + 8);
      mbar_wait(sBar + 8, rep & 1);
      unsigned long long t3 = clk();
      tcgen05_fence_after();
      res->v[8] = t1 - t0; res->v[11] = t3 - t0;
    }
  }
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_rt(128, N, a_mn, b_mn);
    const uint64_t da = a_mn ? desc_mnmajor(sA, 0, 2048) : desc_kmajor(sA, 0);
    const uint64_t db = b_mn ? desc_mnmajor(sB, 0, 2048) : desc_kmajor(sB, 0);
    for (int rep = 0; rep < 3; ++rep) {           // rep 2 is the one reported (warm)
      unsigned long long t0 = clk();
      for (int i = 0; i < n; ++i) umma_bf16_ss(tm + (same_d ? 0 : ((i * N) & 255)), da, db, idesc, 0u);
      unsigned long long t1 = clk();
      umma_commit(sBar);
      unsigned long long t2 = clk();
      mbar_wait(sBar, rep & 1);
      unsigned long long t3 = clk();
      tcgen05_fence_after();
      res->v[0] = t1 - t0; res->v[1] = t2 - t1; res->v[2] = t3 - t2; res->v[3] = t3 - t0;
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

// mode 1: misc primitive costs
__global__ void __launch_bounds__(256, 1) k_misc(Res* res, int which) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, sBar = base + 16384 + 32768, slot = sBar + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(sBar + 8 * i, i < 8 ? 1 : 32); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(slot));
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  if (which == 0 && warp == 1 && lane == 0) {
    // (a) 8 commits with nothing pending
    unsigned long long t0 = clk();
    for (int i = 0; i < 8; ++i) umma_commit(sBar + 8 * i);
    unsigned long long t1 = clk();
    res->v[0] = (t1 - t0) / 8;
    for (int i = 0; i < 8; ++i) mbar_wait(sBar + 8 * i, 0);
    // (b) wait on an already-completed phase, 16 times
    t0 = clk();
    for (int i = 0; i < 16; ++i) mbar_wait(sBar + 8 * (i & 7), 0);
    t1 = clk();
    res->v[1] = (t1 - t0) / 16;
    // (c) one MMA N=16 then commit then wait: single small MMA round trip
    const uint32_t idesc = make_idesc_rt(128, 16, 0, 0);
    t0 = clk();
    umma_bf16_ss(tm, desc_kmajor(sA, 0), desc_kmajor(sB, 0), idesc, 0u);
    umma_commit(sBar);
    mbar_wait(sBar, 1);
    t1 = clk();
    res->v[2] = t1 - t0;
    // (d) clock read overhead
    t0 = clk(); t1 = clk();
    res->v[3] = t1 - t0;
    // (e) mbarrier arrive (local) x16
    t0 = clk();
    for (int i = 0; i < 16; ++i) mbar_arrive(sBar + 8 * 1);
    t1 = clk();
    res->v[4] = (t1 - t0) / 16;
  }
  __syncthreads();
  // (f) tcgen05.ld latencies, warp 4 alone, then warps 4-7 together
  if (which == 1 && warp == 4) {
    uint32_t r8[8], r16[16], r32[32];
    unsigned long long t0 = clk();
    tmem_ld_32x32b_x8(tm + lane_addr, r8); tmem_wait_ld();
    unsigned long long t1 = clk();
    tmem_ld_32x32b_x16(tm + lane_addr, r16); tmem_wait_ld();
    unsigned long long t2 = clk();
    tmem_ld_32x32b_x32(tm + lane_addr, r32); tmem_wait_ld();
    unsigned long long t3 = clk();
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) acc += r8[i];
    for (int i = 0; i < 16; ++i) acc += r16[i];
    for (int i = 0; i < 32; ++i) acc += r32[i];
    if (lane == 0) { res->v[8] = t1 - t0; res->v[9] = t2 - t1; res->v[10] = t3 - t2; res->v[63] = acc; }
    // 16 x (x32) back to back with one wait: throughput
    t0 = clk();
#pragma unroll
    for (int i = 0; i < 8; ++i) { tmem_ld_32x32b_x32(tm + lane_addr + i * 32, r32); for (int j = 0; j < 32; ++j) acc += r32[j]; }
    tmem_wait_ld();
    t1 = clk();
    if (lane == 0) { res->v[11] = (t1 - t0) / 8; res->v[62] = acc; }
  }
  __syncthreads();
  if (which == 1 && warp >= 4) {
    uint32_t r32[32]; uint32_t acc = 0;
    unsigned long long t0 = clk();
#pragma unroll
    for (int i = 0; i < 8; ++i) { tmem_ld_32x32b_x32(tm + lane_addr + i * 32, r32); for (int j = 0; j < 32; ++j) acc += r32[j]; }
    tmem_wait_ld();
    unsigned long long t1 = clk();
    if (lane == 0 && warp == 4) { res->v[12] = (t1 - t0) / 8; res->v[61] = acc; }
  }
  __syncthreads();
  // (g) st.shared + fence.proxy.async cost (one warp; then 4 warps)
  if (which == 2 && warp == 4) {
    unsigned long long t0 = clk();
    for (int i = 0; i < 8; ++i) st_shared_v4(sA + (lane * 16 + i * 512), 1, 2, 3, 4);
    unsigned long long t1 = clk();
    fence_proxy_async_smem();
    unsigned long long t2 = clk();
    for (int i = 0; i < 24; ++i) st_shared_u16(sA + ((lane & 15) * 2 + (lane >> 4) * 1024 + i * 128), (uint16_t)i);
    unsigned long long t3 = clk();
    fence_proxy_async_smem();
    unsigned long long t4 = clk();
    if (lane == 0) { res->v[16] = t1 - t0; res->v[17] = t2 - t1; res->v[18] = t3 - t2; res->v[19] = t4 - t3; }
  }
  __syncthreads();
  // (h) ping-pong hand-off between warp 5 (lane 0) and warp 6 (lane 0): arrive -> wake latency
  if (which != 3) {
  } else if (warp == 5 && lane == 0) {
    unsigned long long t0 = clk();
    for (int i = 0; i < 16; ++i) { mbar_arrive(sBar + 8 * 2); mbar_wait(sBar + 8 * 3, i & 1); }
    unsigned long long t1 = clk();
    res->v[20] = (t1 - t0) / 32;      // one-way hand-off
  } else if (warp == 6 && lane == 0) {
    for (int i = 0; i < 16; ++i) { mbar_wait(sBar + 8 * 2, i & 1); mbar_arrive(sBar + 8 * 3); }
  }
  __syncthreads();
  // (i) whole-warp hand-off as the converters do it: st.shared, fence.proxy.async, tcgen05 fence, syncwarp, lane-0 arrive
  if (which != 4) {
  } else if (warp == 5) {
    unsigned long long t0 = clk();
    for (int i = 0; i < 16; ++i) {
      st_shared_v4(sA + lane * 16, i, 2, 3, 4);
      fence_proxy_async_smem(); tcgen05_fence_before(); __syncwarp();
      if (lane == 0) mbar_arrive(sBar + 8 * 4);
      mbar_wait(sBar + 8 * 5, i & 1);
      tcgen05_fence_after();
    }
    unsigned long long t1 = clk();
    if (lane == 0) res->v[21] = (t1 - t0) / 32;
  } else if (warp == 6) {
    for (int i = 0; i < 16; ++i) {
      mbar_wait(sBar + 8 * 4, i & 1);
      tcgen05_fence_after();
      st_shared_v4(sB + lane * 16, i, 2, 3, 4);
      fence_proxy_async_smem(); tcgen05_fence_before(); __syncwarp();
      if (lane == 0) mbar_arrive(sBar + 8 * 5);
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  Res* d; cudaMalloc(&d, sizeof(Res)); Res h;
  const size_t smem = 16384 + 32768 + 1024 + 1024;
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_misc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("# MMA M=128 x N x 16 (bf16), n back to back from one thread: issue cycles, commit cycles, wait cycles, total\n");
  const int Ns[] = {16, 64, 96, 192, 256};
  const int ns[] = {1, 4, 16, 32};
  for (int same_d = 0; same_d < 2; ++same_d)
    for (int N : Ns) for (int n : ns) {
      k_mma<<<1, 192, smem>>>(d, N, n, same_d, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
      printf("same_d=%d N=%3d n=%2d issue=%5llu commit=%4llu wait=%5llu total=%5llu per_mma=%.1f\n", same_d, N, n, h.v[0], h.v[1], h.v[2], h.v[3], (double)h.v[3] / n);
      fflush(stdout);
    }
  printf("# two issuer threads (warps 1 and 2) concurrently, each n MMAs: issue / total cycles of each\n");
  for (int N : {16, 64, 192}) for (int n : {4, 16, 32}) {
    k_mma<<<1, 192, smem>>>(d, N, n, 0, 0, 0, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("dual N=%3d n=%2d  t1: issue=%5llu total=%5llu   t2: issue=%5llu total=%5llu  per_mma(2n)=%.1f\n", N, n, h.v[0], h.v[3], h.v[8], h.v[11],
           (double)(h.v[3] > h.v[11] ? h.v[3] : h.v[11]) / (2 * n));
    fflush(stdout);
  }
  printf("# MN-major A (as F1) N=16\n");
  for (int n : ns) {
    k_mma<<<1, 192, smem>>>(d, 16, n, 0, 1, 0); cudaDeviceSynchronize(); cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    printf("a_mn N= 16 n=%2d issue=%5llu commit=%4llu wait=%5llu total=%5llu per_mma=%.1f\n", n, h.v[0], h.v[1], h.v[2], h.v[3], (double)h.v[3] / n);
    fflush(stdout);
  }
  printf("# misc\n"); fflush(stdout);
  Res acc_h; memset(&acc_h, 0, sizeof(acc_h));
  for (int which = 0; which < 5; ++which) {
    cudaMemset(d, 0, sizeof(Res));
    k_misc<<<1, 256, smem>>>(d, which);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("misc %d error %s\n", which, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 64; ++i) acc_h.v[i] += h.v[i];
    printf("misc step %d done\n", which); fflush(stdout);
  }
  h = acc_h;
  printf("commit (nothing pending) %llu cyc; wait on completed phase %llu; 1 MMA(N=16)+commit+wait round trip %llu; clock read %llu; mbar_arrive %llu\n",
         h.v[0], h.v[1], h.v[2], h.v[3], h.v[4]);
  printf("tcgen05.ld+wait one warp: x8 %llu, x16 %llu, x32 %llu; x32 pipelined per ld: 1 warp %llu, 4 warps %llu\n", h.v[8], h.v[9], h.v[10], h.v[11], h.v[12]);
  printf("8 x st.shared.v4 %llu, fence.proxy.async %llu; 24 x st.shared.u16 (2-way conflict) %llu, fence %llu\n", h.v[16], h.v[17], h.v[18], h.v[19]);
  printf("thread->thread mbarrier hand-off (one way) %llu; warp hand-off incl. st.shared + fences %llu\n", h.v[20], h.v[21]);
  return 0;
}
