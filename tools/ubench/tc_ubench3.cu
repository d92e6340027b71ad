// Micro-benchmark 3: tcgen05.mma issue rate of ONE thread when the issuing branch is entered through `lane == 0`
// (ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY waterfall loop: it cannot prove a single active lane) vs through
// elect.sync (ptxas emits the bare UTCHMMA).  Measurement tool, not product code.
#include <cstdio>
#include <cstring>
#include "cti_common.cuh"
#include "tc_tiles.cuh"
namespace cti { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace cti;
struct Res { unsigned long long v[16]; };

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <bool ELECT>
__global__ void __launch_bounds__(256, 1) k_issue(Res* res, int M, int N, int n, int nthreads) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + 4 * (16384 + 32768), slot = sBar + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(sBar + 8 * i, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(slot));
  if (warp >= 1 && warp <= nthreads) {
    const bool me = ELECT ? elect_one() : (lane == 0);
    if (me) {
      const int w = warp - 1;
      const uint32_t sA = base + w * (16384 + 32768), sB = sA + 16384;
      const uint32_t idesc = make_idesc_rt(M, N, 0, 0);
      const uint64_t da = desc_kmajor(sA, 0), db = desc_kmajor(sB, 0);
      for (int rep = 0; rep < 3; ++rep) {
        unsigned long long t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < n; ++i)
          umma_bf16_ss(tm + (N > 64 ? 0 : w * 128 + ((i * 16) & 63)), da + (uint64_t)((i & 3) * 2), db, idesc, 0u);
        unsigned long long t1 = clock64();
        umma_commit(sBar + 8 * w);
        mbar_wait(sBar + 8 * w, rep & 1);
        unsigned long long t2 = clock64();
        tcgen05_fence_after();
        res->v[w] = t1 - t0; res->v[4 + w] = t2 - t0;
      }
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  Res* d; cudaMalloc(&d, sizeof(Res)); Res h;
  const size_t smem = 4 * (16384 + 32768) + 2048;
  cudaFuncSetAttribute(k_issue<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_issue<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("# M=128 x N x 16, n = 32 MMAs per issuing thread: issue cycles per MMA (per thread) and total cycles per MMA SM-wide\n");
  for (int N : {16, 64, 128, 256}) for (int T = 1; T <= 4; T += (T == 1 ? 1 : 2)) for (int el = 0; el < 2; ++el) {
    const int n = 32;
    cudaMemset(d, 0, sizeof(Res));
    if (el) k_issue<true><<<1, 256, smem>>>(d, 128, N, n, T); else k_issue<false><<<1, 256, smem>>>(d, 128, N, n, T);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int w = 0; w < T; ++w) mx = h.v[4 + w] > mx ? h.v[4 + w] : mx;
    printf("N=%3d threads=%d %-8s issue/MMA %.1f  total/MMA SM-wide %.1f (tensor floor %.0f)\n", N, T, el ? "elect" : "lane==0",
           (double)h.v[0] / n, (double)mx / (n * T), 128.0 * N / 256);
  }
  return 0;
}
