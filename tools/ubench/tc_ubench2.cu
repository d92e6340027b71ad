// Micro-benchmark 2: SM-wide tcgen05.mma throughput with 1..4 concurrently issuing threads (one per warp), small N,
// M = 128 vs 64, distinct operand tiles per issuer.  Measurement tool, not product code.
#include <cstdio>
#include <cstring>
#include "cti_common.cuh"
#include "tc_tiles.cuh"
namespace cti { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace cti;
struct Res { unsigned long long v[16]; };

__global__ void __launch_bounds__(256, 1) k_multi(Res* res, int M, int N, int n, int nthreads, int commits_each, int a_mn = 0, int b_mn = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + 4 * (16384 + 8192), slot = sBar + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(sBar + 8 * i, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  uint32_t tm; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tm) : "r"(slot));
  if (warp >= 1 && warp <= nthreads && lane == 0) {
    const int w = warp - 1;
    const uint32_t sA = base + w * (16384 + 8192), sB = sA + 16384;
    // a_mn == 2: MIXED -- issuer w uses shape / major-ness variant w (different instruction descriptors per thread)
    const int va = a_mn == 2 ? (w & 1) : a_mn, vb = a_mn == 2 ? ((w >> 1) & 1) : b_mn;
    const int vN = a_mn == 2 ? (w == 3 ? 64 : N) : N;
    const uint32_t idesc = make_idesc_rt(M, vN, va, vb);
    const uint64_t da = va ? desc_mnmajor(sA, 0, 2048) : desc_kmajor(sA, 0);
    const uint64_t db = vb ? desc_mnmajor(sB, 0, 2048) : desc_kmajor(sB, 0);
    for (int rep = 0; rep < 3; ++rep) {
      unsigned long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        umma_bf16_ss(tm + (N > 64 ? 0 : w * 128 + ((i * 16) & 63)), (a_mn ? da : da + (uint64_t)((i & 3) * 2)), db, idesc, 0u);
        if (commits_each) umma_commit(sBar + 8 * (8 + w));
      }
      unsigned long long t1 = clock64();
      umma_commit(sBar + 8 * w);
      mbar_wait(sBar + 8 * w, rep & 1);
      unsigned long long t2 = clock64();
      tcgen05_fence_after();
      res->v[w] = t1 - t0; res->v[4 + w] = t2 - t0;
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  Res* d; cudaMalloc(&d, sizeof(Res)); Res h;
  const size_t smem = 4 * (16384 + 8192) + 2048;
  cudaFuncSetAttribute(k_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("# M x N x 16, n MMAs per thread, T issuer threads: max total cycles / (n*T) = SM-wide cycles per MMA\n");
  for (int commits_each = 0; commits_each < 2; ++commits_each)
  for (int M : {128, 64}) for (int N : {16, 64}) for (int T = 1; T <= 4; ++T) {
    const int n = 32;
    cudaMemset(d, 0, sizeof(Res));
    k_multi<<<1, 256, smem>>>(d, M, N, n, T, commits_each);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int w = 0; w < T; ++w) mx = h.v[4 + w] > mx ? h.v[4 + w] : mx;
    printf("commit_each=%d M=%3d N=%2d threads=%d: per-thread issue %llu %llu %llu %llu total max %llu -> %.1f cycles/MMA SM-wide, %.1f per thread\n",
           commits_each, M, N, T, h.v[0], h.v[1], h.v[2], h.v[3], mx, (double)mx / (n * T), (double)mx / n);
    fflush(stdout);
  }
  printf("# operand major-ness (n = 32 per thread)\n");
  struct C { int M, N, a_mn, b_mn; const char* what; };
  const C cases[] = {{128, 16, 1, 0, "F1: A MN-major (T_r^T), B K-major"}, {128, 16, 0, 0, "F2: both K-major"},
                     {128, 192, 0, 1, "III: A K-major, B MN-major N=192"}, {64, 192, 0, 1, "III with M=64"},
                     {128, 192, 0, 0, "N=192 both K-major"}, {128, 16, 1, 1, "B2-like: both MN-major"}, {128, 16, 0, 1, "B3-like: A K, B MN"}};
  for (const C& c : cases) for (int T = 1; T <= 3; T += 2) {
    cudaMemset(d, 0, sizeof(Res));
    k_multi<<<1, 256, smem>>>(d, c.M, c.N, 32, T, 0, c.a_mn, c.b_mn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int w = 0; w < T; ++w) mx = h.v[4 + w] > mx ? h.v[4 + w] : mx;
    printf("%-40s M=%3d N=%3d threads=%d: %.1f cycles/MMA SM-wide\n", c.what, c.M, c.N, T, (double)mx / (32 * T));
    fflush(stdout);
  }
  printf("# mixed instruction descriptors across issuers (thread w: a_mn = w&1, b_mn = (w>>1)&1, N = 64 for w = 3) vs uniform\n");
  for (int mixed = 0; mixed < 2; ++mixed) for (int T = 2; T <= 4; ++T) {
    cudaMemset(d, 0, sizeof(Res));
    k_multi<<<1, 256, smem>>>(d, 128, 16, 32, T, 0, mixed ? 2 : 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int w = 0; w < T; ++w) mx = h.v[4 + w] > mx ? h.v[4 + w] : mx;
    printf("%s threads=%d: %.1f cycles/MMA SM-wide\n", mixed ? "mixed  " : "uniform", T, (double)mx / (32 * T));
  }
  return 0;
}
