// Small shared pieces of the per-sample tensor-core kernels (trilinear / pooling / bilinear):
// WMMA fragment aliases for bf16 16x16x16 tiles with fp32 accumulation, cp.async helpers.
#pragma once

#include "cti_common.cuh"

#include <mma.h>

namespace cti {
namespace tiles {

using namespace nvcuda;
using bf16 = __nv_bfloat16;
using FragAR = wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major>;
using FragAC = wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::col_major>;
using FragBR = wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major>;
using FragBC = wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::col_major>;
using FragC = wmma::fragment<wmma::accumulator, 16, 16, 16, float>;

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kScrLd = 20;                 // fp32 scratch tile pitch (floats)
constexpr int kScrFloats = 16 * kScrLd;
constexpr int kLdS = 24;                   // pitch of the 16-wide bf16 operand tiles
constexpr int kMaxAcc = 8;                 // accumulator tiles per warp

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace tiles
}  // namespace cti
