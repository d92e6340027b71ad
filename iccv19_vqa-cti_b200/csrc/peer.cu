// Gradient sum over NVLink peer memory without SM-side transfer kernels (dp.PeerAllReducer).
//
// NCCL's all-reduce runs as a persistent kernel: overlapped with backward it takes SMs away from the path's own
// persistent 148-CTA kernels (measured: 3.03 ms overlapped vs 2.97 ms serial at N = 2).  Here the bytes move on the COPY
// ENGINES: every rank owns one cudaMalloc'ed region (flag block | gradient slab | staging), exported to its peers through
// CUDA IPC.  Per bucket of gradients, on a side stream:
//   1. push   : my copy of chunk p  -> rank p's staging slot            (cudaMemcpyAsync, peer write over NVLink)
//   2. barrier: a one-CTA kernel; release-store of an epoch into every peer's flag block, acquire-spin on my own
//   3. reduce : chunk[me] += staged copies, in rank order               (one streaming kernel over 1/N of the bucket)
//   4. push   : the reduced chunk -> chunk[me] of every peer's slab
//   5. barrier
// Every rank ends with bit-identical sums (each element is added once, by its owner, in rank order).  The only SM work
// is step 3 (a few microseconds) and the two single-CTA barriers; everything is capturable in a CUDA graph.
#include "cti_common.cuh"
#include "cti_kernels.h"

#include <cuda.h>

namespace cti {

namespace {

constexpr int kMaxPeers = 16;
constexpr int kSlots = 8;
// flag block layout (uint32): [slot][peer] arrival epochs | [slot] my epoch counters | error word
constexpr int kEpochOff = kSlots * kMaxPeers;
constexpr int kErrorOff = kEpochOff + kSlots;

struct PeerBlocks {
  uint32_t* blk[kMaxPeers];
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One CTA, one thread per peer.  Stream order puts every earlier copy / kernel of this rank before the release-store; the
// acquire-load orders every later operation of the stream after the peers' signals.  Epochs only grow (compared as a signed
// difference), so a replayed graph needs no host-side argument: the counter lives in the flag block.
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerBlocks blocks, int rank, int world, int slot,
                                                          unsigned long long timeout_ns) {
  uint32_t* mine = blocks.blk[rank];
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) {
    epoch = mine[kEpochOff + slot] + 1u;
    mine[kEpochOff + slot] = epoch;
  }
  __syncthreads();
  const int p = threadIdx.x;
  if (p >= world || p == rank) return;
  const uint32_t ep = epoch;
  __threadfence_system();
  uint32_t* theirs = blocks.blk[p] + slot * kMaxPeers + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(ep) : "memory");
  const uint32_t* flag = mine + slot * kMaxPeers + p;
  const unsigned long long t0 = global_ns();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (static_cast<int32_t>(v - ep) >= 0) break;
    if (global_ns() - t0 > timeout_ns) {        // a peer never arrived: record it and leave (no trap: the context survives)
      atomicExch(mine + kErrorOff, 1u + static_cast<uint32_t>(p));
      break;
    }
    __nanosleep(64);
  }
}

// dst[i] = sum over the ranks' copies of element i, added in rank order: the local copy (dst itself) stands at position
// `rank`, staged copy s at position s (s < rank) or s + 1.  16 bytes per thread and source.
__global__ void __launch_bounds__(256)
sum_staged_kernel(float4* __restrict__ dst, const float4* __restrict__ staged, int n_staged, int rank, long n_vec,
                  long stride_vec) {
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= n_vec) return;
  const float4 own = dst[i];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  bool first = true;
  for (int pos = 0; pos <= n_staged; ++pos) {
    float4 x;
    if (pos == rank) {
      x = own;
    } else {
      const int s = pos < rank ? pos : pos - 1;
      x = __ldcs(staged + s * stride_vec + i);
    }
    if (first) {
      acc = x;
      first = false;
    } else {
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  dst[i] = acc;
}

__global__ void stamp_kernel(unsigned long long* dst) { *dst = global_ns(); }

// ---- the whole exchange of one range as ONE kernel (the last bucket of a step: nothing is left to overlap with, so the
// SMs are free and what counts is latency -- a copy-engine node costs ~25 us inside a graph, a flag round trip ~3 us here).
//   phase 1: my copy of chunk p -> rank p's staging, 16-byte stores over NVLink          | all 148 CTAs
//   flag 1 : the last CTA to finish releases "arrived" into every peer's flag block; every CTA acquires the peers' flags
//   phase 2: chunk[me] = sum of the copies in rank order -> my slab AND every peer's slab (the all-gather is the store)
//   flag 2 : as flag 1: nobody leaves before its slab is complete, nobody's staging is overwritten while it is being read
// Epochs live in the flag block (bumped by the last CTA to leave), so a replayed graph needs no host argument.  Every CTA
// spins, so all of them must be resident: the grid is one CTA per SM and the kernel is issued where nothing else runs.
constexpr int kFusedOff = 640;                  // uint32 index: [0,16) arrive1, [16,32) arrive2, 32..34 counters, 35 epoch
struct FusedArgs {
  uint32_t* blk[kMaxPeers];
  float* slab[kMaxPeers];                        // start of the range in every rank's slab
  float* stage[kMaxPeers];                       // every rank's staging buffer for this kernel
  int rank, world;
  long n4, c4;                                   // range / chunk length in float4
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void fused_flag_round(const FusedArgs& a, uint32_t* mine, int cnt, int arrive, uint32_t ep) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned old = atomicAdd(mine + kFusedOff + cnt, 1u);
    if (old == gridDim.x - 1) {                  // every CTA of this rank has issued its stores
      mine[kFusedOff + cnt] = 0u;
      __threadfence_system();
      for (int k = 1; k < a.world; ++k) {
        uint32_t* theirs = a.blk[(a.rank + k) % a.world] + kFusedOff + arrive + a.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(ep) : "memory");
      }
    }
  }
  if (threadIdx.x < a.world && static_cast<int>(threadIdx.x) != a.rank) {
    const uint32_t* flag = mine + kFusedOff + arrive + threadIdx.x;
    const unsigned long long t0 = global_ns();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (static_cast<int32_t>(v - ep) >= 0) break;
      if (global_ns() - t0 > a.timeout_ns) {
        atomicExch(mine + kErrorOff, 1u + threadIdx.x);
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) fused_allreduce_kernel(const FusedArgs a) {
  uint32_t* mine = a.blk[a.rank];
  const uint32_t ep = *reinterpret_cast<volatile uint32_t*>(mine + kFusedOff + 35) + 1u;
  const int W = a.world, me = a.rank;
  const long tid = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long nthr = static_cast<long>(gridDim.x) * blockDim.x;
  for (int k = 1; k < W; ++k) {
    const int p = (me + k) % W;
    const long lo = min(p * a.c4, a.n4), m = min(lo + a.c4, a.n4) - lo;
    const float4* src = reinterpret_cast<const float4*>(a.slab[me]) + lo;
    float4* dst = reinterpret_cast<float4*>(a.stage[p]) + (me < p ? me : me - 1) * a.c4;
    for (long i = tid; i < m; i += nthr) dst[i] = src[i];
  }
  fused_flag_round(a, mine, 32, 0, ep);
  {
    const long lo = min(me * a.c4, a.n4), m = min(lo + a.c4, a.n4) - lo;
    const float4* staged = reinterpret_cast<const float4*>(a.stage[me]);
    for (long i = tid; i < m; i += nthr) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int pos = 0; pos < W; ++pos) {
        const float4 x = pos == me ? reinterpret_cast<const float4*>(a.slab[me])[lo + i]
                                   : __ldcg(staged + (pos < me ? pos : pos - 1) * a.c4 + i);
        if (pos == 0) {
          acc = x;
        } else {
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
      }
      for (int k = 0; k < W; ++k) reinterpret_cast<float4*>(a.slab[(me + k) % W])[lo + i] = acc;
    }
  }
  fused_flag_round(a, mine, 33, 16, ep);
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(mine + kFusedOff + 34, 1u);
    if (old == gridDim.x - 1) {                  // the last CTA to leave: the next launch runs in the next epoch
      mine[kFusedOff + 34] = 0u;
      mine[kFusedOff + 35] = ep;
    }
  }
}

}  // namespace

int peer_stamp(unsigned long long* dst, cudaStream_t s) {
  stamp_kernel<<<1, 1, 0, s>>>(dst);
  return check_launch("stamp_kernel");
}

int peer_alloc(size_t bytes, void** ptr) {
  CTI_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) {
    set_error("peer_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_error("peer_free: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int peer_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) {
    set_error("peer_export: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int peer_import(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("peer_import: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    set_error("peer_close: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int peer_barrier(void* const* flag_blocks, int rank, int world, int slot, double timeout_s, cudaStream_t s) {
  CTI_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && slot >= 0 && slot < kSlots,
              "peer_barrier: bad arguments (rank %d of %d, slot %d)", rank, world, slot);
  if (world == 1) return 0;
  PeerBlocks b;
  for (int i = 0; i < kMaxPeers; ++i) b.blk[i] = i < world ? static_cast<uint32_t*>(flag_blocks[i]) : nullptr;
  peer_barrier_kernel<<<1, 32, 0, s>>>(b, rank, world, slot, static_cast<unsigned long long>(timeout_s * 1e9));
  return check_launch("peer_barrier_kernel");
}

// The same barrier without a resident kernel: stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32).  A
// spinning barrier CTA holds a CTA slot of one SM for as long as the slowest rank takes to arrive; the path's persistent
// kernels (one CTA per SM, all of its shared memory) launched meanwhile find 147 free SMs and run their last CTA as a
// second wave (measured: +0.09 ms per step at N = 2).  Memory operations wait in the stream's front end instead.
// The values are immediates, so a replayed graph writes the same ones: the flag is 1 = arrived, and the waiter resets it to
// 0 right after its wait.  That is safe as long as two consecutive barriers never use the same slot: a peer's next signal
// on this slot follows its pass of a barrier on another slot, which needed my signal there -- issued after my reset here.
namespace {
constexpr int kMemopFlagOff = 256;      // uint32 index of the [slot][peer] flags of the memop barrier (byte 1024)
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

typedef CUresult (*BatchFn)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int);

void* driver_ptr(const char* name) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  return ptr;
}

StreamValueFn driver_fn(const char* name) { return reinterpret_cast<StreamValueFn>(driver_ptr(name)); }
}  // namespace

int peer_barrier_memops(void* const* flag_blocks, int rank, int world, int slot, cudaStream_t s) {
  CTI_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && slot >= 0 && slot < kSlots,
              "peer_barrier_memops: bad arguments (rank %d of %d, slot %d)", rank, world, slot);
  if (world == 1) return 0;
  static BatchFn batch_fn = reinterpret_cast<BatchFn>(driver_ptr("cuStreamBatchMemOp"));
  CTI_REQUIRE(batch_fn != nullptr, "peer_barrier_memops: stream memory operations are not available");
  auto flag = [&](int owner, int from) {
    return reinterpret_cast<CUdeviceptr>(static_cast<uint32_t*>(flag_blocks[owner]) + kMemopFlagOff + slot * kMaxPeers + from);
  };
  // one batch = one graph node: signal every peer, then wait for and reset every peer's signal (a node per operation costs
  // a few microseconds each -- 21 of them per barrier at eight ranks)
  CUstreamBatchMemOpParams ops[3 * kMaxPeers];
  memset(ops, 0, sizeof(ops));
  int n = 0;
  for (int k = 1; k < world; ++k) {
    ops[n].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
    ops[n].writeValue.address = flag((rank + k) % world, rank);
    ops[n].writeValue.value = 1u;
    ops[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
    ++n;
  }
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;
    ops[n].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
    ops[n].waitValue.address = flag(rank, p);
    ops[n].waitValue.value = 1u;
    ops[n].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
    ++n;
    ops[n].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
    ops[n].writeValue.address = flag(rank, p);
    ops[n].writeValue.value = 0u;
    ops[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
    ++n;
  }
  const CUresult r = batch_fn(reinterpret_cast<CUstream>(s), static_cast<unsigned>(n), ops, 0);
  if (r != CUDA_SUCCESS) {
    set_error("peer_barrier_memops: driver error %d", static_cast<int>(r));
    return static_cast<int>(r);
  }
  return 0;
}

// `count` operations on words of ONE flag block (local or a peer's) as a single batch: ops[i] = {index, value, wait}.
int peer_flag_ops(void* flag_block, const int* index, const uint32_t* value, const int* wait, int count, cudaStream_t s) {
  CTI_REQUIRE(flag_block != nullptr && count >= 1 && count <= 16, "peer_flag_ops: bad arguments");
  static BatchFn batch_fn = reinterpret_cast<BatchFn>(driver_ptr("cuStreamBatchMemOp"));
  CTI_REQUIRE(batch_fn != nullptr, "peer_flag_ops: stream memory operations are not available");
  CUstreamBatchMemOpParams ops[16];
  memset(ops, 0, sizeof(ops));
  for (int i = 0; i < count; ++i) {
    CTI_REQUIRE(index[i] >= 0 && index[i] < 1024, "peer_flag_ops: bad flag index %d", index[i]);
    const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(static_cast<uint32_t*>(flag_block) + index[i]);
    if (wait[i]) {
      ops[i].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
      ops[i].waitValue.address = a;
      ops[i].waitValue.value = value[i];
      ops[i].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
    } else {
      ops[i].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
      ops[i].writeValue.address = a;
      ops[i].writeValue.value = value[i];
      ops[i].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
    }
  }
  const CUresult r = batch_fn(reinterpret_cast<CUstream>(s), static_cast<unsigned>(count), ops, 0);
  if (r != CUDA_SUCCESS) {
    set_error("peer_flag_ops: driver error %d", static_cast<int>(r));
    return static_cast<int>(r);
  }
  return 0;
}

int peer_flag_op(void* flag_block, int index, uint32_t value, int wait, cudaStream_t s) {
  CTI_REQUIRE(flag_block != nullptr && index >= 0 && index < 1024, "peer_flag_op: bad arguments");
  static StreamValueFn write_fn = driver_fn("cuStreamWriteValue32");
  static StreamValueFn wait_fn = driver_fn("cuStreamWaitValue32");
  CTI_REQUIRE(write_fn != nullptr && wait_fn != nullptr, "peer_flag_op: stream memory operations are not available");
  const CUdeviceptr a = reinterpret_cast<CUdeviceptr>(static_cast<uint32_t*>(flag_block) + index);
  const CUresult r = wait ? wait_fn(reinterpret_cast<CUstream>(s), a, value, CU_STREAM_WAIT_VALUE_GEQ)
                          : write_fn(reinterpret_cast<CUstream>(s), a, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) {
    set_error("peer_flag_op: driver error %d", static_cast<int>(r));
    return static_cast<int>(r);
  }
  return 0;
}

int peer_error(const void* flag_block, int* out) {
  uint32_t v = 0;
  cudaError_t e = cudaMemcpy(&v, static_cast<const uint32_t*>(flag_block) + kErrorOff, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_error("peer_error: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  *out = static_cast<int>(v);
  return 0;
}

int peer_allreduce_fused(void* const* flag_blocks, void* const* slab_ranges, void* const* stagings, int rank, int world,
                         long n, double timeout_s, cudaStream_t s) {
  CTI_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && n >= 0 && (n & 3) == 0,
              "peer_allreduce_fused: bad arguments (rank %d of %d, n=%ld)", rank, world, n);
  if (world == 1 || n == 0) return 0;
  FusedArgs a;
  for (int i = 0; i < kMaxPeers; ++i) {
    a.blk[i] = i < world ? static_cast<uint32_t*>(flag_blocks[i]) : nullptr;
    a.slab[i] = i < world ? static_cast<float*>(slab_ranges[i]) : nullptr;
    a.stage[i] = i < world ? static_cast<float*>(stagings[i]) : nullptr;
    CTI_REQUIRE(i >= world || (((uintptr_t)a.slab[i] | (uintptr_t)a.stage[i]) & 15) == 0,
                "peer_allreduce_fused: buffers must be 16-byte aligned");
  }
  a.rank = rank;
  a.world = world;
  a.n4 = n / 4;
  a.c4 = (a.n4 + world - 1) / world;
  a.timeout_ns = static_cast<unsigned long long>(timeout_s * 1e9);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  fused_allreduce_kernel<<<sms, 512, 0, s>>>(a);
  return check_launch("fused_allreduce_kernel");
}

int peer_copy(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return 0;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s);
  if (e != cudaSuccess) {
    set_error("peer_copy(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

int sum_staged(float* dst, const float* staged, int n_staged, int rank, long n, long stride, cudaStream_t s) {
  CTI_REQUIRE(n >= 0 && (n & 3) == 0 && (stride & 3) == 0 && n_staged >= 0 && rank >= 0 && rank <= n_staged,
              "sum_staged: bad arguments (n=%ld stride=%ld n_staged=%d rank=%d)", n, stride, n_staged, rank);
  if (n == 0 || n_staged == 0) return 0;
  CTI_REQUIRE(((uintptr_t)dst & 15) == 0 && ((uintptr_t)staged & 15) == 0, "sum_staged: buffers must be 16-byte aligned");
  const long nv = n / 4;
  sum_staged_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, s>>>(reinterpret_cast<float4*>(dst),
                                                                 reinterpret_cast<const float4*>(staged), n_staged, rank, nv,
                                                                 stride / 4);
  return check_launch("sum_staged_kernel");
}

}  // namespace cti
