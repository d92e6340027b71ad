// GRU cell pointwise stages: the step before the hot path, `QuestionEmbedding`'s
// one-layer unidirectional nn.GRU (reference src/language_model.py:56-61,93-98) that produces q and a.
// The matrix products run on the tcgen05 GEMM (all timesteps of x W_ih^T at once, h W_hh^T per step); these two
// kernels are everything else of a timestep, fused:
//   forward   r = sigmoid(gx_r + gh_r), z = sigmoid(gx_z + gh_z), n = tanh(gx_n + r * gh_n), h' = (1 - z) n + z h
//   backward  the derivative of the above w.r.t. gx (3H), gh (3H) and h
// gx / gh already contain b_ih / b_hh (GEMM epilogue).  fp32 gate math; h' is written as fp32 into the (B, T, H)
// output and as bf16 for the next step's GEMM; r, z, n and gh_n are kept in bf16 for the backward pass.
// HBM-bound: forward reads 6H*4 + H*4 and writes H*(4 + 2 + 4*2) bytes per row.
#include "cti_common.cuh"
#include "cti_kernels.h"

namespace cti {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256)
gru_gate_fwd_kernel(const float* __restrict__ gx, long gx_row_stride, const float* __restrict__ gh,
                    const float* __restrict__ h_prev, long hp_row_stride, float* __restrict__ h_out, long ho_row_stride,
                    __nv_bfloat16* __restrict__ h_bf16, __nv_bfloat16* __restrict__ r_s, __nv_bfloat16* __restrict__ z_s,
                    __nv_bfloat16* __restrict__ n_s, __nv_bfloat16* __restrict__ ghn_s, long rows, int H) {
  pdl_prologue_done();
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= rows * H) return;
  const long b = i / H;
  const int c = static_cast<int>(i - b * H);
  const float* gxr = gx + b * gx_row_stride;
  const float* ghr = gh + b * 3l * H;
  const float r = sigmoidf_(gxr[c] + ghr[c]);
  const float z = sigmoidf_(gxr[H + c] + ghr[H + c]);
  const float ghn = ghr[2 * H + c];
  const float n = tanhf(gxr[2 * H + c] + r * ghn);
  const float hp = h_prev != nullptr ? h_prev[b * hp_row_stride + c] : 0.f;
  const float h = (1.f - z) * n + z * hp;
  h_out[b * ho_row_stride + c] = h;
  h_bf16[i] = __float2bfloat16(h);
  r_s[i] = __float2bfloat16(r);
  z_s[i] = __float2bfloat16(z);
  n_s[i] = __float2bfloat16(n);
  ghn_s[i] = __float2bfloat16(ghn);
}

// dh (rows, H) fp32 is the total gradient w.r.t. h_t on entry (carried + this step's output gradient, added here) and
// the direct part of the gradient w.r.t. h_{t-1} (dh * z) on exit; the W_hh part is added by the following GEMM.
__global__ void __launch_bounds__(256)
gru_gate_bwd_kernel(float* __restrict__ dh, const float* __restrict__ dout, long do_row_stride,
                    const float* __restrict__ h_prev, long hp_row_stride, const __nv_bfloat16* __restrict__ r_s,
                    const __nv_bfloat16* __restrict__ z_s, const __nv_bfloat16* __restrict__ n_s,
                    const __nv_bfloat16* __restrict__ ghn_s, __nv_bfloat16* __restrict__ dgx, long dgx_row_stride,
                    __nv_bfloat16* __restrict__ dgh, long rows, int H) {
  pdl_prologue_done();
  const long i = blockIdx.x * 256l + threadIdx.x;
  if (i >= rows * H) return;
  const long b = i / H;
  const int c = static_cast<int>(i - b * H);
  const float d = dh[i] + dout[b * do_row_stride + c];
  const float r = __bfloat162float(r_s[i]), z = __bfloat162float(z_s[i]), n = __bfloat162float(n_s[i]);
  const float ghn = __bfloat162float(ghn_s[i]);
  const float hp = h_prev != nullptr ? h_prev[b * hp_row_stride + c] : 0.f;
  const float dn_pre = d * (1.f - z) * (1.f - n * n);
  const float dz_pre = d * (hp - n) * z * (1.f - z);
  const float dr_pre = dn_pre * ghn * r * (1.f - r);
  __nv_bfloat16* gxo = dgx + b * dgx_row_stride;
  __nv_bfloat16* gho = dgh + b * 3l * H;
  const __nv_bfloat16 drb = __float2bfloat16(dr_pre), dzb = __float2bfloat16(dz_pre);
  gxo[c] = drb;
  gxo[H + c] = dzb;
  gxo[2 * H + c] = __float2bfloat16(dn_pre);
  gho[c] = drb;
  gho[H + c] = dzb;
  gho[2 * H + c] = __float2bfloat16(dn_pre * r);
  dh[i] = d * z;
}

}  // namespace

int gru_gate_fwd(const float* gx, long gx_row_stride, const float* gh, const float* h_prev, long hp_row_stride, float* h_out,
                 long ho_row_stride, __nv_bfloat16* h_bf16, __nv_bfloat16* r_s, __nv_bfloat16* z_s, __nv_bfloat16* n_s,
                 __nv_bfloat16* ghn_s, long rows, int H, cudaStream_t s) {
  CTI_REQUIRE(rows > 0 && H > 0, "gru_gate_fwd: empty problem");
  const long n = rows * H;
  CTI_REQUIRE((n + 255) / 256 < (1l << 31), "gru_gate_fwd: too many elements");
  launch_pdl(gru_gate_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, gx, gx_row_stride, gh, h_prev, hp_row_stride, h_out,
                                                                  ho_row_stride, h_bf16, r_s, z_s, n_s, ghn_s, rows, H);
  return check_launch("gru_gate_fwd_kernel");
}

int gru_gate_bwd(float* dh, const float* dout, long do_row_stride, const float* h_prev, long hp_row_stride,
                 const __nv_bfloat16* r_s, const __nv_bfloat16* z_s, const __nv_bfloat16* n_s, const __nv_bfloat16* ghn_s,
                 __nv_bfloat16* dgx, long dgx_row_stride, __nv_bfloat16* dgh, long rows, int H, cudaStream_t s) {
  CTI_REQUIRE(rows > 0 && H > 0, "gru_gate_bwd: empty problem");
  const long n = rows * H;
  CTI_REQUIRE((n + 255) / 256 < (1l << 31), "gru_gate_bwd: too many elements");
  launch_pdl(gru_gate_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, dh, dout, do_row_stride, h_prev, hp_row_stride, r_s, z_s,
                                                                  n_s, ghn_s, dgx, dgx_row_stride, dgh, rows, H);
  return check_launch("gru_gate_bwd_kernel");
}

}  // namespace cti
