// FFMA implicit-GEMM convolution / linear layer (fp32 accumulation, fp32 or bf16 storage).
//
// This is the fp32 parity engine (north_star: "fp32 rtol 1e-4 on features and logits" cannot be met
// by a 10-bit-mantissa tensor-core format through 53 layers) and the on-device reference the
// tcgen05 kernel in igemm_sm100.cu is checked against. Same operator contract as conv_fwd_tc:
// y = act(scale * conv(x, w) + shift + residual), any kernel size / stride / pad / channel count,
// arbitrary input strides (so NCHW fp32 images are consumed without a layout pass).
//
// Reference sites: models/resnet.py:31-47,128-148,184-189 (Conv2d/BN/ReLU),
// models/backbones/blocks.py:41-60 (Linear/ReLU).
#include "common.cuh"
#include "ops.h"

namespace rmv {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
  const void* x; const void* w; void* y; const void* residual;
  const float* scale; const float* shift;
  long long x_sn, x_sh, x_sw, x_sc;
  long long y_sn, y_sh, y_sw;
  long long r_sn, r_sh, r_sw;
  int n_img, in_h, in_w, c_in;
  int c_out, kh, kw, stride, pad;
  int out_h, out_w;
  int relu;
  long long m_total;
  int k_total;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) simt_conv_kernel(const SimtArgs a) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const TIn* __restrict__ x = reinterpret_cast<const TIn*>(a.x);
  const TIn* __restrict__ w = reinterpret_cast<const TIn*>(a.w);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;

  // global->smem assignment: one (row, 4 consecutive k) strip of A and of B per thread
  const int l_row = tid >> 2;
  const int l_k0 = (tid & 3) * 4;
  const long long lm = m0 + l_row;
  const bool m_ok = lm < a.m_total;
  int ln = 0, loh = 0, low = 0;
  if (m_ok) {
    low = (int)(lm % a.out_w);
    const long long t = lm / a.out_w;
    loh = (int)(t % a.out_h);
    ln = (int)(t / a.out_h);
  }
  const int ih0 = loh * a.stride - a.pad, iw0 = low * a.stride - a.pad;
  const long long x_base = (long long)ln * a.x_sn;
  const int wn = n0 + l_row;
  const bool n_ok = wn < a.c_out;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.k_total; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + l_k0 + i;
      float av = 0.f, bv = 0.f;
      if (k < a.k_total) {
        if (m_ok) {
          const int tap = k / a.c_in;
          const int c = k - tap * a.c_in;
          const int r = tap / a.kw;
          const int s = tap - r * a.kw;
          const int ih = ih0 + r, iw = iw0 + s;
          if (ih >= 0 && ih < a.in_h && iw >= 0 && iw < a.in_w)
            av = ld_as_float(x + x_base + ih * a.x_sh + iw * a.x_sw + c * a.x_sc);
        }
        if (n_ok) bv = ld_as_float(w + (long long)wn * a.k_total + k);
      }
      As[l_k0 + i][l_row] = av;
      Bs[l_k0 + i][l_row] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float am[4] = {av.x, av.y, av.z, av.w};
      const float bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    __syncthreads();
  }

  TOut* __restrict__ y = reinterpret_cast<TOut*>(a.y);
  const TOut* __restrict__ res = reinterpret_cast<const TOut*>(a.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= a.m_total) continue;
    const int ow = (int)(m % a.out_w);
    const long long t = m / a.out_w;
    const int oh = (int)(t % a.out_h);
    const int n = (int)(t / a.out_h);
    const long long yo = n * a.y_sn + oh * a.y_sh + ow * a.y_sw;
    const long long ro = n * a.r_sn + oh * a.r_sh + ow * a.r_sw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = n0 + tx * 4 + j;
      if (k >= a.c_out) continue;
      float v = acc[i][j];
      if (a.scale) v *= __ldg(a.scale + k);
      if (a.shift) v += __ldg(a.shift + k);
      if (res) v += ld_as_float(res + ro + k);
      if (a.relu) v = fmaxf(v, 0.f);
      st_from_float(y + yo + k, v);
    }
  }
}

}  // namespace

int conv_fwd_simt(const ConvArgs& p, cudaStream_t stream) {
  SimtArgs a;
  a.x = p.x; a.w = p.w; a.y = p.y; a.residual = p.residual;
  a.scale = p.scale; a.shift = p.shift;
  a.x_sn = p.x_sn; a.x_sh = p.x_sh; a.x_sw = p.x_sw; a.x_sc = p.x_sc ? p.x_sc : 1;
  a.y_sn = p.y_sn; a.y_sh = p.y_sh; a.y_sw = p.y_sw;
  a.r_sn = p.r_sn; a.r_sh = p.r_sh; a.r_sw = p.r_sw;
  a.n_img = p.n_img; a.in_h = p.in_h; a.in_w = p.in_w; a.c_in = p.c_in;
  a.c_out = p.c_out; a.kh = p.kh; a.kw = p.kw; a.stride = p.stride; a.pad = p.pad;
  a.out_h = p.out_h; a.out_w = p.out_w; a.relu = p.relu;
  a.m_total = (long long)p.n_img * p.out_h * p.out_w;
  a.k_total = p.kh * p.kw * p.c_in;
  if (a.m_total == 0 || p.c_out == 0) return 0;
  dim3 grid((unsigned)ceil_div(a.m_total, TM), (unsigned)ceil_div(p.c_out, TN));
  RMV_CHECK_ARG(grid.y <= 65535, "simt conv: c_out too large");
  if (p.x_dtype == RMV_DTYPE_F32 && p.y_dtype == RMV_DTYPE_F32)
    simt_conv_kernel<float, float><<<grid, 256, 0, stream>>>(a);
  else if (p.x_dtype == RMV_DTYPE_BF16 && p.y_dtype == RMV_DTYPE_BF16)
    simt_conv_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(a);
  else if (p.x_dtype == RMV_DTYPE_BF16 && p.y_dtype == RMV_DTYPE_F32)
    simt_conv_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(a);
  else if (p.x_dtype == RMV_DTYPE_F32 && p.y_dtype == RMV_DTYPE_BF16)
    simt_conv_kernel<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(a);
  else
    RMV_CHECK_ARG(false, "simt conv: bad dtype combination %d/%d", p.x_dtype, p.y_dtype);
  RMV_LAUNCH_CHECK();
  return 0;
}

}  // namespace rmv
