// Re-layout kernels of the constructor variants (encode_rotmat, share_feature):
//  * rmv_strided_copy  -- 3-D strided copy with dtype conversion, optional per-column scale and
//                         accumulation: the zero-padded corners of the 3593-wide layers
//                         (models/rot_mv.py:53-67), the 9 rotation entries appended to a fuser row
//                         (:225-231), the [3][2][512] interleave of RotFeatFuser's input and of the
//                         head input (:80-84, 243-248) and their backward scatter/accumulate.
//  * rmv_intensity_bn_train -- train-mode IntensityBatchNorm statistics of one call (:13-32).
//  * rmv_fill_zero     -- zero fill (gradient / padding buffers).
// All HBM-bound byte movers; none of them is on the path main.py builds (the default
// ImageFeatFuser configuration is concat-free and needs no re-layout at all).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ops.h"

namespace {

using rmv::griddep_launch;
using rmv::griddep_wait;

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

struct Strides3 {
  long long s0, s1, s2;
};

// one element per thread, i2 fastest: coalesced whenever both innermost strides are 1
template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
strided_copy_kernel(const TS* __restrict__ src, Strides3 ss, TD* __restrict__ dst, Strides3 ds,
                    int n1, int n2, long long total, const float* __restrict__ scale,
                    int accumulate) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i2 = (int)(idx % n2);
  const long long t = idx / n2;
  const int i1 = (int)(t % n1);
  const long long i0 = t / n1;
  float v = ld_f(src + i0 * ss.s0 + i1 * ss.s1 + i2 * ss.s2);
  if (scale != nullptr) v = __fmul_rn(v, scale[i2]);   // never contracted into an FMA with the
  TD* d = dst + i0 * ds.s0 + i1 * ds.s1 + i2 * ds.s2;
  if (accumulate) v = __fadd_rn(v, ld_f(d));           // accumulation: product and sum round separately
  st_f(d, v);
}

// eight consecutive i2 elements per thread (16-byte accesses) when both innermost strides are 1 and
// every offset is a multiple of eight elements: the interleave / padded-corner / output-assembly copies
template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = rmv::unpack_bf16x2(w[i]);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float* f);
template <> __device__ __forceinline__ void st8<float>(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  uint4 u;
  u.x = rmv::pack_bf16x2(f[0], f[1]); u.y = rmv::pack_bf16x2(f[2], f[3]);
  u.z = rmv::pack_bf16x2(f[4], f[5]); u.w = rmv::pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
strided_copy_vec8_kernel(const TS* __restrict__ src, Strides3 ss, TD* __restrict__ dst, Strides3 ds,
                         int n1, int n2v, long long total, const float* __restrict__ scale,
                         int accumulate) {
  griddep_wait();
  griddep_launch();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over groups of 8
  if (idx >= total) return;
  const int i2 = (int)(idx % n2v) * 8;
  const long long t = idx / n2v;
  const int i1 = (int)(t % n1);
  const long long i0 = t / n1;
  float v[8];
  ld8(src + i0 * ss.s0 + i1 * ss.s1 + i2, v);
  if (scale != nullptr) {
    float sc[8];
    ld8(scale + i2, sc);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fmul_rn(v[i], sc[i]);
  }
  TD* d = dst + i0 * ds.s0 + i1 * ds.s1 + i2;
  if (accumulate) {
    float a[8];
    ld8(d, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], a[i]);
  }
  st8(d, v);
}

// transposing form: the source is contiguous along i1 and the destination along i2 (a weight matrix
// and its transpose). 32 x 32 tiles through shared memory, coalesced on both sides; grid.z = i0.
template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
strided_copy_transpose_kernel(const TS* __restrict__ src, Strides3 ss, TD* __restrict__ dst,
                              Strides3 ds, int n1, int n2, const float* __restrict__ scale,
                              int accumulate) {
  griddep_wait();
  griddep_launch();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int b1 = blockIdx.x * 32, b2 = blockIdx.y * 32;
  const long long i0 = blockIdx.z;
  const TS* s = src + i0 * ss.s0;
  TD* d = dst + i0 * ds.s0;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {  // r walks i2, tx walks i1 (source-contiguous)
    const int i1 = b1 + tx, i2 = b2 + r;
    tile[r][tx] = (i1 < n1 && i2 < n2) ? ld_f(s + (long long)i1 * ss.s1 + (long long)i2 * ss.s2) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {  // r walks i1, tx walks i2 (destination-contiguous)
    const int i1 = b1 + r, i2 = b2 + tx;
    if (i1 < n1 && i2 < n2) {
      float v = tile[tx][r];
      if (scale != nullptr) v = __fmul_rn(v, scale[i2]);
      TD* p = d + (long long)i1 * ds.s1 + (long long)i2 * ds.s2;
      if (accumulate) v = __fadd_rn(v, ld_f(p));
      st_f(p, v);
    }
  }
}

// IntensityBatchNorm (train): one block = 32 feature vectors (columns) x 8 row lanes. Sums in
// fp64 (the variance of a few hundred norms of similar size; no cancellation to speak of).
template <typename T>
__global__ void __launch_bounds__(256)
intensity_bn_kernel(const T* __restrict__ feat, long long ld, int rows, int nvec,
                    float* __restrict__ running, float momentum, float eps,
                    float* __restrict__ scale_out) {
  griddep_wait();
  griddep_launch();
  __shared__ double s_sum[8][33], s_sq[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  double sum = 0.0, sq = 0.0;
  if (j < nvec) {
    for (int r = ty; r < rows; r += 8) {
      const T* p = feat + (long long)r * ld + j;
      const float a = ld_f(p), b = ld_f(p + nvec), c = ld_f(p + 2 * nvec);
      const float nrm = sqrtf(a * a + b * b + c * c);   // torch.norm(x, dim=-2)
      sum += (double)nrm;
      sq += (double)nrm * (double)nrm;
    }
  }
  s_sum[ty][tx] = sum;
  s_sq[ty][tx] = sq;
  __syncthreads();
  if (ty == 0 && j < nvec) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { sum += s_sum[k][tx]; sq += s_sq[k][tx]; }
    const double mean = sum / rows;
    double var = sq / rows - mean * mean;  // biased (unbiased=False)
    if (var < 0.0) var = 0.0;
    const float sd = sqrtf(fmaxf((float)var, eps));
    const float run = running[j] * (1.f - momentum) + sd * momentum;
    running[j] = run;
    scale_out[j] = 1.f / (run + eps);
  }
}

// bytes [0, n_head) and the n_tail bytes after the 16-byte body are written byte-wise by block 0
__global__ void __launch_bounds__(256) fill_zero_kernel(unsigned char* __restrict__ head, int n_head,
                                                        uint4* __restrict__ body, long long n16,
                                                        unsigned char* __restrict__ tail,
                                                        int n_tail) {
  griddep_wait();
  griddep_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
    body[i] = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < n_head) head[threadIdx.x] = 0;
    if ((int)threadIdx.x < n_tail) tail[threadIdx.x] = 0;
  }
}

template <typename TS, typename TD>
int launch_copy(const void* src, Strides3 ss, void* dst, Strides3 ds, int n0, int n1, int n2,
                const float* scale, int accumulate, cudaStream_t stream) {
  const bool transposing = ss.s2 != 1 && ss.s1 == 1 && ds.s2 == 1 && n1 >= 16 && n2 >= 16 && n0 <= 65535;
  const auto mult8 = [](long long x) { return (x & 7) == 0; };
  const bool vec8 = ss.s2 == 1 && ds.s2 == 1 && (n2 & 7) == 0 &&
                    (n1 == 1 || (mult8(ss.s1) && mult8(ds.s1))) && (n0 == 1 || (mult8(ss.s0) && mult8(ds.s0))) &&
                    ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
                      reinterpret_cast<uintptr_t>(scale)) & 15) == 0;
  if (vec8) {
    const long long total = (long long)n0 * n1 * (n2 / 8);
    const long long blocks = (total + 255) / 256;
    RMV_CHECK_ARG(blocks <= 0x7fffffffLL, "strided_copy: too many elements");
    RMV_CUDA(rmv::launch_pdl(strided_copy_vec8_kernel<TS, TD>, dim3((unsigned)blocks), dim3(256), 0, stream,
                             (const TS*)src, ss, (TD*)dst, ds, n1, n2 / 8, total, scale, accumulate));
  } else if (transposing) {
    dim3 grid((unsigned)((n1 + 31) / 32), (unsigned)((n2 + 31) / 32), (unsigned)n0);
    RMV_CHECK_ARG(grid.y <= 65535u, "strided_copy: n2 too large for the transposing form");
    RMV_CUDA(rmv::launch_pdl(strided_copy_transpose_kernel<TS, TD>, grid, dim3(256), 0, stream,
                             (const TS*)src, ss, (TD*)dst, ds, n1, n2, scale, accumulate));
  } else {
    const long long total = (long long)n0 * n1 * n2;
    const long long blocks = (total + 255) / 256;
    RMV_CHECK_ARG(blocks <= 0x7fffffffLL, "strided_copy: too many elements");
    RMV_CUDA(rmv::launch_pdl(strided_copy_kernel<TS, TD>, dim3((unsigned)blocks), dim3(256), 0, stream,
                             (const TS*)src, ss, (TD*)dst, ds, n1, n2, total, scale, accumulate));
  }
  RMV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int rmv_strided_copy(const void* src, int src_dtype, long long ss0, long long ss1,
                                long long ss2, void* dst, int dst_dtype, long long ds0,
                                long long ds1, long long ds2, int n0, int n1, int n2,
                                const float* scale, int accumulate, void* stream) {
  RMV_CHECK_ARG(n0 >= 0 && n1 >= 0 && n2 >= 0, "strided_copy: negative extent");
  if ((long long)n0 * n1 * n2 == 0) return 0;
  RMV_CHECK_ARG(src != nullptr && dst != nullptr, "strided_copy: null tensor pointer");
  RMV_CHECK_ARG((src_dtype == RMV_DTYPE_F32 || src_dtype == RMV_DTYPE_BF16) &&
                    (dst_dtype == RMV_DTYPE_F32 || dst_dtype == RMV_DTYPE_BF16),
                "strided_copy: dtypes must be RMV_DTYPE_F32 or RMV_DTYPE_BF16");
  const Strides3 ss{ss0, ss1, ss2}, ds{ds0, ds1, ds2};
  cudaStream_t st = (cudaStream_t)stream;
  const bool sb = src_dtype == RMV_DTYPE_BF16, db = dst_dtype == RMV_DTYPE_BF16;
  if (sb && db) return launch_copy<__nv_bfloat16, __nv_bfloat16>(src, ss, dst, ds, n0, n1, n2, scale, accumulate, st);
  if (sb) return launch_copy<__nv_bfloat16, float>(src, ss, dst, ds, n0, n1, n2, scale, accumulate, st);
  if (db) return launch_copy<float, __nv_bfloat16>(src, ss, dst, ds, n0, n1, n2, scale, accumulate, st);
  return launch_copy<float, float>(src, ss, dst, ds, n0, n1, n2, scale, accumulate, st);
}

extern "C" int rmv_intensity_bn_train(const void* feat, long long ld, int dtype, int rows, int nvec,
                                      float* running, float momentum, float eps, float* scale_out,
                                      void* stream) {
  RMV_CHECK_ARG(feat && running && scale_out, "intensity_bn_train: null pointer");
  RMV_CHECK_ARG(rows > 0 && nvec > 0 && ld >= 3LL * nvec, "intensity_bn_train: bad shape (rows=%d nvec=%d ld=%lld)",
                rows, nvec, ld);
  RMV_CHECK_ARG(dtype == RMV_DTYPE_F32 || dtype == RMV_DTYPE_BF16, "intensity_bn_train: bad dtype");
  const dim3 grid((unsigned)((nvec + 31) / 32));
  if (dtype == RMV_DTYPE_BF16) {
    RMV_CUDA(rmv::launch_pdl(intensity_bn_kernel<__nv_bfloat16>, grid, dim3(256), 0, (cudaStream_t)stream,
                             (const __nv_bfloat16*)feat, ld, rows, nvec, running, momentum, eps, scale_out));
  } else {
    RMV_CUDA(rmv::launch_pdl(intensity_bn_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream,
                             (const float*)feat, ld, rows, nvec, running, momentum, eps, scale_out));
  }
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_fill_zero(void* dst, size_t bytes, void* stream) {
  if (bytes == 0) return 0;
  RMV_CHECK_ARG(dst != nullptr, "fill_zero: null pointer");
  unsigned char* base = (unsigned char*)dst;
  size_t n_head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (n_head > bytes) n_head = bytes;
  const long long n16 = (long long)((bytes - n_head) / 16);
  const int n_tail = (int)((bytes - n_head) % 16);
  long long blocks = (n16 + 256 * 8 - 1) / (256 * 8);   // ~8 stores per thread
  const long long cap = (long long)rmv::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  RMV_CUDA(rmv::launch_pdl(fill_zero_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                           base, (int)n_head, (uint4*)(base + n_head), n16,
                           base + n_head + n16 * 16, n_tail));
  RMV_LAUNCH_CHECK();
  return 0;
}
