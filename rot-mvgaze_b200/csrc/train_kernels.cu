// Training-side kernels of the Rot-MV path (all HBM-bound except the FFMA weight gradient):
// batch-statistic BatchNorm forward/backward with PER-VIEW statistics (the reference runs the trunk
// once per view, models/rot_mv.py:196-197), pooling backward, ReLU masks, bias/column
// sums, the analytic angular-loss gradient, the gaze-head tail backward, weight layout transforms,
// gradient dilation for stride-2 dgrad, the FFMA weight-gradient GEMM and the fused multi-tensor
// Adam step (coupled L2 = torch.optim.Adam(weight_decay) as in trainer.py:54, or decoupled).
#include "common.cuh"
#include "bn_finalize.cuh"
#include "ops.h"

namespace rmv {
namespace {

constexpr float kRadToDeg = 57.29577951308232f;

template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
  typedef uint4 Raw;  // packed form: keeps unrolled loads in 4 registers until they are consumed
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) {
    return *reinterpret_cast<const uint4*>(p);
  }
  static __device__ __forceinline__ void unpack(const Raw& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = unpack_bf16x2(w[i]);
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = unpack_bf16x2(w[i]);
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct V8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float* f) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
    f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

inline unsigned nblk(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------
// BatchNorm (train): per-(view, channel) statistics. Image n belongs to view n % views.
// Partial sums are fp32 inside a thread (<= rows_per_block terms), fp64 across threads/blocks.
// ---------------------------------------------------------------------------------------------
// MASK: 0 = no ReLU mask, 1 = y_mask is the ReLU output tensor (mask = y > 0), 2 = y_mask is the
// packed bit mask bn_apply wrote (one byte per 8-channel vector, bit e = channel e passed the ReLU).
__device__ __forceinline__ void apply_bits(float* d, uint32_t bits) {
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = (bits >> i) & 1u ? d[i] : 0.f;
}

// grid = (G, views) with G*views = the number of co-resident blocks (one balanced wave, no tail):
// block (x, v) reduces a contiguous slab of the rows (image-of-view-v, pixel) of view v;
// block = 256 threads = (c/8 channel groups) x (256/(c/8)) row lanes; each thread keeps U 16-byte
// loads per stream in flight and tracks (image, pixel) incrementally -- no division in the loop.
template <typename T, bool BWD, int MASK>
__global__ void __launch_bounds__(256, 3)
bn_reduce_kernel(const T* __restrict__ z, const T* __restrict__ dy, const void* __restrict__ y_mask_v,
                 const float* __restrict__ mean, const float* __restrict__ invstd, int pix, int c,
                 int views, int imgs_per_view, double* __restrict__ acc /* [views][c][2] */,
                 const BnFinalize fin) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  extern __shared__ float s_red[];  // [row_lanes][c][2]
  __shared__ unsigned int s_ticket;
  const T* y_mask = reinterpret_cast<const T*>(y_mask_v);
  const uint8_t* y_bits = reinterpret_cast<const uint8_t*>(y_mask_v);
  // blockIdx.z = channel slice of cs = c / gridDim.z channels (<= 256): wide layers are split
  // along the channels instead of the rows, so that the number of fp64 atomics per launch
  // (blocks x slice channels x 2) does not grow with c
  const int cs = c / gridDim.z, c0 = blockIdx.z * cs;
  const int cg = cs / 8;
  const int lanes = blockDim.x / cg;  // row lanes per block (>= 1)
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int v = blockIdx.y;
  const long long rows_total = (long long)imgs_per_view * pix;
  const long long r0 = rows_total * blockIdx.x / gridDim.x;
  const long long r1 = rows_total * (blockIdx.x + 1) / gridDim.x;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  // BWD accumulates the raw sums S1 = sum d and S2' = sum d*z; sum d*xhat = invstd*(S2' - mean*S1)
  // is formed in fp64 when the block's partial sums are combined (keeps mean/invstd out of the
  // streaming loop: 16 fewer live registers, one FMA per element instead of three operations).
  // independent 16-byte (bf16) / 32-byte (fp32) loads in flight per tensor
  constexpr int U = sizeof(T) == 2 ? (BWD ? 4 : 8) : (BWD ? 2 : 4);
  if (lane < lanes && r0 + lane < r1) {
    // this thread's rows: r0 + lane + k*lanes; (img, p) of the current row, element offset `off`
    long long img = (r0 + lane) / pix;
    int p = (int)((r0 + lane) - img * pix);
    long long off = ((img * views + v) * pix + p) * c + c0 + g * 8;
    const long long row_step = (long long)lanes * c;                 // next row, same image
    const long long wrap_step = (long long)(views - 1) * pix * c;    // extra when the image wraps
    long long left = (r1 - (r0 + lane) + lanes - 1) / lanes;          // rows this thread owns
    while (left > 0) {
      typename V8<T>::Raw zr[U], dr[U], mr[U];
      uint32_t mb[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (j < left) {
          zr[j] = V8<T>::load_raw(z + off);
          if (BWD) {
            dr[j] = V8<T>::load_raw(dy + off);
            if (MASK == 1) mr[j] = V8<T>::load_raw(y_mask + off);
            if (MASK == 2) mb[j] = y_bits[off >> 3];
          }
          off += row_step;
          p += lanes;
          if (p >= pix) { p -= pix; off += wrap_step; }  // lanes <= 32 < pix: at most one wrap
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (j < left) {
          float f[8];
          V8<T>::unpack(zr[j], f);
          if (!BWD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1[i] += f[i]; s2[i] = fmaf(f[i], f[i], s2[i]); }
          } else {
            float d[8];
            V8<T>::unpack(dr[j], d);
            if (MASK == 1) {
              float m[8];
              V8<T>::unpack(mr[j], m);
#pragma unroll
              for (int i = 0; i < 8; ++i) d[i] = m[i] > 0.f ? d[i] : 0.f;
            }
            if (MASK == 2) apply_bits(d, mb[j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              s1[i] += d[i];
              s2[i] = fmaf(d[i], f[i], s2[i]);
            }
          }
        }
      }
      left -= U;
    }
  }
  // reduce over row lanes through shared memory, then one fp64 atomic per (channel, stat)
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s_red[(lane * cs + g * 8 + i) * 2] = s1[i];
      s_red[(lane * cs + g * 8 + i) * 2 + 1] = s2[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < cs * 2; j += blockDim.x) {
    double s = 0.0;
    for (int l = 0; l < lanes; ++l) s += (double)s_red[l * cs * 2 + j];
    const int ch = c0 + (j >> 1);
    if (BWD && (j & 1)) {
      double sd = 0.0;
      for (int l = 0; l < lanes; ++l) sd += (double)s_red[l * cs * 2 + j - 1];
      s = (double)__ldg(invstd + v * c + ch) * (s - (double)__ldg(mean + v * c + ch) * sd);
    }
    atomicAdd(acc + ((long long)v * c + ch) * 2 + (j & 1), s);
  }
  if (fin.ticket == nullptr) return;
  // ---- the last block to finish turns the sums into the per-(view, channel) coefficients ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(fin.ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x * gridDim.y * gridDim.z - 1) return;
  __threadfence();
  if (BWD) bn_bwd_finalize_block(acc, fin, mean, invstd, c, views, threadIdx.x, blockDim.x);
  else bn_finalize_block(acc, fin, c, views, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) {
    *fin.ticket = 0;
    if (!BWD && fin.nbt != nullptr) *fin.nbt += views;
  }
}

// stand-alone finalize launches (one block)
__global__ void __launch_bounds__(256) bn_finalize_kernel(double* __restrict__ acc, const BnFinalize fin, int c, int views) {
  if (threadIdx.x == 0 && fin.nbt != nullptr) *fin.nbt += views;
  bn_finalize_block(acc, fin, c, views, threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(double* __restrict__ acc, const BnFinalize fin,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int c, int views) {
  bn_bwd_finalize_block(acc, fin, mean, invstd, c, views, threadIdx.x, blockDim.x);
}

// y = relu?(a[v,c] * z + b[v,c] + residual)
// grid = (x, n_img): every thread owns ONE 8-channel group (its coefficients live in registers) and
// walks the image's pixels with 4 independent 16-byte loads in flight per tensor.
constexpr int kEwUnroll = 4;

// UNROLL = kEwUnroll with a residual (two input streams); without one (most launches: bn1/bn2 of
// every block, the stem, the downsample branches) twice that, so that a thread keeps the same
// 128 bytes of loads in flight either way.
template <typename T, int UNROLL, bool HAS_RES>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2)
bn_apply_kernel(const T* __restrict__ z, const float* __restrict__ a, const float* __restrict__ b,
                const T* __restrict__ residual, T* __restrict__ y,
                uint8_t* __restrict__ relu_bits, int pix, int c, int views, int relu) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const int cg = c / 8;
  const int n = blockIdx.y, v = n % views;
  const long long per_img = (long long)pix * cg;  // 8-channel vectors in this image
  const long long stride = (long long)gridDim.x * blockDim.x;  // multiple of cg (cg | 256)
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = (int)(i0 % cg);
  float ca[8], cb[8];
  V8<float>::load(a + v * c + g * 8, ca);
  V8<float>::load(b + v * c + g * 8, cb);
  const T* zi = z + (long long)n * pix * c;
  const T* ri = HAS_RES ? residual + (long long)n * pix * c : nullptr;
  T* yi = y + (long long)n * pix * c;
  uint8_t* bi = relu_bits ? relu_bits + (long long)n * per_img : nullptr;
  for (long long i = i0; i < per_img; i += stride * UNROLL) {
    typename V8<T>::Raw zr[UNROLL], rr[HAS_RES ? UNROLL : 1];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const long long k = i + j * stride;
      if (k < per_img) {
        zr[j] = V8<T>::load_raw(zi + k * 8);
        if (HAS_RES) rr[j] = V8<T>::load_raw(ri + k * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const long long k = i + j * stride;
      if (k < per_img) {
        float f[8], r[8], o[8];
        V8<T>::unpack(zr[j], f);
        if (HAS_RES) V8<T>::unpack(rr[j], r);
        uint32_t bits = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = fmaf(f[e], ca[e], cb[e]);
          if (HAS_RES) o[e] += r[e];
          bits |= (o[e] > 0.f ? 1u : 0u) << e;
          if (relu) o[e] = fmaxf(o[e], 0.f);
        }
        V8<T>::store(yi + k * 8, o);
        if (bi) bi[k] = (uint8_t)bits;
      }
    }
  }
}

// dz = k0*dyr + k1*z + k2, dyr = dy * (y_mask > 0); optionally also writes dyr (skip-path grad)
constexpr int kBwdUnroll = 4;  // 3 input streams x 4 packed 16-byte loads in flight per thread

template <typename T, int MASK>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ z, const T* __restrict__ dy, const void* __restrict__ y_mask_v,
                    const float* __restrict__ k0, const float* __restrict__ k1,
                    const float* __restrict__ k2, T* __restrict__ dz, T* __restrict__ dyr_out,
                    int pix, int c, int views) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const T* y_mask = reinterpret_cast<const T*>(y_mask_v);
  const uint8_t* y_bits = reinterpret_cast<const uint8_t*>(y_mask_v);
  const int cg = c / 8;
  const int n = blockIdx.y, v = n % views;
  const long long per_img = (long long)pix * cg;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = (int)(i0 % cg);
  float c0[8], c1[8], c2[8];
  V8<float>::load(k0 + v * c + g * 8, c0);
  V8<float>::load(k1 + v * c + g * 8, c1);
  V8<float>::load(k2 + v * c + g * 8, c2);
  const long long img = (long long)n * pix * c;
  for (long long i = i0; i < per_img; i += stride * kBwdUnroll) {
    typename V8<T>::Raw zr[kBwdUnroll], dr[kBwdUnroll], mr[kBwdUnroll];
    uint32_t mb[kBwdUnroll];
#pragma unroll
    for (int j = 0; j < kBwdUnroll; ++j) {
      const long long k = i + j * stride;
      if (k < per_img) {
        zr[j] = V8<T>::load_raw(z + img + k * 8);
        dr[j] = V8<T>::load_raw(dy + img + k * 8);
        if (MASK == 1) mr[j] = V8<T>::load_raw(y_mask + img + k * 8);
        if (MASK == 2) mb[j] = y_bits[(img >> 3) + k];
      }
    }
#pragma unroll
    for (int j = 0; j < kBwdUnroll; ++j) {
      const long long k = i + j * stride;
      if (k < per_img) {
        float f[8], d[8], m[8], o[8];
        V8<T>::unpack(zr[j], f);
        V8<T>::unpack(dr[j], d);
        if (MASK == 1) {
          V8<T>::unpack(mr[j], m);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[e] = m[e] > 0.f ? d[e] : 0.f;
        }
        if (MASK == 2) apply_bits(d, mb[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(c0[e], d[e], fmaf(c1[e], f[e], c2[e]));
        V8<T>::store(dz + img + k * 8, o);
        if (dyr_out) V8<T>::store(dyr_out + img + k * 8, d);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Elementwise helpers
// ---------------------------------------------------------------------------------------------
// dst = (mask > 0 ? src : 0) [+ add]; 2-D with row strides (columns contiguous, multiple of 8)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ src, long long ld_src,
                                const T* __restrict__ mask, long long ld_mask,
                                const T* __restrict__ add, long long ld_add, T* __restrict__ dst,
                                long long ld_dst, int cols, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = cols / 8;
  const int g = (int)(idx % cg);
  const long long row = idx / cg;
  float s[8];
  V8<T>::load(src + row * ld_src + g * 8, s);
  if (mask != nullptr) {
    float m[8];
    V8<T>::load(mask + row * ld_mask + g * 8, m);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = m[i] > 0.f ? s[i] : 0.f;
  }
  if (add != nullptr) {
    float a[8];
    V8<T>::load(add + row * ld_add + g * 8, a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += a[i];
  }
  V8<T>::store(dst + row * ld_dst + g * 8, s);
}

// dst[i] = bit i of `bits` ? src[i] : 0 over a flat tensor (8 elements / one mask byte per thread):
// multiplies a gradient by the ReLU derivative recorded as a packed sign mask.
template <typename T>
__global__ void mask_bits_kernel(const T* __restrict__ src, const uint8_t* __restrict__ bits,
                                 T* __restrict__ dst, long long n8) {
  griddep_wait();
  griddep_launch();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    V8<T>::load(src + i * 8, f);
    apply_bits(f, bits[i]);
    V8<T>::store(dst + i * 8, f);
  }
}

// out[col] (+)= sum_rows x[row, col]  (fp32 output; bias gradients)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  __shared__ float s[8][32 + 1];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int lane_r = threadIdx.x >> 5;  // 8 row lanes
  const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float acc = 0.f;
  if (col < cols)
    for (int r = r0 + lane_r; r < r1; r += 8) {
      if constexpr (sizeof(T) == 2) acc += __bfloat162float(x[(long long)r * ld + col]);
      else acc += x[(long long)r * ld + col];
    }
  s[lane_r][threadIdx.x & 31] = acc;
  __syncthreads();
  if (lane_r == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x & 31];
    atomicAdd(out + col, t);
  }
}

// Generic 4-D permute + cast (+ flips on output dims 1,2): weight layout transforms
//   dst[i0,i1,i2,i3] (contiguous) = src[i0*s0 + f(i1)*s1 + f(i2)*s2 + i3*s3]
template <typename TD>
__global__ void permute_cast_kernel(const float* __restrict__ src, TD* __restrict__ dst, int d0,
                                    int d1, int d2, int d3, long long s0, long long s1,
                                    long long s2, long long s3, int flip1, int flip2,
                                    long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i3 = (int)(idx % d3);
  long long t = idx / d3;
  int i2 = (int)(t % d2); t /= d2;
  int i1 = (int)(t % d1);
  const int i0 = (int)(t / d1);
  if (flip1) i1 = d1 - 1 - i1;
  if (flip2) i2 = d2 - 1 - i2;
  const float v = __ldg(src + i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3);
  if constexpr (sizeof(TD) == 2) dst[idx] = __float2bfloat16_rn(v);
  else dst[idx] = v;
}

// Batched variant: one launch re-derives EVERY engine-layout filter tensor of a training step from
// the fp32 masters (job table in device memory; block b belongs to the last job with
// first_block <= b; 1024 elements per block).
template <typename TD>
__device__ __forceinline__ void permute_job_elems(const rmv_permute_job& j, long long base, int tid) {
  const long long total = (long long)j.d0 * j.d1 * j.d2 * j.d3;
  TD* dst = reinterpret_cast<TD*>(j.dst);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long idx = base + tid + u * 256;
    if (idx >= total) return;
    const int i3 = (int)(idx % j.d3);
    long long t = idx / j.d3;
    int i2 = (int)(t % j.d2); t /= j.d2;
    int i1 = (int)(t % j.d1);
    const int i0 = (int)(t / j.d1);
    if (j.flip1) i1 = j.d1 - 1 - i1;
    if (j.flip2) i2 = j.d2 - 1 - i2;
    const float v = __ldg(j.src + i0 * j.s0 + i1 * j.s1 + i2 * j.s2 + i3 * j.s3);
    if constexpr (sizeof(TD) == 2) dst[idx] = __float2bfloat16_rn(v);
    else dst[idx] = v;
  }
}

// kind 1: dst = cast(src), both contiguous; 1024 elements per block.
template <typename TD>
__device__ __forceinline__ void permute_job_cast(const rmv_permute_job& j, long long blk, int tid) {
  const long long total = (long long)j.d0 * j.d1 * j.d2 * j.d3;
  TD* dst = reinterpret_cast<TD*>(j.dst);
  const long long i = blk * 1024 + tid * 4;
  if (i + 3 < total && ((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(j.src + i);
    if constexpr (sizeof(TD) == 2) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
      *reinterpret_cast<float4*>(dst + i) = v;
    }
  } else {
    for (long long e = i; e < total && e < i + 4; ++e) {
      if constexpr (sizeof(TD) == 2) dst[e] = __float2bfloat16_rn(j.src[e]);
      else dst[e] = j.src[e];
    }
  }
}

// kind 2: src [K][J] (J = C*T contiguous) -> dst [J'][K], J' = c*T + (flip ? T-1-t : t): the
// transposed (and tap-flipped) filters of the data-gradient convolutions / Linear dX GEMMs.
// 64 x 64 tiles through shared memory, coalesced on both sides. (d0,d1,d2,d3) = (C,R,S,K).
template <typename TD>
__device__ __forceinline__ void permute_job_transpose(const rmv_permute_job& j, long long blk,
                                                      int tid, float (*tile)[65]) {
  const int K = j.d3, T = j.d1 * j.d2, J = j.d0 * T;
  const int tiles_j = (J + 63) / 64;
  const int j0 = (int)(blk % tiles_j) * 64, k0 = (int)(blk / tiles_j) * 64;
  TD* dst = reinterpret_cast<TD*>(j.dst);
  const int tx = tid & 63, ty = tid >> 6;  // 64 x 4
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int k = k0 + r, jj = j0 + tx;
    tile[r][tx] = (k < K && jj < J) ? __ldg(j.src + (long long)k * J + jj) : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int jj = j0 + r, k = k0 + tx;
    if (jj < J && k < K) {
      int c = jj / T, t = jj - c * T;
      if (j.flip1) t = T - 1 - t;
      const float v = tile[tx][r];
      const long long o = ((long long)c * T + t) * K + k;
      if constexpr (sizeof(TD) == 2) dst[o] = __float2bfloat16_rn(v);
      else dst[o] = v;
    }
  }
}

// kind 3: src [K][C][T] -> dst [K][T][C] (forward KRSC filters of the 3x3 convs): one block per k.
// (d0,d1,d2,d3) = (K,R,S,C)
template <typename TD>
__device__ __forceinline__ void permute_job_krsc(const rmv_permute_job& j, long long blk, int tid) {
  const int T = j.d1 * j.d2, C = j.d3;
  const long long per_k = (long long)C * T;
  const float* src = j.src + blk * per_k;
  TD* dst = reinterpret_cast<TD*>(j.dst) + blk * per_k;
  // reads walk dst order (t, c): consecutive threads read stride-T floats (T odd: every 32-byte
  // sector is fully used by the block within a few iterations, L1 keeps it); writes are coalesced
  for (int o = tid; o < per_k; o += 256) {
    const int t = o / C, c = o - t * C;
    const float v = __ldg(src + (long long)c * T + t);
    if constexpr (sizeof(TD) == 2) dst[o] = __float2bfloat16_rn(v);
    else dst[o] = v;
  }
}

__global__ void __launch_bounds__(256)
permute_cast_batch_kernel(const rmv_permute_job* __restrict__ jobs, int n_jobs) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  __shared__ int s_job;
  __shared__ float s_tile[64][65];
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_job = lo;
  }
  __syncthreads();
  const rmv_permute_job j = jobs[s_job];
  const long long blk = (long long)(blockIdx.x - j.first_block);
  const bool bf = j.dst_dtype == RMV_DTYPE_BF16;
  switch (j.kind) {
    case 1:
      if (bf) permute_job_cast<__nv_bfloat16>(j, blk, threadIdx.x);
      else permute_job_cast<float>(j, blk, threadIdx.x);
      break;
    case 2:
      if (bf) permute_job_transpose<__nv_bfloat16>(j, blk, threadIdx.x, s_tile);
      else permute_job_transpose<float>(j, blk, threadIdx.x, s_tile);
      break;
    case 3:
      if (bf) permute_job_krsc<__nv_bfloat16>(j, blk, threadIdx.x);
      else permute_job_krsc<float>(j, blk, threadIdx.x);
      break;
    default:
      if (bf) permute_job_elems<__nv_bfloat16>(j, blk * 1024, threadIdx.x);
      else permute_job_elems<float>(j, blk * 1024, threadIdx.x);
  }
}

// dst[n, 2h, 2w, :] = src[n, h, w, :], zeros elsewhere (dst is [n, 2H, 2W, c]) -- stride-2 dgrad
template <typename T>
__global__ void dilate2_kernel(const T* __restrict__ src, T* __restrict__ dst, int h, int w, int c,
                               long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over dst / 8
  if (idx >= total) return;
  const int cg = c / 8;
  const int g = (int)(idx % cg);
  long long t = idx / cg;
  const int ow = (int)(t % (2 * w)); t /= (2 * w);
  const int oh = (int)(t % (2 * h));
  const long long n = t / (2 * h);
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (((ow | oh) & 1) == 0) V8<T>::load(src + ((n * h + oh / 2) * w + ow / 2) * c + g * 8, f);
  V8<T>::store(dst + idx * 8, f);
}

// MaxPool 3x3 s2 p1 backward (NHWC): gather form -- each input pixel sums the gradients of the
// (<= 4) windows whose maximum it is (first maximum in scan order wins, as in ATen).
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                   T* __restrict__ dx, int in_h, int in_w, int c, int out_h,
                                   int out_w, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c / 8;
  const int g = (int)(idx % cg);
  long long t = idx / cg;
  const int iw = (int)(t % in_w); t /= in_w;
  const int ih = (int)(t % in_h);
  const long long n = t / in_h;
  float xv[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  V8<T>::load(x + idx * 8, xv);
  for (int oh = max(0, (ih - 1 + 1) / 2); oh <= min(out_h - 1, (ih + 1) / 2); ++oh)
    for (int ow = max(0, (iw - 1 + 1) / 2); ow <= min(out_w - 1, (iw + 1) / 2); ++ow) {
      // window of (oh, ow): rows 2oh-1..2oh+1, cols 2ow-1..2ow+1; find its first argmax per channel
      float best[8];
      int arg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { best[i] = -3.4e38f; arg[i] = -1; }
      for (int r = 0; r < 3; ++r) {
        const int hh = 2 * oh - 1 + r;
        if (hh < 0 || hh >= in_h) continue;
        for (int s = 0; s < 3; ++s) {
          const int ww = 2 * ow - 1 + s;
          if (ww < 0 || ww >= in_w) continue;
          float f[8];
          V8<T>::load(x + ((n * in_h + hh) * in_w + ww) * c + g * 8, f);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (f[i] > best[i]) { best[i] = f[i]; arg[i] = hh * in_w + ww; }
        }
      }
      float d[8];
      V8<T>::load(dy + ((n * out_h + oh) * out_w + ow) * c + g * 8, d);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (arg[i] == ih * in_w + iw) acc[i] += d[i];
    }
  V8<T>::store(dx + idx * 8, acc);
}

// MaxPool 3x3 s2 p1 forward that also records, per output element, WHICH of the 9 window positions
// won (first maximum in scan order, as ATen): idx in 0..8 = r*3+s. The backward then needs 4 index
// bytes + 4 gradient values per input element instead of re-scanning 4 windows x 9 inputs.
template <typename T>
__global__ void maxpool_fwd_idx_kernel(const T* __restrict__ x, T* __restrict__ y,
                                       uint8_t* __restrict__ idx, int in_h, int in_w, int c,
                                       int out_h, int out_w, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = c / 8;
  const int g = (int)(i % cg);
  long long t = i / cg;
  const int ow = (int)(t % out_w); t /= out_w;
  const int oh = (int)(t % out_h);
  const long long n = t / out_h;
  float best[8];
  uint32_t arg[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -3.4e38f; arg[e] = 0; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = 2 * oh - 1 + r;
    if (ih < 0 || ih >= in_h) continue;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int iw = 2 * ow - 1 + q;
      if (iw < 0 || iw >= in_w) continue;
      float f[8];
      V8<T>::load(x + ((n * in_h + ih) * in_w + iw) * c + g * 8, f);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (f[e] > best[e]) { best[e] = f[e]; arg[e] = r * 3 + q; }
    }
  }
  V8<T>::store(y + i * 8, best);
  uint2 packed;
  packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
  *reinterpret_cast<uint2*>(idx + i * 8) = packed;
}

// Thread = a 2x2 block of input pixels (8 channels): rows 2k, 2k+1 and columns 2j, 2j+1 are
// covered by exactly the four windows (k..k+1, j..j+1) -- the even row/column only by window k/j
// (its centre), the odd one by k (bottom/right edge) and k+1 (top/left edge). Four (index, dy)
// loads give four outputs; every (pixel, window) pair has a compile-time index code.
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const uint8_t* __restrict__ idx, const T* __restrict__ dy,
                       T* __restrict__ dx, int in_h, int in_w, int c, int out_h, int out_w,
                       int blk_h, int blk_w, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = c / 8;
  const int g = (int)(i % cg);
  long long t = i / cg;
  const int j = (int)(t % blk_w); t /= blk_w;
  const int k = (int)(t % blk_h);
  const long long n = t / blk_h;
  uint2 code[2][2];
  float d[2][2][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int oh = k + a, ow = j + b;
      if (oh < out_h && ow < out_w) {
        const long long o = ((n * out_h + oh) * out_w + ow) * c + g * 8;
        code[a][b] = *reinterpret_cast<const uint2*>(idx + o);
        V8<T>::load(dy + o, d[a][b]);
      } else {
        code[a][b] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);  // matches no position
#pragma unroll
        for (int e = 0; e < 8; ++e) d[a][b][e] = 0.f;
      }
    }
  // input pixel (2k+p, 2j+q): windows a <= p, b <= q; position inside window (k+a, j+b) is
  // (p - 2a + 1, q - 2b + 1), stored as r*3 + s
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ih = 2 * k + p, iw = 2 * j + q;
      if (ih >= in_h || iw >= in_w) continue;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a <= p; ++a)
#pragma unroll
        for (int b = 0; b <= q; ++b) {
          const uint32_t want = (uint32_t)((p - 2 * a + 1) * 3 + (q - 2 * b + 1));
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t got = ((e < 4 ? code[a][b].x : code[a][b].y) >> (8 * (e & 3))) & 0xffu;
            if (got == want) acc[e] += d[a][b][e];
          }
        }
      V8<T>::store(dx + (((n * in_h + ih) * in_w + iw) * c + g * 8), acc);
    }
}

// dx[n, p, :] = dfeat[n, :] / hw
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dfeat, long long ld, T* __restrict__ dx,
                                   int hw, int c, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = c / 8;
  const int g = (int)(idx % cg);
  const long long n = idx / cg / hw;
  float f[8];
  V8<T>::load(dfeat + n * ld + g * 8, f);
  const float inv = 1.f / (float)hw;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] *= inv;
  V8<T>::store(dx + idx * 8, f);
}

// ---------------------------------------------------------------------------------------------
// Angular-loss gradient + gaze-head tail backward
//   dpred[m,:] = loss_scale * d_m * d(angular_deg(pred_m, gt_m))/dpred   (0 where the cosine clamps)
//   dhidden[m,:] = (hidden > 0) * (dpred[m,0] w2[0,:] + dpred[m,1] w2[1,:])
//   dw2 += dpred^T hidden,  db2 += colsum(dpred)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void loss_grad_row(float p, float y, float gp, float gy, float w,
                                              float* dp, float* dyw) {
  float sp, cp, sy, cy, sgp, cgp, sgy, cgy;
  sincosf(p, &sp, &cp); sincosf(y, &sy, &cy);
  sincosf(gp, &sgp, &cgp); sincosf(gy, &sgy, &cgy);
  const float b[3] = {cp * sy, sp, cp * cy};
  const float a[3] = {cgp * sgy, sgp, cgp * cgy};
  const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-6f);
  const float nb = fmaxf(sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), 1e-6f);
  const float an[3] = {a[0] / na, a[1] / na, a[2] / na};
  const float bn[3] = {b[0] / nb, b[1] / nb, b[2] / nb};
  const float sim = an[0] * bn[0] + an[1] * bn[1] + an[2] * bn[2];
  if (!(sim > -1.f && sim < 1.f)) { *dp = 0.f; *dyw = 0.f; return; }  // hardtanh saturated
  const float dth = -kRadToDeg * w * rsqrtf(1.f - sim * sim);  // d(deg)/d(sim) * weight
  // d sim / d b = (an - sim * bn) / nb
  const float g[3] = {(an[0] - sim * bn[0]) / nb, (an[1] - sim * bn[1]) / nb,
                      (an[2] - sim * bn[2]) / nb};
  *dp = dth * (g[0] * (-sp * sy) + g[1] * cp + g[2] * (-sp * cy));
  *dyw = dth * (g[0] * (cp * cy) + g[2] * (-cp * sy));
}

// One warp finishes kHeadRows rows per iteration (grid-stride): their loads are issued first, lanes
// 0..kHeadRows-1 each compute the loss gradient of one row meanwhile and broadcast it by shuffle.
// NCH = hid / 256 (1 or 2; the reference head has hid = 512): each lane owns the same 8*NCH columns
// for every row it sees, so the dw2 partial sums live in registers; the warps of a block meet in
// shared memory once and the block issues one set of global atomics whatever the row count.
// NCH = 0: generic hid, shared-memory accumulation per row.
constexpr int kHeadRows = 4;

template <typename T, int NCH>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 2 : 1)
head_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                     const T* __restrict__ hidden, long long ld_h, const float* __restrict__ w2,
                     int rows, int hid, float loss_scale, int views, float aux_decay,
                     T* __restrict__ dhidden, long long ld_dh, float* __restrict__ dpred_out,
                     float* __restrict__ dw2, float* __restrict__ db2) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  extern __shared__ float s_dw[];  // [2][hid] + [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * hid + 2; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  constexpr int R = NCH > 0 ? NCH : 1;
  float ga[R][8], gb[R][8];
  float g_dp = 0.f, g_dy = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) { ga[j][i] = 0.f; gb[j][i] = 0.f; }
  for (int row0 = (blockIdx.x * 8 + warp) * kHeadRows; row0 < rows;
       row0 += gridDim.x * 8 * kHeadRows) {
    typename V8<T>::Raw raw[kHeadRows][R];
    if (NCH > 0) {
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        const int row = min(row0 + r, rows - 1);   // past the end: re-read the last row, dropped below
#pragma unroll
        for (int j = 0; j < R; ++j)
          raw[r][j] = V8<T>::load_raw(hidden + (long long)row * ld_h + j * 256 + lane * 8);
      }
    }
    // lane r: loss gradient of row row0 + r
    float my_dp = 0.f, my_dy = 0.f;
    {
      const int row = row0 + lane;
      if (lane < kHeadRows && row < rows) {
        if (gt != nullptr) {
          const float wgt = loss_scale * ((row % views) == 0 ? 1.f : aux_decay);
          const float2 p2 = __ldg(reinterpret_cast<const float2*>(pred + (long long)row * 2));
          const float2 g2 = __ldg(reinterpret_cast<const float2*>(gt + (long long)row * 2));
          loss_grad_row(p2.x, p2.y, g2.x, g2.y, wgt, &my_dp, &my_dy);
          *reinterpret_cast<float2*>(dpred_out + (long long)row * 2) = make_float2(my_dp, my_dy);
        } else {
          // external d(loss)/d(pred): the caller's own loss produced it (autograd bridge)
          const float2 d2 = *reinterpret_cast<const float2*>(dpred_out + (long long)row * 2);
          my_dp = d2.x; my_dy = d2.y;
        }
        g_dp += my_dp; g_dy += my_dy;
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const float dp = __shfl_sync(0xffffffffu, my_dp, r), dyw = __shfl_sync(0xffffffffu, my_dy, r);
      const int row = row0 + r;
      if (row >= rows) break;   // warp-uniform
      if (NCH > 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
          float h[8], a[8], b[8], o[8];
          V8<T>::unpack(raw[r][j], h);
          V8<float>::load(w2 + j * 256 + lane * 8, a);        // 4 KB table, L1-resident
          V8<float>::load(w2 + hid + j * 256 + lane * 8, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] = h[i] > 0.f ? fmaf(dp, a[i], dyw * b[i]) : 0.f;
            ga[j][i] = fmaf(dp, h[i], ga[j][i]);
            gb[j][i] = fmaf(dyw, h[i], gb[j][i]);
          }
          V8<T>::store(dhidden + (long long)row * ld_dh + j * 256 + lane * 8, o);
        }
      } else {
        for (int k = lane * 8; k < hid; k += 256) {
          float h[8], a[8], b[8], o[8];
          V8<T>::load(hidden + (long long)row * ld_h + k, h);
          V8<float>::load(w2 + k, a);
          V8<float>::load(w2 + hid + k, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] = h[i] > 0.f ? fmaf(dp, a[i], dyw * b[i]) : 0.f;
            atomicAdd(&s_dw[k + i], dp * h[i]);
            atomicAdd(&s_dw[hid + k + i], dyw * h[i]);
          }
          V8<T>::store(dhidden + (long long)row * ld_dh + k, o);
        }
      }
    }
  }
  if (NCH > 0) {
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_dw[j * 256 + lane * 8 + i], ga[j][i]);
        atomicAdd(&s_dw[hid + j * 256 + lane * 8 + i], gb[j][i]);
      }
  }
  if (lane < kHeadRows) { atomicAdd(&s_dw[2 * hid], g_dp); atomicAdd(&s_dw[2 * hid + 1], g_dy); }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * hid; i += blockDim.x) atomicAdd(dw2 + i, s_dw[i]);
  if (threadIdx.x < 2) atomicAdd(db2 + threadIdx.x, s_dw[2 * hid + threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// FFMA weight gradient: dW[k, c, r, s] (fp32, PyTorch OIHW layout, +=) =
//     sum_p dY[p, k] * X[p shifted by (r,s), c]           (split over pixels, fp32 atomics)
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const void* x; const void* dy; float* dw;
  long long x_sn, x_sh, x_sw, x_sc;
  long long dy_sn, dy_sh, dy_sw;
  int n_img, in_h, in_w, c_in, c_out, kh, kw, stride, pad, out_h, out_w;
  long long p_total;
  int k_cols;  // kh*kw*c_in
  int p_per_split;
};

template <typename T>
__global__ void __launch_bounds__(256) simt_wgrad_kernel(const WgradArgs a) {
  constexpr int TM = 64, TN = 64, TP = 16;
  __shared__ float As[TP][TM + 4];  // dY tile  [pixel][k_out]
  __shared__ float Bs[TP][TN + 4];  // X tile   [pixel][(r,s,c) column]
  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ dy = reinterpret_cast<const T*>(a.dy);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM;   // k_out tile
  const int n0 = blockIdx.x * TN;   // (r,s,c) column tile
  const long long p_begin = (long long)blockIdx.z * a.p_per_split;
  const long long p_end = min(a.p_total, p_begin + a.p_per_split);
  // loader assignment: thread -> (pixel lane 0..15, 4 consecutive columns)
  const int l_p = tid >> 4, l_c0 = (tid & 15) * 4;
  // decode this thread's 4 B columns once
  int b_r[4], b_s[4], b_c[4];
  bool b_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int col = n0 + l_c0 + i;
    b_ok[i] = col < a.k_cols;
    const int tap = b_ok[i] ? col / a.c_in : 0;
    b_c[i] = b_ok[i] ? col - tap * a.c_in : 0;
    b_r[i] = tap / a.kw;
    b_s[i] = tap - b_r[i] * a.kw;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long p0 = p_begin; p0 < p_end; p0 += TP) {
    const long long p = p0 + l_p;
    const bool p_ok = p < p_end;
    int ow = 0, oh = 0, n = 0;
    if (p_ok) {
      ow = (int)(p % a.out_w);
      const long long t = p / a.out_w;
      oh = (int)(t % a.out_h);
      n = (int)(t / a.out_h);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float av = 0.f, bv = 0.f;
      const int k = m0 + l_c0 + i;
      if (p_ok && k < a.c_out) {
        const T* q = dy + n * a.dy_sn + oh * a.dy_sh + ow * a.dy_sw + k;
        if constexpr (sizeof(T) == 2) av = __bfloat162float(*q); else av = *q;
      }
      if (p_ok && b_ok[i]) {
        const int ih = oh * a.stride - a.pad + b_r[i], iw = ow * a.stride - a.pad + b_s[i];
        if (ih >= 0 && ih < a.in_h && iw >= 0 && iw < a.in_w) {
          const T* q = x + n * a.x_sn + ih * a.x_sh + iw * a.x_sw + b_c[i] * a.x_sc;
          if constexpr (sizeof(T) == 2) bv = __bfloat162float(*q); else bv = *q;
        }
      }
      As[l_p][l_c0 + i] = av;
      Bs[l_p][l_c0 + i] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TP; ++q) {
      const float4 av = *reinterpret_cast<const float4*>(&As[q][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[q][tx * 4]);
      const float am[4] = {av.x, av.y, av.z, av.w};
      const float bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = m0 + ty * 4 + i;
    if (k >= a.c_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= a.k_cols) continue;
      const int tap = col / a.c_in, c = col - tap * a.c_in;
      const int r = tap / a.kw, s = tap - r * a.kw;
      atomicAdd(a.dw + (((long long)k * a.c_in + c) * a.kh + r) * a.kw + s, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused Adam over one flat fp32 buffer. hyper = double{lr, beta1, beta2, eps, weight_decay, step} lives in
// DEVICE memory (graph-capturable; the step counter is advanced by adam_tick_kernel).
//   decoupled == 0: torch.optim.Adam(weight_decay): g += wd * p   (trainer.py:54)
//   decoupled == 1: AdamW: p *= 1 - lr * wd
// ---------------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(double* hyper) { hyper[5] += 1.0; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v,
                            const double* __restrict__ hyper, int decoupled, float grad_scale,
                            long long n) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  // scalar prefactors in fp64, exactly as torch.optim.Adam computes them on the host
  const double lr_d = hyper[0], b1_d = hyper[1], b2_d = hyper[2], step_d = hyper[5];
  const float lr = (float)lr_d, b1 = (float)b1_d, b2 = (float)b2_d;
  const float eps = (float)hyper[3], wd = (float)hyper[4];
  const float bc2_sqrt = (float)sqrt(1.0 - pow(b2_d, step_d));
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, step_d)));
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      float4 pp = *reinterpret_cast<float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gr = G[j] * grad_scale;
        if (decoupled) P[j] *= 1.f - lr * wd; else gr = fmaf(wd, P[j], gr);
        M[j] = fmaf(b1, M[j], (1.f - b1) * gr);       // m.lerp_(g, 1-b1) up to rounding
        V[j] = fmaf(b2, V[j], (1.f - b2) * gr * gr);
        const float denom = sqrtf(V[j]) / bc2_sqrt + eps;
        P[j] -= step_size * (M[j] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long k = i; k < n; ++k) {
        float gr = g[k] * grad_scale;
        float pk = p[k];
        if (decoupled) pk *= 1.f - lr * wd; else gr = fmaf(wd, pk, gr);
        const float mk = fmaf(b1, m[k], (1.f - b1) * gr);
        const float vk = fmaf(b2, v[k], (1.f - b2) * gr * gr);
        m[k] = mk; v[k] = vk;
        p[k] = pk - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
      }
    }
  }
}

}  // namespace
}  // namespace rmv

using namespace rmv;

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == RMV_DTYPE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else { using T = float; __VA_ARGS__; }

// RMV_BN_APPLY_U8=0 keeps the 4-deep form for launches without a residual (A/B switch)
static bool apply_unroll8() {
  static const int on = [] { const char* e = getenv("RMV_BN_APPLY_U8"); return e ? atoi(e) : 1; }();
  return on != 0;
}

// blocks along x for the per-image elementwise kernels: enough threads for kEwUnroll vectors each,
// but at least ~4 waves of blocks over the whole launch
static unsigned ew_blocks_x(int pix, int c, int n_img, int unroll = rmv::kEwUnroll) {
  const long long per_img = (long long)pix * (c / 8);
  long long bx = (per_img + 256LL * unroll - 1) / (256LL * unroll);
  if (bx < 1) bx = 1;
  return (unsigned)bx;
}

// grid (G, views, channel slices): one wave of co-resident blocks (occ per SM), but no more row
// slabs than there are 16-row (per thread) pieces; channels in slices of at most 256.
static int bn_reduce_cfg(int pix, int c, int n_img, int views, int occ, dim3* grid, int* smem) {
  RMV_CHECK_ARG(c % 8 == 0 && (c / 8) <= 256 && 256 % (c / 8) == 0,
                "batchnorm: channels=%d must be 8*2^k with c <= 2048", c);
  const int cs = c > 256 ? 256 : c, slices = c / cs;
  const int cg = cs / 8;
  RMV_CHECK_ARG(pix > 32 || 256 / cg <= pix, "batchnorm: %d pixels per image is too few", pix);
  const int lanes = 256 / cg;
  const long long rows_total = (long long)(n_img / views) * pix;
  long long gx = (rows_total + lanes * 16 - 1) / (lanes * 16);
  long long want = (long long)occ * num_sms() / (views * slices);
  if (want < 1) want = 1;
  if (gx > want) gx = want;
  if (gx < 1) gx = 1;
  *grid = dim3((unsigned)gx, (unsigned)views, (unsigned)slices);
  *smem = lanes * cs * 2 * (int)sizeof(float);
  return 0;
}

static int bn_stats_launch(const void* z, int dtype, int n_img, int pix, int c, int views,
                           double* acc, const BnFinalize& fin, cudaStream_t stream) {
  RMV_CHECK_ARG(views >= 1 && n_img % views == 0, "bn_stats: n_img=%d not a multiple of views=%d", n_img, views);
  if (n_img == 0 || pix == 0) return 0;
  dim3 grid; int smem;
  if (int rc = bn_reduce_cfg(pix, c, n_img, views, 3, &grid, &smem)) return rc;
  DISPATCH_T(dtype, (rmv::launch_pdl(bn_reduce_kernel<T, false, 0>, dim3(grid), dim3(256), smem, stream, 
      (const T*)z, nullptr, nullptr, nullptr, nullptr, pix, c, views, n_img / views, acc, fin)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_bn_stats(const void* z, int dtype, int n_img, int pix, int c, int views,
                            double* acc, void* stream) {
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  return bn_stats_launch(z, dtype, n_img, pix, c, views, acc, fin, (cudaStream_t)stream);
}

static BnFinalize bn_fin_fwd(unsigned int* ticket, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long long* num_batches,
                             float* mean, float* invstd, float* a, float* b,
                             long long count_per_view, float eps, float momentum) {
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  fin.ticket = ticket; fin.gamma = gamma; fin.beta = beta;
  fin.running_mean = running_mean; fin.running_var = running_var; fin.nbt = num_batches;
  fin.mean = mean; fin.invstd = invstd; fin.a = a; fin.b = b;
  fin.count = (double)count_per_view; fin.eps = eps; fin.momentum = momentum;
  return fin;
}

extern "C" int rmv_bn_finalize(double* acc, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, long long* num_batches,
                               float* mean, float* invstd, float* a, float* b, int c, int views,
                               long long count_per_view, float eps, float momentum, void* stream) {
  RMV_CHECK_ARG(count_per_view > 1, "bn_finalize: need more than one value per channel");
  const BnFinalize fin = bn_fin_fwd(nullptr, gamma, beta, running_mean, running_var, num_batches,
                                    mean, invstd, a, b, count_per_view, eps, momentum);
  bn_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(acc, fin, c, views);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_bn_stats_finalize(const void* z, int dtype, int n_img, int pix, int c, int views,
                                     double* acc, unsigned int* ticket, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var,
                                     long long* num_batches, float* mean, float* invstd, float* a,
                                     float* b, float eps, float momentum, void* stream) {
  RMV_CHECK_ARG(ticket != nullptr, "bn_stats_finalize: null ticket");
  RMV_CHECK_ARG(views >= 1 && n_img % views == 0 && (long long)(n_img / views) * pix > 1,
                "bn_stats_finalize: need more than one value per (view, channel)");
  const BnFinalize fin = bn_fin_fwd(ticket, gamma, beta, running_mean, running_var, num_batches,
                                    mean, invstd, a, b, (long long)(n_img / views) * pix, eps,
                                    momentum);
  return bn_stats_launch(z, dtype, n_img, pix, c, views, acc, fin, (cudaStream_t)stream);
}

extern "C" int rmv_bn_apply(const void* z, const float* a, const float* b, const void* residual,
                            void* y, unsigned char* relu_bits, int dtype, int n_img, int pix, int c,
                            int views, int relu, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0 && 256 % (c / 8) == 0, "bn_apply: c=%d must be 8*2^k, <= 2048", c);
  if ((long long)n_img * pix == 0) return 0;
  if (residual == nullptr && dtype == RMV_DTYPE_BF16 && apply_unroll8()) {
    const dim3 grid(ew_blocks_x(pix, c, n_img, 2 * rmv::kEwUnroll), (unsigned)n_img);
    rmv::launch_pdl(bn_apply_kernel<__nv_bfloat16, 2 * rmv::kEwUnroll, false>, dim3(grid), dim3(256), 0,
                    (cudaStream_t)stream, (const __nv_bfloat16*)z, a, b, (const __nv_bfloat16*)nullptr,
                    (__nv_bfloat16*)y, relu_bits, pix, c, views, relu);
    RMV_LAUNCH_CHECK();
    return 0;
  }
  const dim3 grid(ew_blocks_x(pix, c, n_img), (unsigned)n_img);
  if (residual != nullptr) {
    DISPATCH_T(dtype, (rmv::launch_pdl(bn_apply_kernel<T, rmv::kEwUnroll, true>, dim3(grid), dim3(256), 0,
        (cudaStream_t)stream, (const T*)z, a, b, (const T*)residual, (T*)y, relu_bits, pix, c, views, relu)));
  } else {
    DISPATCH_T(dtype, (rmv::launch_pdl(bn_apply_kernel<T, rmv::kEwUnroll, false>, dim3(grid), dim3(256), 0,
        (cudaStream_t)stream, (const T*)z, a, b, (const T*)nullptr, (T*)y, relu_bits, pix, c, views, relu)));
  }
  RMV_LAUNCH_CHECK();
  return 0;
}

#define DISPATCH_MASK(y_mask, mask_is_bits, ...)                                  \
  if ((y_mask) == nullptr) { constexpr int MASK = 0; __VA_ARGS__; }              \
  else if (mask_is_bits) { constexpr int MASK = 2; __VA_ARGS__; }                \
  else { constexpr int MASK = 1; __VA_ARGS__; }

static int bn_bwd_reduce_launch(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                                const float* mean, const float* invstd, int dtype, int n_img,
                                int pix, int c, int views, double* acc, const BnFinalize& fin,
                                cudaStream_t stream) {
  RMV_CHECK_ARG(views >= 1 && n_img % views == 0, "bn_bwd_reduce: n_img not a multiple of views");
  if (n_img == 0 || pix == 0) return 0;
  dim3 grid; int smem;
  if (int rc = bn_reduce_cfg(pix, c, n_img, views, 3, &grid, &smem)) return rc;
  DISPATCH_T(dtype, DISPATCH_MASK(y_mask, mask_is_bits,
      (rmv::launch_pdl(bn_reduce_kernel<T, true, MASK>, dim3(grid), dim3(256), smem, stream, 
          (const T*)z, (const T*)dy, y_mask, mean, invstd, pix, c, views, n_img / views, acc, fin))));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_bn_bwd_reduce(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                                 const float* mean, const float* invstd, int dtype, int n_img,
                                 int pix, int c, int views, double* acc, void* stream) {
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  return bn_bwd_reduce_launch(z, dy, y_mask, mask_is_bits, mean, invstd, dtype, n_img, pix, c, views,
                              acc, fin, (cudaStream_t)stream);
}

static BnFinalize bn_fin_bwd(unsigned int* ticket, const float* gamma, float* dgamma, float* dbeta,
                             float* k0, float* k1, float* k2, long long count_per_view) {
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  fin.ticket = ticket; fin.gamma = gamma; fin.dgamma = dgamma; fin.dbeta = dbeta;
  fin.k0 = k0; fin.k1 = k1; fin.k2 = k2; fin.count = (double)count_per_view;
  return fin;
}

extern "C" int rmv_bn_bwd_finalize(double* acc, const float* gamma, const float* mean,
                                   const float* invstd, float* dgamma, float* dbeta, float* k0,
                                   float* k1, float* k2, int c, int views, long long count_per_view,
                                   void* stream) {
  const BnFinalize fin = bn_fin_bwd(nullptr, gamma, dgamma, dbeta, k0, k1, k2, count_per_view);
  bn_bwd_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(acc, fin, mean, invstd, c, views);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_bn_bwd_reduce_finalize(const void* z, const void* dy, const void* y_mask,
                                          int mask_is_bits, const float* mean, const float* invstd,
                                          int dtype, int n_img, int pix, int c, int views,
                                          double* acc, unsigned int* ticket, const float* gamma,
                                          float* dgamma, float* dbeta, float* k0, float* k1,
                                          float* k2, void* stream) {
  RMV_CHECK_ARG(ticket != nullptr, "bn_bwd_reduce_finalize: null ticket");
  RMV_CHECK_ARG(views >= 1 && n_img % views == 0, "bn_bwd_reduce_finalize: n_img not a multiple of views");
  const BnFinalize fin =
      bn_fin_bwd(ticket, gamma, dgamma, dbeta, k0, k1, k2, (long long)(n_img / views) * pix);
  return bn_bwd_reduce_launch(z, dy, y_mask, mask_is_bits, mean, invstd, dtype, n_img, pix, c, views,
                              acc, fin, (cudaStream_t)stream);
}

extern "C" int rmv_bn_bwd_apply(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                                const float* k0, const float* k1, const float* k2, void* dz,
                                void* dyr_out,
                                int dtype, int n_img, int pix, int c, int views, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0 && 256 % (c / 8) == 0, "bn_bwd_apply: c=%d must be 8*2^k, <= 2048", c);
  if ((long long)n_img * pix == 0) return 0;
  const dim3 grid(ew_blocks_x(pix, c, n_img, rmv::kBwdUnroll), (unsigned)n_img);
  DISPATCH_T(dtype, DISPATCH_MASK(y_mask, mask_is_bits,
      (rmv::launch_pdl(bn_bwd_apply_kernel<T, MASK>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
          (const T*)z, (const T*)dy, y_mask, k0, k1, k2, (T*)dz, (T*)dyr_out, pix, c, views))));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_relu_bwd(const void* src, long long ld_src, const void* mask, long long ld_mask,
                            const void* add, long long ld_add, void* dst, long long ld_dst,
                            int rows, int cols, int dtype, void* stream) {
  RMV_CHECK_ARG(cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && ld_mask % 8 == 0 && ld_add % 8 == 0,
                "relu_bwd: cols/ld must be multiples of 8");
  const long long total = (long long)rows * (cols / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (rmv::launch_pdl(relu_bwd_kernel<T>, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)src, ld_src, (const T*)mask, ld_mask, (const T*)add, ld_add, (T*)dst, ld_dst, cols,
      total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_colsum(const void* x, long long ld, int rows, int cols, int dtype, float* out,
                          void* stream) {
  if (rows == 0 || cols == 0) return 0;
  int ysplit = (rows + 255) / 256;
  if (ysplit > 64) ysplit = 64;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)ysplit);
  DISPATCH_T(dtype, (rmv::launch_pdl(colsum_kernel<T>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const T*)x, ld, rows, cols, out)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_permute_cast(const float* src, void* dst, int d0, int d1, int d2, int d3,
                                long long s0, long long s1, long long s2, long long s3, int flip1,
                                int flip2, int dst_dtype, void* stream) {
  const long long total = (long long)d0 * d1 * d2 * d3;
  if (total == 0) return 0;
  if (dst_dtype == RMV_DTYPE_BF16)
    permute_cast_kernel<__nv_bfloat16><<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, (__nv_bfloat16*)dst, d0, d1, d2, d3, s0, s1, s2, s3, flip1, flip2, total);
  else
    permute_cast_kernel<float><<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, (float*)dst, d0, d1, d2, d3, s0, s1, s2, s3, flip1, flip2, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_dilate2(const void* src, void* dst, int n_img, int h, int w, int c, int dtype,
                           void* stream) {
  RMV_CHECK_ARG(c % 8 == 0, "dilate2: c must be a multiple of 8");
  const long long total = (long long)n_img * 2 * h * 2 * w * (c / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (dilate2_kernel<T><<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const T*)src, (T*)dst, h, w, c, total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int n_img, int in_h,
                                    int in_w, int c, int dtype, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0, "maxpool_bwd: c must be a multiple of 8");
  const int out_h = (in_h - 1) / 2 + 1, out_w = (in_w - 1) / 2 + 1;
  const long long total = (long long)n_img * in_h * in_w * (c / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (maxpool_bwd_kernel<T><<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const T*)x, (const T*)dy, (T*)dx, in_h, in_w, c, out_h, out_w, total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_avgpool_bwd(const void* dfeat, long long ld, void* dx, int n_img, int hw, int c,
                               int dtype, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0 && ld % 8 == 0, "avgpool_bwd: c/ld must be multiples of 8");
  const long long total = (long long)n_img * hw * (c / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (rmv::launch_pdl(avgpool_bwd_kernel<T>, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)dfeat, ld, (T*)dx, hw, c, total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_mask_bits(const void* src, const void* bits, void* dst, long long n, int dtype,
                             void* stream) {
  RMV_CHECK_ARG(n % 8 == 0, "mask_bits: element count must be a multiple of 8");
  if (n == 0) return 0;
  RMV_CHECK_ARG(src && bits && dst, "mask_bits: null pointer");
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
  DISPATCH_T(dtype, (rmv::launch_pdl(mask_bits_kernel<T>, dim3((unsigned)blocks), dim3(256), 0,
                                     (cudaStream_t)stream, (const T*)src, (const uint8_t*)bits, (T*)dst,
                                     n / 8)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_head_loss_bwd(const float* pred, const float* gt, const void* hidden,
                                 long long ld_hidden, int hid_dtype, const float* w2, int rows,
                                 int hid, float loss_scale, int views, float aux_decay,
                                 void* dhidden, long long ld_dhidden, float* dpred, float* dw2,
                                 float* db2, void* stream) {
  RMV_CHECK_ARG(hid % 8 == 0 && ld_hidden % 8 == 0 && ld_dhidden % 8 == 0,
                "head_loss_bwd: hid/ld must be multiples of 8");
  RMV_CHECK_ARG(dpred != nullptr && (gt == nullptr || pred != nullptr),
                "head_loss_bwd: dpred (and pred when gt is given) must not be null");
  if (rows == 0) return 0;
  const int smem = (2 * hid + 2) * (int)sizeof(float);
  // grid-stride over groups of 8 warps x kHeadRows rows; one set of global atomics per block
  long blocks = ((long)rows + 8 * rmv::kHeadRows - 1) / (8 * rmv::kHeadRows);
  if (blocks > 2L * num_sms()) blocks = 2L * num_sms();
  const dim3 grid((unsigned)blocks);
#define RMV_HEAD_BWD(NCH)                                                                          \
  DISPATCH_T(hid_dtype, (rmv::launch_pdl(head_loss_bwd_kernel<T, NCH>, grid, dim3(256), smem,       \
      (cudaStream_t)stream, pred, gt, (const T*)hidden, ld_hidden, w2, rows, hid, loss_scale, views, \
      aux_decay, (T*)dhidden, ld_dhidden, dpred, dw2, db2)))
  if (hid == 512) { RMV_HEAD_BWD(2); }
  else if (hid == 256) { RMV_HEAD_BWD(1); }
  else { RMV_HEAD_BWD(0); }
#undef RMV_HEAD_BWD
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_conv2d_wgrad(const rmv_conv_args* args, const void* dy, float* dw, void* stream) {
  RMV_CHECK_ARG(args && dy && dw, "conv2d_wgrad: null pointer");
  const rmv_conv_args& p = *args;
  WgradArgs a;
  a.x = p.x; a.dy = dy; a.dw = dw;
  a.x_sn = p.x_sn; a.x_sh = p.x_sh; a.x_sw = p.x_sw; a.x_sc = p.x_sc ? p.x_sc : 1;
  a.dy_sn = p.y_sn; a.dy_sh = p.y_sh; a.dy_sw = p.y_sw;
  a.n_img = p.n_img; a.in_h = p.in_h; a.in_w = p.in_w; a.c_in = p.c_in; a.c_out = p.c_out;
  a.kh = p.kh; a.kw = p.kw; a.stride = p.stride; a.pad = p.pad; a.out_h = p.out_h; a.out_w = p.out_w;
  a.p_total = (long long)p.n_img * p.out_h * p.out_w;
  a.k_cols = p.kh * p.kw * p.c_in;
  if (a.p_total == 0) return 0;
  const int gx = ceil_div(a.k_cols, 64), gy = ceil_div(p.c_out, 64);
  long splits = (6L * num_sms() + (long)gx * gy - 1) / ((long)gx * gy);
  const long max_splits = (a.p_total + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  a.p_per_split = (int)(((a.p_total + splits - 1) / splits + 15) / 16 * 16);
  const int gz = ceil_div(a.p_total, a.p_per_split);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  DISPATCH_T(p.x_dtype, (simt_wgrad_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(a)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                             double* hyper, long long n, int decoupled, float grad_scale,
                             void* stream) {
  RMV_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && hyper, "adam_step: null pointer");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, s>>>(hyper);
  long blocks = (n / 4 + 255) / 256;
  if (blocks > 16L * num_sms()) blocks = 16L * num_sms();
  rmv::launch_pdl(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, s, params, grads, exp_avg, exp_avg_sq, hyper, decoupled,
                                               grad_scale, n);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_maxpool3x3s2_fwd_idx(const void* x, void* y, void* idx, int n_img, int in_h,
                                        int in_w, int c, int dtype, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0, "maxpool_fwd_idx: c must be a multiple of 8");
  const int out_h = (in_h - 1) / 2 + 1, out_w = (in_w - 1) / 2 + 1;
  const long long total = (long long)n_img * out_h * out_w * (c / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (rmv::launch_pdl(maxpool_fwd_idx_kernel<T>, dim3(nblk(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const T*)x, (T*)y, (uint8_t*)idx, in_h, in_w, c, out_h, out_w, total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_maxpool3x3s2_bwd_idx(const void* idx, const void* dy, void* dx, int n_img,
                                        int in_h, int in_w, int c, int dtype, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0, "maxpool_bwd_idx: c must be a multiple of 8");
  const int out_h = (in_h - 1) / 2 + 1, out_w = (in_w - 1) / 2 + 1;
  const int blk_h = (in_h + 1) / 2, blk_w = (in_w + 1) / 2;
  const long long total = (long long)n_img * blk_h * blk_w * (c / 8);
  if (total == 0) return 0;
  DISPATCH_T(dtype, (rmv::launch_pdl(maxpool_bwd_idx_kernel<T>, dim3(nblk(total, 256)), dim3(256), 0,
                                     (cudaStream_t)stream, (const uint8_t*)idx, (const T*)dy, (T*)dx,
                                     in_h, in_w, c, out_h, out_w, blk_h, blk_w, total)));
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_permute_cast_batch(const rmv_permute_job* jobs_dev, int n_jobs,
                                      unsigned total_blocks, void* stream) {
  RMV_CHECK_ARG(jobs_dev != nullptr || n_jobs == 0, "permute_cast_batch: null job table");
  if (n_jobs == 0 || total_blocks == 0) return 0;
  rmv::launch_pdl(permute_cast_batch_kernel, dim3(total_blocks), dim3(256), 0, (cudaStream_t)stream, jobs_dev, n_jobs);
  RMV_LAUNCH_CHECK();
  return 0;
}
