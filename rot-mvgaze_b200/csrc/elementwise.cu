// HBM-bound kernels of the Rot-MV path: layout changes, pooling, the rotation-constrained
// cross-view gather, the gaze-head tail + angular loss, metrics and pose->rotation conversion.
// All are one-pass, 128-bit vectorised along the contiguous (channel / feature) axis.
#include "common.cuh"
#include "ops.h"

#include <math_constants.h>

namespace rmv {
namespace {

constexpr float kRadToDeg = 57.29577951308232f;  // 180 / pi

// ---- 8-wide vector helpers: bf16 x8 (16 B) or fp32 x8 (2 x 16 B) -> float[8] ----------------
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  typedef uint4 Raw;  // packed form: unrolled loads stay in 4 registers until they are consumed
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
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
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
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
template <> struct Vec8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float* f) {
    f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
    f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// ---------------------------------------------------------------------------------------------
// Stem im2col (fp32 NCHW -> [pixels, k_pad])
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void stem_im2col_kernel(const float* __restrict__ x, T* __restrict__ a, int n_img,
                                   int c_in, int in_h, int in_w, int kh, int kw, int stride,
                                   int pad, int out_h, int out_w, int k_pad, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int chunks = k_pad / 8;
  const int chunk = (int)(idx % chunks);
  const long long m = idx / chunks;
  const int ow = (int)(m % out_w);
  const long long t = m / out_w;
  const int oh = (int)(t % out_h);
  const int n = (int)(t / out_h);
  const int k_real = kh * kw * c_in;
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = chunk * 8 + i;
    float v = 0.f;
    if (k < k_real) {
      const int tap = k / c_in, c = k - tap * c_in;
      const int r = tap / kw, s = tap - r * kw;
      const int ih = oh * stride - pad + r, iw = ow * stride - pad + s;
      if (ih >= 0 && ih < in_h && iw >= 0 && iw < in_w)
        v = __ldg(x + (((long long)n * c_in + c) * in_h + ih) * in_w + iw);
    }
    f[i] = v;
  }
  Vec8<T>::store(a + m * k_pad + chunk * 8, f);
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int c,
                                    long long hw, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw
  if (idx >= total) return;
  const long long n = idx / hw, p = idx % hw;
  for (int ch = 0; ch < c; ++ch) {
    const float v = __ldg(x + (n * c + ch) * hw + p);
    if constexpr (sizeof(T) == 2) y[idx * c + ch] = __float2bfloat16_rn(v);
    else y[idx * c + ch] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool 3x3 s2 p1, NHWC, 8 channels per thread
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int in_h, int in_w,
                               int c, int out_h, int out_w, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = c / 8;
  const int cc = (int)(idx % cv) * 8;
  long long t = idx / cv;
  const int ow = (int)(t % out_w); t /= out_w;
  const int oh = (int)(t % out_h);
  const long long n = t / out_h;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -CUDART_INF_F;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * 2 - 1 + r;
    if (ih < 0 || ih >= in_h) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int iw = ow * 2 - 1 + s;
      if (iw < 0 || iw >= in_w) continue;
      float f[8];
      Vec8<T>::load(x + ((n * in_h + ih) * in_w + iw) * c + cc, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
    }
  }
  Vec8<T>::store(y + ((n * out_h + oh) * out_w + ow) * c + cc, m);
}

// bf16 fast path: each thread produces TWO horizontally adjacent outputs (8 channels each) from a
// 3 x 5 input window (15 instead of 18 16-byte loads), all loads issued before the first use, and
// the maximum taken on packed bf16 pairs (max of bf16 values is exact, no fp32 round trip).
__global__ void __launch_bounds__(256)
maxpool_bf16x2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int in_h,
                      int in_w, int c, int out_h, int out_w2 /* ceil(out_w/2) */, int out_w,
                      long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = c / 8;
  const int cc = (int)(idx % cv) * 8;
  long long t = idx / cv;
  const int owp = (int)(t % out_w2); t /= out_w2;
  const int oh = (int)(t % out_h);
  const long long n = t / out_h;
  const int ow = owp * 2;
  const uint32_t ninf = 0xFF80FF80u;  // (-inf, -inf)
  uint4 v[3][5];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * 2 - 1 + r;
    const bool rok = ih >= 0 && ih < in_h;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int iw = ow * 2 - 1 + q;
      v[r][q] = (rok && iw >= 0 && iw < in_w)
                    ? *reinterpret_cast<const uint4*>(x + ((n * in_h + ih) * in_w + iw) * c + cc)
                    : make_uint4(ninf, ninf, ninf, ninf);
    }
  }
  auto mx = [](uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  };
  auto mx4 = [&](uint4 a, uint4 b) { return make_uint4(mx(a.x, b.x), mx(a.y, b.y), mx(a.z, b.z), mx(a.w, b.w)); };
  uint4 col[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) col[q] = mx4(mx4(v[0][q], v[1][q]), v[2][q]);
  const uint4 o0 = mx4(mx4(col[0], col[1]), col[2]);
  __nv_bfloat16* yo = y + ((n * out_h + oh) * out_w + ow) * c + cc;
  *reinterpret_cast<uint4*>(yo) = o0;
  if (ow + 1 < out_w) *reinterpret_cast<uint4*>(yo + c) = mx4(mx4(col[2], col[3]), col[4]);
}

// ---------------------------------------------------------------------------------------------
// Global average pool, NHWC [n, hw, c] -> [n, c] (one or two destinations, row strides ld0/ld1)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, int hw, int c, T* __restrict__ y0,
                               long long ld0, T* __restrict__ y1, long long ld1, long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = c / 8;
  const int cc = (int)(idx % cv) * 8;
  const long long n = idx / cv;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* p = x + n * hw * c + cc;
  for (int i = 0; i < hw; ++i) {
    float f[8];
    Vec8<T>::load(p + (long long)i * c, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
  const float inv = 1.f / (float)hw;
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] *= inv;
  Vec8<T>::store(y0 + n * ld0 + cc, s);
  if (y1 != nullptr) Vec8<T>::store(y1 + n * ld1 + cc, s);
}

// ---------------------------------------------------------------------------------------------
// Rotation-constrained cross-view gather
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__global__ void rotate_gather_kernel(const T* __restrict__ feat, long long ld_feat,
                                     const float* __restrict__ rot, T* __restrict__ dst,
                                     long long ld_dst, int views, int nvec, int apply_rot,
                                     long long total) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int kv = nvec / 8;
  const int k0 = (int)(idx % kv) * 8;
  const long long row = idx / kv;  // b*V + v
  const int v = (int)(row % views);
  const long long b = row / views;
  float o[3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) o[r][i] = 0.f;
  for (int u = 0; u < views; ++u) {
    if (u == v) continue;
    const T* fp = feat + (b * views + u) * ld_feat + k0;
    float f[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) Vec8<T>::load(fp + (long long)c * nvec, f[c]);
    float R[9];
    if (apply_rot & 2) {
      // backward of the gather: this row (self = v) receives rot[b,u,v]^T applied to d/dA of row u
      const float* rp = rot + ((b * views + u) * views + v) * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + (i % 3) * 3 + i / 3);
    } else if (apply_rot & 1) {
      const float* rp = rot + ((b * views + v) * views + u) * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + i);
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = R[r * 3 + 0] * f[0][i];
        t = fmaf(R[r * 3 + 1], f[1][i], t);
        t = fmaf(R[r * 3 + 2], f[2][i], t);
        o[r][i] += t;
      }
  }
  if (views > 2) {
    const float inv = 1.f / (float)(views - 1);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[r][i] *= inv;
  }
  T* dp = dst + row * ld_dst + k0;
#pragma unroll
  for (int r = 0; r < 3; ++r) Vec8<T>::store(dp + (long long)r * nvec, o[r]);
}

// More than two views: every feature row of a sample is a partner of the V-1 other
// rows. One thread = 8 feature columns of ALL views of one sample: the V x 3 slices are copied once
// with cp.async into a per-thread scratch in shared memory (all V*3 16-byte copies in flight at
// once, no registers held, no barrier: a thread reads back only what it copied), then the V outputs
// are produced one after the other from that scratch -- HBM sees every row once instead of V-1
// times. Partners are added in ascending view order, as in the general kernel.
template <typename T>
__global__ void __launch_bounds__(256)
rotate_gather_staged_kernel(const T* __restrict__ feat, long long ld_feat,
                            const float* __restrict__ rot, T* __restrict__ dst, long long ld_dst,
                            int views, int nvec, int apply_rot, long long total /* batch * nvec/8 */) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  using Raw = typename Vec8<T>::Raw;
  constexpr int kPieces = (int)sizeof(Raw) / 16;
  extern __shared__ uint4 s_stage[];   // [view][component][piece][256 threads]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int kv = nvec / 8;
  const int k0 = (int)(idx % kv) * 8;
  const long long b = idx / kv;
  const T* fbase = feat + b * views * ld_feat + k0;
  for (int u = 0; u < views; ++u)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const char* src = reinterpret_cast<const char*>(fbase + u * ld_feat + (long long)c * nvec);
#pragma unroll
      for (int q = 0; q < kPieces; ++q)
        cp_async_16(s_stage + ((u * 3 + c) * kPieces + q) * 256 + threadIdx.x, src + 16 * q);
    }
  cp_async_commit();
  cp_async_wait<0>();
  const float inv = 1.f / (float)(views - 1);
  for (int v = 0; v < views; ++v) {
    float o[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[r][i] = 0.f;
    for (int u = 0; u < views; ++u) {
      if (u == v) continue;
      float f[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Raw raw;
        uint4* rp = reinterpret_cast<uint4*>(&raw);
#pragma unroll
        for (int q = 0; q < kPieces; ++q) rp[q] = s_stage[((u * 3 + c) * kPieces + q) * 256 + threadIdx.x];
        Vec8<T>::unpack(raw, f[c]);
      }
      float R[9];
      if (apply_rot & 2) {   // backward of the gather: rot[b,u,v]^T
        const float* rp = rot + ((b * views + u) * views + v) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + (i % 3) * 3 + i / 3);
      } else if (apply_rot & 1) {
        const float* rp = rot + ((b * views + v) * views + u) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + i);
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = R[r * 3 + 0] * f[0][i];
          t = fmaf(R[r * 3 + 1], f[1][i], t);
          t = fmaf(R[r * 3 + 2], f[2][i], t);
          o[r][i] += t;
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[r][i] *= inv;
    T* dp = dst + (b * views + v) * ld_dst + k0;
#pragma unroll
    for (int r = 0; r < 3; ++r) Vec8<T>::store(dp + (long long)r * nvec, o[r]);
  }
}

// Two views (the reference's configuration, models/rot_mv.py:234,238): persistent grid-stride
// kernel. One item = 8 feature columns of both views of one sample: the two partner slices (6 x
// 16 B) are loaded together, the NEXT item's six loads are issued before this item is computed, and
// both directions out_0 = R_01 F_1, out_1 = R_10 F_0 are produced from them (6 x 16 B stores).
template <typename T>
__global__ void __launch_bounds__(128)
rotate_gather_pair_kernel(const T* __restrict__ feat, long long ld_feat,
                          const float* __restrict__ rot, T* __restrict__ dst, long long ld_dst,
                          int nvec, int apply_rot, long long total /* batch * nvec/8 */) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  const int kv = nvec / 8;
  const long long step = (long long)gridDim.x * blockDim.x;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  typename Vec8<T>::Raw cur[2][3], nxt[2][3];
  auto issue = [&](long long item, typename Vec8<T>::Raw (&r)[2][3]) {
    const int k0 = (int)(item % kv) * 8;
    const T* fp = feat + (item / kv) * 2 * ld_feat + k0;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int c = 0; c < 3; ++c) r[u][c] = Vec8<T>::load_raw(fp + u * ld_feat + (long long)c * nvec);
  };
  issue(idx, cur);
#pragma unroll 1
  for (; idx < total; idx += step) {
    const bool more = idx + step < total;
    if (more) issue(idx + step, nxt);
    const int k0 = (int)(idx % kv) * 8;
    const long long b = idx / kv;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int u = 1 - v;
      float R[9];
      if (apply_rot & 2) {   // backward of the gather: rot[b,u,v]^T
        const float* rp = rot + ((b * 2 + u) * 2 + v) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + (i % 3) * 3 + i / 3);
      } else if (apply_rot & 1) {
        const float* rp = rot + ((b * 2 + v) * 2 + u) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rp + i);
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.f : 0.f;
      }
      float f[3][8], o[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c) Vec8<T>::unpack(cur[u][c], f[c]);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a = R[r * 3 + 0] * f[0][i];
          a = fmaf(R[r * 3 + 1], f[1][i], a);
          o[r][i] = fmaf(R[r * 3 + 2], f[2][i], a) + 0.f;   // same rounding as 0 + t of the general kernel
        }
      T* dp = dst + (b * 2 + v) * ld_dst + k0;
#pragma unroll
      for (int r = 0; r < 3; ++r) Vec8<T>::store(dp + (long long)r * nvec, o[r]);
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[u][c] = nxt[u][c];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gaze head tail (hid -> 2 GEMV) + pitch-yaw -> vector + angular loss
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pitchyaw_to_vec(float p, float y, float* v) {
  float sp, cp, sy, cy;
  sincosf(p, &sp, &cp);
  sincosf(y, &sy, &cy);
  v[0] = cp * sy; v[1] = sp; v[2] = cp * cy;
}
// F.cosine_similarity(a, b, eps) of torch >= 1.12: normalise each vector first.
__device__ __forceinline__ float cos_sim_torch(const float* a, const float* b, float eps) {
  const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), eps);
  const float nb = fmaxf(sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), eps);
  return (a[0] / na) * (b[0] / nb) + (a[1] / na) * (b[1] / nb) + (a[2] / na) * (b[2] / nb);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp finishes kHeadRows rows per iteration (grid-stride). NCH = hid / 256 (1 or 2: the
// reference head has hid = 512): the rows are staged through a per-thread ring in shared memory
// with cp.async (kHeadStages groups of kHeadRows rows in flight per warp, no registers held by loads
// in flight; every thread reads back exactly the 16-byte pieces it copied, so the ring needs no
// barrier), the eight dot products of a group are reduced with a multi-value butterfly, and lanes
// 0, 4, 8, 12 each do the transcendental tail (pitch-yaw -> vector, cosine, acos) of one row.
// NCH = 0 is the generic form (any hid) without the ring. One loss atomic per block.
constexpr int kHeadRows = 4;     // the shuffle tree in head_loss_kernel is written for 4
constexpr int kHeadStages = 3;


// prediction write + loss term of one row per lane (models/backbones/blocks.py:41-60 output,
// utils/math.py:52-60, losses/gaze_loss.py:42-52)
__device__ __forceinline__ void head_tail(float p0, float p1, int row, int rows,
                                          float* __restrict__ pred, const float* __restrict__ gt,
                                          int views, float aux_decay, float& ang) {
  if (row >= rows) return;
  *reinterpret_cast<float2*>(pred + (long long)row * 2) = make_float2(p0, p1);
  if (gt != nullptr) {
    float vg[3], vp[3];
    const float2 g2 = __ldg(reinterpret_cast<const float2*>(gt + (long long)row * 2));
    pitchyaw_to_vec(g2.x, g2.y, vg);
    pitchyaw_to_vec(p0, p1, vp);
    float sim = cos_sim_torch(vg, vp, 1e-6f);
    sim = fminf(fmaxf(sim, -1.f), 1.f);
    ang += acosf(sim) * kRadToDeg * ((row % views) == 0 ? 1.f : aux_decay);
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256, 2)
head_loss_kernel(const T* __restrict__ hidden, long long ld, const float* __restrict__ w2,
                 const float* __restrict__ b2, int rows, int hid, float* __restrict__ pred,
                 const float* __restrict__ gt, float loss_scale, int views, float aux_decay,
                 float* __restrict__ loss_out) {
  griddep_wait();    // PDL: predecessors complete + visible
  griddep_launch();  // let the next kernel of the stream get scheduled

  using Raw = typename Vec8<T>::Raw;
  constexpr int R = NCH > 0 ? NCH : 1;
  constexpr int kPieces = (int)sizeof(Raw) / 16;           // 16-byte pieces per 8-element vector
  extern __shared__ uint4 s_ring[];                        // [stage][row][chunk][piece][256 threads]
  __shared__ float s_part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float bias0 = __ldg(b2), bias1 = __ldg(b2 + 1);
  const int step = gridDim.x * 8 * kHeadRows;
  float ang = 0.f;
  auto slot = [&](int stage, int r, int j, int piece) -> uint4* {
    return s_ring + ((((stage * kHeadRows + r) * R + j) * kPieces + piece) * 256 + threadIdx.x);
  };
  auto issue = [&](int row0, int stage) {   // always commits (possibly empty) so group counts line up
    if (row0 < rows) {
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        const int row = min(row0 + r, rows - 1);  // past the end: re-read the last row, result dropped
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const char* src = reinterpret_cast<const char*>(hidden + (long long)row * ld + j * 256 + lane * 8);
#pragma unroll
          for (int q = 0; q < kPieces; ++q) cp_async_16(slot(stage, r, j, q), src + 16 * q);
        }
      }
    }
    cp_async_commit();
  };
  int row0 = (blockIdx.x * 8 + warp) * kHeadRows;
  if (NCH > 0) {
#pragma unroll
    for (int st = 0; st < kHeadStages - 1; ++st) issue(row0 + st * step, st);
  }
  int stage = 0, pend = 0, qrow = rows;
  float q0 = 0.f, q1 = 0.f;
#pragma unroll 1
  for (; row0 < rows; row0 += step) {
    float d0[kHeadRows], d1[kHeadRows];
    if (NCH > 0) {
      {
        int st_next = stage + kHeadStages - 1;
        if (st_next >= kHeadStages) st_next -= kHeadStages;
        issue(row0 + (kHeadStages - 1) * step, st_next);
      }
      cp_async_wait<kHeadStages - 1>();   // this thread's copies of the current group have landed
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) { d0[r] = 0.f; d1[r] = 0.f; }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float a[8], b[8];                  // 4 KB weight table: L1-resident, read once per group
        Vec8<float>::load(w2 + j * 256 + lane * 8, a);
        Vec8<float>::load(w2 + hid + j * 256 + lane * 8, b);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) {
          Raw raw;
          uint4* rp = reinterpret_cast<uint4*>(&raw);
#pragma unroll
          for (int q = 0; q < kPieces; ++q) rp[q] = *slot(stage, r, j, q);
          float h[8];
          Vec8<T>::unpack(raw, h);
#pragma unroll
          for (int i = 0; i < 8; ++i) { d0[r] = fmaf(h[i], a[i], d0[r]); d1[r] = fmaf(h[i], b[i], d1[r]); }
        }
      }
      if (++stage == kHeadStages) stage = 0;
    } else {
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        d0[r] = 0.f; d1[r] = 0.f;
        const T* hp = hidden + (long long)min(row0 + r, rows - 1) * ld;
        for (int k = lane * 8; k < hid; k += 256) {
          float h[8], a[8], b[8];
          Vec8<T>::load(hp + k, h);
          Vec8<float>::load(w2 + k, a);
          Vec8<float>::load(w2 + hid + k, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) { d0[r] = fmaf(h[i], a[i], d0[r]); d1[r] = fmaf(h[i], b[i], d1[r]); }
        }
      }
    }
    // eight warp sums in 9 shuffles: each butterfly level halves the number of values a lane still
    // carries (same pairing and order of additions as warp_sum, so the sums are bit-identical);
    // afterwards lane 4k holds value k (k < 4: d0[k], k >= 4: d1[k-4]); one more shuffle brings
    // d1[r] next to d0[r] in lane 4r, which does the tail of row row0 + r
    float x4[4], x2[2], x1;
    {
      const bool up = lane & 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float keep = up ? d1[i] : d0[i], send = up ? d0[i] : d1[i];
        x4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool up = lane & 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float keep = up ? x4[i + 2] : x4[i], send = up ? x4[i] : x4[i + 2];
        x2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool up = lane & 4;
      const float keep = up ? x2[1] : x2[0], send = up ? x2[0] : x2[1];
      x1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    x1 += __shfl_xor_sync(0xffffffffu, x1, 2);
    x1 += __shfl_xor_sync(0xffffffffu, x1, 1);
    const float p0 = x1 + bias0;
    const float p1 = __shfl_down_sync(0xffffffffu, x1, 16) + bias1;
    // park the four results in lanes 4*pend .. 4*pend+3; after eight groups every lane owns one
    // row and the transcendental tail runs once for 32 rows with all lanes active
    const float s0 = __shfl_sync(0xffffffffu, p0, 4 * (lane & 3));
    const float s1 = __shfl_sync(0xffffffffu, p1, 4 * (lane & 3));
    if ((lane >> 2) == pend) { q0 = s0; q1 = s1; qrow = row0 + (lane & 3); }
    if (++pend == 8) {
      head_tail(q0, q1, qrow, rows, pred, gt, views, aux_decay, ang);
      pend = 0; qrow = rows;
    }
  }
  if (pend != 0) head_tail(q0, q1, qrow, rows, pred, gt, views, aux_decay, ang);
  if (NCH > 0) cp_async_wait<0>();
  if (gt != nullptr) {
    ang = warp_sum(ang);
    if (lane == 0) s_part[warp] = ang;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += s_part[i];
      atomicAdd(loss_out, s * loss_scale);
    }
  }
}

__global__ void angular_error_kernel(const float* __restrict__ pred, long long ld_pred,
                                     const float* __restrict__ gt, long long ld_gt, int rows,
                                     float* __restrict__ err_sum) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float e = 0.f;
  if (row < rows) {
    float a[3], b[3];
    pitchyaw_to_vec(__ldg(pred + row * ld_pred), __ldg(pred + row * ld_pred + 1), a);
    pitchyaw_to_vec(__ldg(gt + row * ld_gt), __ldg(gt + row * ld_gt + 1), b);
    const float ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-7f);
    const float nb = fmaxf(sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]), 1e-7f);
    // The reference metric does not clamp (utils/math.py:118-120,135-137) and can return NaN
    // when rounding pushes the similarity past 1; the device metric clamps.
    const float sim = fminf(fmaxf(ab / (na * nb), -1.f), 1.f);
    e = acosf(sim) * kRadToDeg;
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(err_sum, e);
    const int first = blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    const int cnt = max(0, min(32, rows - first));
    atomicAdd(err_sum + 1, (float)cnt);
  }
}

__global__ void pose_to_rot_kernel(const float* __restrict__ pose, float* __restrict__ rot,
                                   int views, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, i, j)
  if (idx >= total) return;
  const int j = (int)(idx % views);
  const int i = (int)((idx / views) % views);
  const long long b = idx / ((long long)views * views);
  float R[2][9];
  const int vi[2] = {i, j};
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float* p = pose + (b * views + vi[t]) * 2;
    float sp, cp, sy, cy;
    sincosf(-__ldg(p), &sp, &cp);   // head-pose convention: pitch enters negated
    sincosf(__ldg(p + 1), &sy, &cy);
    // R = R_y(yaw) * R_x(-pitch)
    R[t][0] = cy;  R[t][1] = sy * sp;  R[t][2] = sy * cp;
    R[t][3] = 0.f; R[t][4] = cp;       R[t][5] = -sp;
    R[t][6] = -sy; R[t][7] = cy * sp;  R[t][8] = cy * cp;
  }
  float* o = rot + idx * 9;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = R[0][r * 3 + 0] * R[1][c * 3 + 0];
      s = fmaf(R[0][r * 3 + 1], R[1][c * 3 + 1], s);
      s = fmaf(R[0][r * 3 + 2], R[1][c * 3 + 2], s);
      o[r * 3 + c] = s;
    }
}

__global__ void relative_rot_kernel(const float* __restrict__ rot, float* __restrict__ out,
                                    int views, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (b, i, j)
  if (idx >= total) return;
  const int j = (int)(idx % views);
  const int i = (int)((idx / views) % views);
  const long long b = idx / ((long long)views * views);
  const float* ri = rot + (b * views + i) * 9;
  const float* rj = rot + (b * views + j) * 9;
  float* o = out + idx * 9;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = __ldg(ri + r * 3) * __ldg(rj + c * 3);
      s = fmaf(__ldg(ri + r * 3 + 1), __ldg(rj + c * 3 + 1), s);
      s = fmaf(__ldg(ri + r * 3 + 2), __ldg(rj + c * 3 + 2), s);
      o[r * 3 + c] = s;
    }
}

inline unsigned blocks_for(long long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}

}  // namespace
}  // namespace rmv

using namespace rmv;

extern "C" int rmv_stem_im2col(const float* x, void* a, int n_img, int c_in, int in_h, int in_w,
                               int kh, int kw, int stride, int pad, int out_h, int out_w,
                               int k_pad, int a_dtype, void* stream) {
  RMV_CHECK_ARG(k_pad % 8 == 0 && k_pad >= kh * kw * c_in, "stem_im2col: bad k_pad %d", k_pad);
  const long long total = (long long)n_img * out_h * out_w * (k_pad / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (a_dtype == RMV_DTYPE_BF16)
    stem_im2col_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, s>>>(
        x, (__nv_bfloat16*)a, n_img, c_in, in_h, in_w, kh, kw, stride, pad, out_h, out_w, k_pad, total);
  else
    stem_im2col_kernel<float><<<blocks_for(total, 256), 256, 0, s>>>(
        x, (float*)a, n_img, c_in, in_h, in_w, kh, kw, stride, pad, out_h, out_w, k_pad, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_nchw_to_nhwc(const float* x, void* y, int n_img, int c, int h, int w,
                                int y_dtype, void* stream) {
  const long long hw = (long long)h * w, total = hw * n_img;
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (y_dtype == RMV_DTYPE_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks_for(total, 256), 256, 0, s>>>(x, (__nv_bfloat16*)y, c, hw, total);
  else
    nchw_to_nhwc_kernel<float><<<blocks_for(total, 256), 256, 0, s>>>(x, (float*)y, c, hw, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_maxpool3x3s2_fwd(const void* x, void* y, int n_img, int in_h, int in_w, int c,
                                    int dtype, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0, "maxpool: c=%d must be a multiple of 8", c);
  const int out_h = (in_h + 2 - 3) / 2 + 1, out_w = (in_w + 2 - 3) / 2 + 1;
  const long long total = (long long)n_img * out_h * out_w * (c / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RMV_DTYPE_BF16) {
    const int out_w2 = (out_w + 1) / 2;
    const long long total2 = (long long)n_img * out_h * out_w2 * (c / 8);
    rmv::launch_pdl(maxpool_bf16x2_kernel, dim3(blocks_for(total2, 256)), dim3(256), 0, s,
                    (const __nv_bfloat16*)x, (__nv_bfloat16*)y, in_h, in_w, c, out_h, out_w2, out_w,
                    total2);
  } else
    rmv::launch_pdl(maxpool_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, s, (const float*)x, (float*)y, in_h, in_w, c, out_h, out_w, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_avgpool_fwd(const void* x, int n_img, int hw, int c, int dtype, void* y0,
                               long long ld0, void* y1, long long ld1, void* stream) {
  RMV_CHECK_ARG(c % 8 == 0 && ld0 % 8 == 0 && ld1 % 8 == 0, "avgpool: c/ld must be multiples of 8");
  const long long total = (long long)n_img * (c / 8);
  if (total == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == RMV_DTYPE_BF16)
    rmv::launch_pdl(avgpool_kernel<__nv_bfloat16>, dim3(blocks_for(total, 128)), dim3(128), 0, s, 
        (const __nv_bfloat16*)x, hw, c, (__nv_bfloat16*)y0, ld0, (__nv_bfloat16*)y1, ld1, total);
  else
    rmv::launch_pdl(avgpool_kernel<float>, dim3(blocks_for(total, 128)), dim3(128), 0, s, (const float*)x, hw, c, (float*)y0, ld0, (float*)y1, ld1, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_rotate_gather_fwd(const void* feat, long long ld_feat, const float* rot,
                                     void* dst, long long ld_dst, int batch, int views, int nvec,
                                     int dtype, int apply_rot, void* stream) {
  RMV_CHECK_ARG(views >= 2, "rotate_gather: need >= 2 views, got %d", views);
  RMV_CHECK_ARG(nvec % 8 == 0 && ld_feat % 8 == 0 && ld_dst % 8 == 0,
                "rotate_gather: nvec/ld must be multiples of 8");
  cudaStream_t s = (cudaStream_t)stream;
  if (views == 2) {
    // persistent pair kernel: as many blocks as stay resident (register-bound), grid-stride
    const long long total = (long long)batch * (nvec / 8);
    if (total == 0) return 0;
    long long blocks = (total + 127) / 128;
    if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
    if (dtype == RMV_DTYPE_BF16)
      rmv::launch_pdl(rotate_gather_pair_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(128), 0, s,
                      (const __nv_bfloat16*)feat, ld_feat, rot, (__nv_bfloat16*)dst, ld_dst, nvec,
                      apply_rot, total);
    else
      rmv::launch_pdl(rotate_gather_pair_kernel<float>, dim3((unsigned)blocks), dim3(128), 0, s,
                      (const float*)feat, ld_feat, rot, (float*)dst, ld_dst, nvec, apply_rot, total);
    RMV_LAUNCH_CHECK();
    return 0;
  }
  const int esz = dtype == RMV_DTYPE_BF16 ? 2 : 4;
  const int stage_bytes = views * 3 * 8 * esz * 256;   // V x 12 KB (bf16) / V x 24 KB (fp32) per block
  if (views <= 4) {
    // 3 or 4 views: every row read from HBM once, staged per thread in shared memory (measured at
    // B=32768, V=4: 239 us vs 271 us for the general kernel). From 5 views on the V(V-1) 3x3
    // products per column make the kernel instruction-bound and the staging's lower occupancy costs
    // more than the saved L2 re-reads (V=8: 568 us staged vs 476 us general), so those stay general.
    const long long total = (long long)batch * (nvec / 8);
    if (total == 0) return 0;
    if (dtype == RMV_DTYPE_BF16) {
      if (stage_bytes > 48 * 1024)
        RMV_CUDA(cudaFuncSetAttribute(rotate_gather_staged_kernel<__nv_bfloat16>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      rmv::launch_pdl(rotate_gather_staged_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256),
                      stage_bytes, s, (const __nv_bfloat16*)feat, ld_feat, rot, (__nv_bfloat16*)dst,
                      ld_dst, views, nvec, apply_rot, total);
    } else {
      if (stage_bytes > 48 * 1024)
        RMV_CUDA(cudaFuncSetAttribute(rotate_gather_staged_kernel<float>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      rmv::launch_pdl(rotate_gather_staged_kernel<float>, dim3(blocks_for(total, 256)), dim3(256),
                      stage_bytes, s, (const float*)feat, ld_feat, rot, (float*)dst, ld_dst, views, nvec,
                      apply_rot, total);
    }
    RMV_LAUNCH_CHECK();
    return 0;
  }
  // general kernel: one short thread per (row, 8 columns); the V-1 re-reads of a partner row by
  // the rows of the same sample hit L1/L2
  const long long total = (long long)batch * views * (nvec / 8);
  if (total == 0) return 0;
  if (dtype == RMV_DTYPE_BF16)
    rmv::launch_pdl(rotate_gather_kernel<__nv_bfloat16>, dim3(blocks_for(total, 128)), dim3(128), 0, s,
        (const __nv_bfloat16*)feat, ld_feat, rot, (__nv_bfloat16*)dst, ld_dst, views, nvec, apply_rot, total);
  else
    rmv::launch_pdl(rotate_gather_kernel<float>, dim3(blocks_for(total, 128)), dim3(128), 0, s,
        (const float*)feat, ld_feat, rot, (float*)dst, ld_dst, views, nvec, apply_rot, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_head_loss_fwd(const void* hidden, long long ld_hidden, int hid_dtype,
                                 const float* w2, const float* b2, int rows, int hid, float* pred,
                                 const float* gt, float loss_scale, int views,
                                 float aux_decay, float* loss_out, void* stream) {
  RMV_CHECK_ARG(hid % 8 == 0 && ld_hidden % 8 == 0, "head_loss: hid/ld must be multiples of 8");
  RMV_CHECK_ARG(gt == nullptr || loss_out != nullptr, "head_loss: gt given without loss_out");
  RMV_CHECK_ARG(views >= 1, "head_loss: views must be >= 1");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  // grid-stride over groups of 8 warps x kHeadRows rows; one loss atomic per block
  long blocks = ((long)rows + 8 * kHeadRows - 1) / (8 * kHeadRows);
  if (blocks > 2L * num_sms()) blocks = 2L * num_sms();
  const dim3 grid((unsigned)blocks), block(256);
  const int nch = hid == 512 ? 2 : (hid == 256 ? 1 : 0);
  const int esz = hid_dtype == RMV_DTYPE_BF16 ? 2 : 4;
  const int smem = nch ? kHeadStages * kHeadRows * nch * 8 * esz * 256 : 0;   // bf16/512: 96 KB (2 blocks/SM), fp32/512: 192 KB
#define RMV_HEAD_LAUNCH(T, NCH)                                                                    \
  do {                                                                                             \
    static int attr_done = 0; /* benign race: the attribute is idempotent */                       \
    if (smem > 48 * 1024 && !attr_done) {                                                          \
      RMV_CUDA(cudaFuncSetAttribute(head_loss_kernel<T, NCH>,                                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));            \
      attr_done = 1;                                                                               \
    }                                                                                              \
    rmv::launch_pdl(head_loss_kernel<T, NCH>, grid, block, smem, s, (const T*)hidden, ld_hidden,   \
                    w2, b2, rows, hid, pred, gt, loss_scale, views, aux_decay, loss_out);          \
  } while (0)
  if (hid_dtype == RMV_DTYPE_BF16) {
    if (nch == 2) RMV_HEAD_LAUNCH(__nv_bfloat16, 2);
    else if (nch == 1) RMV_HEAD_LAUNCH(__nv_bfloat16, 1);
    else RMV_HEAD_LAUNCH(__nv_bfloat16, 0);
  } else {
    if (nch == 2) RMV_HEAD_LAUNCH(float, 2);
    else if (nch == 1) RMV_HEAD_LAUNCH(float, 1);
    else RMV_HEAD_LAUNCH(float, 0);
  }
#undef RMV_HEAD_LAUNCH
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_angular_error_accum(const float* pred, long long ld_pred, const float* gt,
                                       long long ld_gt, int rows, float* err_sum, void* stream) {
  if (rows == 0) return 0;
  angular_error_kernel<<<blocks_for(rows, 128), 128, 0, (cudaStream_t)stream>>>(pred, ld_pred, gt, ld_gt, rows, err_sum);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_pose_to_rotations(const float* head_pose, float* rotations, int batch,
                                     int views, void* stream) {
  const long long total = (long long)batch * views * views;
  if (total == 0) return 0;
  pose_to_rot_kernel<<<blocks_for(total, 128), 128, 0, (cudaStream_t)stream>>>(head_pose, rotations, views, total);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_relative_rotations(const float* rot, float* rotations, int batch, int views,
                                      void* stream) {
  const long long total = (long long)batch * views * views;
  if (total == 0) return 0;
  relative_rot_kernel<<<blocks_for(total, 128), 128, 0, (cudaStream_t)stream>>>(rot, rotations, views, total);
  RMV_LAUNCH_CHECK();
  return 0;
}
