// Fused ResNet stem: fp32 NCHW image -> 7x7/s2/p3 conv (3->64) -> BatchNorm(eval) -> ReLU -> bf16
// NHWC, in ONE kernel on the tensor cores (reference models/resnet.py:184-188,262-264).
//
// The input has 3 channels, so there is nothing for TMA to tile along K; instead each CTA
//   1. gathers the (21 x 38 x 3) fp32 input patch of a 16x8 output-pixel tile into shared memory
//      as bf16 (the NCHW->NHWC/bf16 conversion the reference never needs is folded in here),
//   2. lays the im2col rows out directly in the 128-byte-swizzled K-major UMMA layout
//      (k = c*56 + kh*8 + kw, 168 real + 24 zero columns = 3 k-blocks of 64),
//   3. issues 12 tcgen05.mma (M=128, N=64, K=16) against the packed filters (TMA-loaded once),
//   4. reads the accumulator from TMEM, applies scale/shift/ReLU, stages bf16 rows in swizzled
//      shared memory and writes the tile with one TMA store.
// Two CTAs per SM overlap each other's gather / MMA / store phases; the im2col matrix
// (4.8 MB per image) never exists in HBM.
#include "common.cuh"
#include "ops.h"

namespace rmv {
namespace {

constexpr int kTW = 16, kTH = 8;        // output tile: 16 wide x 8 high = 128 GEMM rows
constexpr int kCout = 64;
constexpr int kKPad = 192;              // 3 k-blocks of 64
constexpr int kPatchH = 2 * kTH + 5;    // 21 input rows
constexpr int kPatchWUsed = 2 * kTW + 6;  // 38 input columns (incl. the zero-weight kw=7 tap)
constexpr int kPatchPitch = 48;         // elements; 2*pitch*2B = 192 B -> conflict-free LDS.32
constexpr int kPatchElems = 3 * kPatchH * kPatchPitch;  // 3024
constexpr int kThreads = 256;
constexpr int kBBytes = kCout * 128;    // one k-block of filters: 64 rows x 128 B
constexpr int kABytes = 128 * 128;      // one k-block of im2col rows
constexpr int kSmemBytes = 3 * kBBytes + 3 * kABytes + kABytes /*out staging*/ +
                           6144 /*patch*/ + 512 /*scale,shift*/ + 64 /*barriers*/ + 1024 /*align*/;

struct StemArgs {
  CUtensorMap tmap_w;
  CUtensorMap tmap_out;
  const float* x;
  const float* scale;
  const float* shift;
  int n_img, in_h, in_w, out_h, out_w, tiles_w, tiles_h, total_tiles;
  int relu;
};

__device__ __forceinline__ void stem_tma_store(const void* tmap, const void* src, int c0, int c1,
                                               int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

constexpr int kPatchIters = (kPatchElems + kThreads - 1) / kThreads;  // 12

// Issue the (coalesced, bounds-checked) global loads of a tile's input patch into registers...
__device__ __forceinline__ void patch_issue(const StemArgs& a, int tile, int tid,
                                            float (&v)[kPatchIters]) {
  const int tw = tile % a.tiles_w;
  const int th = (tile / a.tiles_w) % a.tiles_h;
  const int n = tile / (a.tiles_w * a.tiles_h);
  const int ih0 = 2 * th * kTH - 3, iw0 = 2 * tw * kTW - 3;
  const float* xn = a.x + (long long)n * 3 * a.in_h * a.in_w;
#pragma unroll
  for (int it = 0; it < kPatchIters; ++it) {
    const int i = tid + it * kThreads;
    float f = 0.f;
    if (i < kPatchElems) {
      const int col = i % kPatchPitch;
      const int r = (i / kPatchPitch) % kPatchH;
      const int c = i / (kPatchPitch * kPatchH);
      const int ih = ih0 + r, iw = iw0 + col;
      if (col < kPatchWUsed && ih >= 0 && ih < a.in_h && iw >= 0 && iw < a.in_w)
        f = __ldg(xn + ((long long)c * a.in_h + ih) * a.in_w + iw);
    }
    v[it] = f;
  }
}
// ... and, one pipeline phase later, convert to bf16 and park them in shared memory.
__device__ __forceinline__ void patch_store(const float (&v)[kPatchIters], __nv_bfloat16* patch,
                                            int tid) {
#pragma unroll
  for (int it = 0; it < kPatchIters; ++it) {
    const int i = tid + it * kThreads;
    if (i < kPatchElems) patch[i] = __float2bfloat16_rn(v[it]);
  }
}

__global__ void __launch_bounds__(kThreads, 2) stem_kernel(const __grid_constant__ StemArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by an integer round trip of the pointer) so the compiler
  // keeps the shared address space and emits LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  uint8_t* sA = sB + 3 * kBBytes;
  uint8_t* sOut = sA + 3 * kABytes;
  __nv_bfloat16* sPatch = reinterpret_cast<__nv_bfloat16*>(sOut + kABytes);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sPatch) + 6144);
  float* s_shift = s_scale + kCout;
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(s_shift + kCout);
  uint64_t* mma_bar = w_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tma_prefetch_desc(&a.tmap_w);
    tma_prefetch_desc(&a.tmap_out);
    mbar_init(w_bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  griddep_wait();    // PDL: global memory is touched from here on
  griddep_launch();
  if (tid < kCout) {
    s_scale[tid] = a.scale ? __ldg(a.scale + tid) : 1.f;
    s_shift[tid] = a.shift ? __ldg(a.shift + tid) : 0.f;
  }
  // zero K-padding columns 168..191 (16-byte units 5,6,7 of k-block 2) never change
  for (int i = tid; i < 128 * 3; i += kThreads) {
    const int row = i / 3;
    const uint32_t j = 5 + i % 3;
    *reinterpret_cast<uint4*>(sA + 2 * kABytes + row * 128 + ((j ^ (uint32_t)(row & 7)) << 4)) =
        make_uint4(0, 0, 0, 0);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  if (tid == 0) {
    mbar_expect_tx(w_bar, 3 * kBBytes);
    for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * kBBytes, &a.tmap_w, w_bar, kb * 64, 0);
  }

  // software pipeline over tiles: patch(T) in smem, patch(T+1) in flight in registers
  float pv[kPatchIters];
  int tile = blockIdx.x;
  if (tile < a.total_tiles) {
    patch_issue(a, tile, tid, pv);
    patch_store(pv, sPatch, tid);
    if (tile + (int)gridDim.x < a.total_tiles) patch_issue(a, tile + gridDim.x, tid, pv);
  }
  __syncthreads();

  const int row = tid & 127, half = tid >> 7;
  const int dy = row >> 4, dx = row & 15;
  const uint32_t sw = (uint32_t)(row & 7);
  const int quarter = warp & 3, colhalf = warp >> 2;
  const int erow = quarter * 32 + lane;
  const uint32_t esw = (uint32_t)(erow & 7);
  constexpr uint32_t idesc = umma_idesc_bf16(128, kCout, 0, 0);
  uint32_t mma_phase = 0;
  bool first = true;

  for (; tile < a.total_tiles; tile += gridDim.x) {
    // ---- im2col rows straight into the swizzled UMMA layout: 16-byte unit q = (c, kh) ----------
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const int q = half * 12 + i;
      if (q < 21) {
        const int c = q / 7, kh = q - c * 7;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(
            sPatch + (c * kPatchH + 2 * dy + kh) * kPatchPitch + 2 * dx);
        const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
        *reinterpret_cast<uint4*>(sA + (q >> 3) * kABytes + row * 128 +
                                  ((((uint32_t)q & 7) ^ sw) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging free
    tc_fence_before_sync();
    __syncthreads();  // SYNC1
    if (tid == 0) {
      tc_fence_after_sync();
      if (first) mbar_wait(w_bar, 0);
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(sA + kb * kABytes), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + kb * kBBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
      }
      umma_commit(mma_bar);
    }
    first = false;
    // ---- the patch buffer is free (everyone is past SYNC1): park the next tile's patch, whose
    //      loads were issued a whole phase ago, and issue the loads of the tile after it ---------
    const int next = tile + gridDim.x;
    if (next < a.total_tiles) patch_store(pv, sPatch, tid);
    if (next + (int)gridDim.x < a.total_tiles) patch_issue(a, next + gridDim.x, tid, pv);
    // ---- epilogue: TMEM -> scale/shift/ReLU -> bf16 -> swizzled staging -> TMA store --------------
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1;
    tc_fence_after_sync();
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + colhalf * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int col = colhalf * 32 + q * 8 + t;
        f[t] = fmaf(__uint_as_float(v[q * 8 + t]), s_scale[col], s_shift[col]);
        if (a.relu) f[t] = fmaxf(f[t], 0.f);
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      const uint32_t j = (uint32_t)(colhalf * 4 + q);
      *reinterpret_cast<uint4*>(sOut + erow * 128 + ((j ^ esw) << 4)) = o;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();  // SYNC2: staging + next patch visible, TMEM drained, sA reusable
    if (tid == 0) {
      const int tw = tile % a.tiles_w;
      const int th = (tile / a.tiles_w) % a.tiles_h;
      const int n = tile / (a.tiles_w * a.tiles_h);
      stem_tma_store(&a.tmap_out, sOut, 0, tw * kTW, th * kTH, n);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 64);
  }
}

// ---------------------------------------------------------------------------------------------
// Stem weight gradient on the tensor cores: dW[col, k] = sum_pixels P[pixel, col] * dZ[pixel, k]
// with P the same in-shared-memory im2col rows the forward builds (col = c*56 + kh*8 + kw) and dZ
// the gradient of the stem conv output (bf16 NHWC, TMA-loaded 16x8-pixel boxes). Both operands are
// MN-major (pixels along K); every CTA accumulates all its tiles in TMEM and adds its partial
// [256 cols][64 k] result to an fp32 scratch with atomics once, at the end.
// ---------------------------------------------------------------------------------------------
struct StemWgArgs {
  CUtensorMap tmap_dz;   // (64 ch, out_w, out_h, n), box (64, 16, 8, 1)
  const float* x;
  float* scratch;        // [192][64] fp32, +=
  int n_img, in_h, in_w, out_h, out_w, tiles_w, tiles_h, total_tiles;
};
constexpr int kWgSmemBytes = 4 * kABytes + 2 * kABytes + 6144 + 64 + 1024;

__global__ void __launch_bounds__(kThreads, 2) stem_wgrad_kernel(const __grid_constant__ StemWgArgs w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // 4 column blocks of [128 px][64 cols] (block 3 = zeros)
  uint8_t* sB = sA + 4 * kABytes;           // 2 x [128 px][64 ch] dZ tiles
  __nv_bfloat16* sPatch = reinterpret_cast<__nv_bfloat16*>(sB + 2 * kABytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPatch) + 6144);  // [2]
  uint64_t* mma_bar = full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // StemArgs view for the shared patch helpers
  StemArgs a;
  a.x = w.x; a.in_h = w.in_h; a.in_w = w.in_w; a.tiles_w = w.tiles_w; a.tiles_h = w.tiles_h;
  a.total_tiles = w.total_tiles;

  if (tid == 0) {
    tma_prefetch_desc(&w.tmap_dz);
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1); mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 128); tmem_relinquish(); }
  // zero block 3 and the K-padding units 5..7 of block 2 once
  for (int i = tid; i < 128 * 8; i += kThreads)
    *reinterpret_cast<uint4*>(sA + 3 * kABytes + i * 16) = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 128 * 3; i += kThreads) {
    const int row = i / 3;
    const uint32_t j = 5 + i % 3;
    *reinterpret_cast<uint4*>(sA + 2 * kABytes + row * 128 + ((j ^ (uint32_t)(row & 7)) << 4)) =
        make_uint4(0, 0, 0, 0);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  auto issue_dz = [&](int tile, int buf) {
    const int tw = tile % w.tiles_w;
    const int th = (tile / w.tiles_w) % w.tiles_h;
    const int n = tile / (w.tiles_w * w.tiles_h);
    mbar_expect_tx(&full_bar[buf], kABytes);
    tma_load_4d(sB + buf * kABytes, &w.tmap_dz, &full_bar[buf], 0, tw * kTW, th * kTH, n);
  };

  float pv[kPatchIters];
  int tile = blockIdx.x;
  const int step = gridDim.x;
  if (tile < w.total_tiles) {
    if (tid == 0) {
      issue_dz(tile, 0);
      if (tile + step < w.total_tiles) issue_dz(tile + step, 1);
    }
    patch_issue(a, tile, tid, pv);
    patch_store(pv, sPatch, tid);
    if (tile + step < w.total_tiles) patch_issue(a, tile + step, tid, pv);
  }
  __syncthreads();

  const int row = tid & 127, half = tid >> 7;
  const int dy = row >> 4, dx = row & 15;
  const uint32_t sw = (uint32_t)(row & 7);
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
  int it = 0;
  for (; tile < w.total_tiles; tile += step, ++it) {
    // the previous tile's MMAs must be done before sA (and its dZ buffer) are overwritten
    if (it > 0) {
      mbar_wait(mma_bar, (uint32_t)((it - 1) & 1));
      tc_fence_after_sync();
      if (tid == 0 && tile + step < w.total_tiles) issue_dz(tile + step, (it + 1) & 1);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const int q = half * 12 + i;
      if (q < 21) {
        const int c = q / 7, kh = q - c * 7;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(
            sPatch + (c * kPatchH + 2 * dy + kh) * kPatchPitch + 2 * dx);
        const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
        *reinterpret_cast<uint4*>(sA + (q >> 3) * kABytes + row * 128 +
                                  ((((uint32_t)q & 7) ^ sw) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      mbar_wait(&full_bar[it & 1], (uint32_t)((it >> 1) & 1));
      const uint32_t sa = smem_u32(sA), sb = smem_u32(sB + (it & 1) * kABytes);
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // 16 pixels per MMA
        const uint64_t bdesc = umma_desc_sw128(sb + k * 2048, 16, 1024);
        umma_f16(tmem, umma_desc_sw128(sa + k * 2048, kABytes, 1024), bdesc, idesc, (it | k) != 0);
        umma_f16(tmem + 64, umma_desc_sw128(sa + 2 * kABytes + k * 2048, kABytes, 1024), bdesc,
                 idesc, (it | k) != 0);
      }
      umma_commit(mma_bar);
    }
    const int next = tile + step;
    if (next < w.total_tiles) patch_store(pv, sPatch, tid);
    if (next + step < w.total_tiles) patch_issue(a, next + step, tid, pv);
    __syncthreads();
  }
  if (it > 0) {
    mbar_wait(mma_bar, (uint32_t)((it - 1) & 1));
    tc_fence_after_sync();
    const int quarter = warp & 3, colhalf = warp >> 2;
    const int erow = quarter * 32 + lane;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + mt * 64 + colhalf * 32, v);
      tmem_ld_wait();
      const int col = mt * 128 + erow;
      if (col < kKPad) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          atomicAdd(w.scratch + col * kCout + colhalf * 32 + j, __uint_as_float(v[j]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem, 128); }
}

// dw[k][c][kh][kw] (fp32 OIHW) = scratch[c*56 + kh*8 + kw][k]
__global__ void stem_wgrad_unpack_kernel(const float* __restrict__ scratch, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 64*3*7*7
  if (i >= kCout * 147) return;
  const int kw = i % 7, kh = (i / 7) % 7, c = (i / 49) % 3, k = i / 147;
  dw[i] = scratch[(c * 56 + kh * 8 + kw) * kCout + k];
}

__global__ void stem_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 64 * 192
  if (i >= kCout * kKPad) return;
  const int n = i / kKPad, k = i % kKPad;
  float v = 0.f;
  if (k < 168) {
    const int c = k / 56, kh = (k % 56) / 8, kw = k % 8;
    if (kw < 7) v = __ldg(w + ((n * 3 + c) * 7 + kh) * 7 + kw);
  }
  out[i] = __float2bfloat16_rn(v);
}

}  // namespace
}  // namespace rmv

using namespace rmv;

extern "C" int rmv_stem_pack_weights(const float* w_oihw, void* packed, void* stream) {
  RMV_CHECK_ARG(w_oihw && packed, "stem_pack_weights: null pointer");
  stem_pack_kernel<<<(kCout * kKPad + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, (__nv_bfloat16*)packed);
  RMV_LAUNCH_CHECK();
  return 0;
}

extern "C" int rmv_stem_conv_fwd(const float* x_nchw, const void* w_packed, const float* scale,
                                 const float* shift, void* y_nhwc, int n_img, int in_h, int in_w,
                                 int relu, void* stream) {
  RMV_CHECK_ARG(x_nchw && w_packed && y_nhwc, "stem_conv_fwd: null pointer");
  RMV_CHECK_ARG(in_h >= 7 && in_w >= 7, "stem_conv_fwd: input %dx%d too small", in_h, in_w);
  RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_nhwc) & 15) == 0,
                "stem_conv_fwd: pointers must be 16-byte aligned");
  if (n_img == 0) return 0;
  StemArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x_nchw; a.scale = scale; a.shift = shift;
  a.n_img = n_img; a.in_h = in_h; a.in_w = in_w; a.relu = relu;
  a.out_h = (in_h + 6 - 7) / 2 + 1;
  a.out_w = (in_w + 6 - 7) / 2 + 1;
  a.tiles_w = ceil_div(a.out_w, kTW);
  a.tiles_h = ceil_div(a.out_h, kTH);
  a.total_tiles = a.tiles_w * a.tiles_h * n_img;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kKPad, (cuuint64_t)kCout};
    cuuint64_t strides[1] = {(cuuint64_t)(kKPad * 2)};
    cuuint32_t box[2] = {64, (cuuint32_t)kCout};
    int rc = encode_map(&a.tmap_w, w_packed, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)kCout, (cuuint64_t)a.out_w, (cuuint64_t)a.out_h,
                          (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)(kCout * 2), (cuuint64_t)a.out_w * kCout * 2,
                             (cuuint64_t)a.out_h * a.out_w * kCout * 2};
    cuuint32_t box[4] = {(cuuint32_t)kCout, (cuuint32_t)kTW, (cuuint32_t)kTH, 1};
    int rc = encode_map(&a.tmap_out, y_nhwc, 4, dims, strides, box);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    RMV_CUDA(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmemBytes));
    attr_set = true;
  }
  const int grid = a.total_tiles < 2 * num_sms() ? a.total_tiles : 2 * num_sms();
  RMV_CUDA(launch_pdl_tc(stem_kernel, dim3(grid), dim3(kThreads), kSmemBytes, (cudaStream_t)stream, a));
  return 0;
}

extern "C" int rmv_stem_wgrad(const float* x_nchw, const void* dz_nhwc, float* scratch, float* dw_oihw,
                              int n_img, int in_h, int in_w, void* stream) {
  RMV_CHECK_ARG(x_nchw && dz_nhwc && scratch && dw_oihw, "stem_wgrad: null pointer");
  RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(dz_nhwc) & 15) == 0, "stem_wgrad: dz must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  RMV_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * kKPad * kCout, s));
  if (n_img > 0) {
    StemWgArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x_nchw; a.scratch = scratch;
    a.n_img = n_img; a.in_h = in_h; a.in_w = in_w;
    a.out_h = (in_h + 6 - 7) / 2 + 1;
    a.out_w = (in_w + 6 - 7) / 2 + 1;
    a.tiles_w = ceil_div(a.out_w, kTW);
    a.tiles_h = ceil_div(a.out_h, kTH);
    a.total_tiles = a.tiles_w * a.tiles_h * n_img;
    cuuint64_t dims[4] = {(cuuint64_t)kCout, (cuuint64_t)a.out_w, (cuuint64_t)a.out_h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)(kCout * 2), (cuuint64_t)a.out_w * kCout * 2,
                             (cuuint64_t)a.out_h * a.out_w * kCout * 2};
    cuuint32_t box[4] = {(cuuint32_t)kCout, (cuuint32_t)kTW, (cuuint32_t)kTH, 1};
    if (int rc = encode_map(&a.tmap_dz, dz_nhwc, 4, dims, strides, box)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
      RMV_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kWgSmemBytes));
      attr_set = true;
    }
    const int grid = a.total_tiles < 2 * num_sms() ? a.total_tiles : 2 * num_sms();
    stem_wgrad_kernel<<<grid, kThreads, kWgSmemBytes, s>>>(a);
    RMV_LAUNCH_CHECK();
  }
  stem_wgrad_unpack_kernel<<<(kCout * 147 + 255) / 256, 256, 0, s>>>(scratch, dw_oihw);
  RMV_LAUNCH_CHECK();
  return 0;
}
