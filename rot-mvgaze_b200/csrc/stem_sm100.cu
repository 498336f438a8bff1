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
  int vec_ok;  // forward loader may use aligned 16-byte loads (in_w % 4 == 0, 16-byte aligned x)
  // uint8 input (x_u8 != nullptr): [n, in_h, in_w, 3] bytes (HWC, as image decoders deliver them);
  // the loader applies ToTensor + Normalize (main.py:38-56): v = byte * u8_scale[c] + u8_offset[c]
  // with u8_scale = 1 / (255 std), u8_offset = -mean / std. Needs 3*in_w % 16 == 0, aligned base.
  const unsigned char* x_u8;
  float u8_scale[3], u8_offset[3];
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
template <int NT>
__device__ __forceinline__ void patch_issue(const StemArgs& a, int tile, int tid,
                                            float (&v)[(kPatchElems + NT - 1) / NT]) {
  const int tw = tile % a.tiles_w;
  const int th = (tile / a.tiles_w) % a.tiles_h;
  const int n = tile / (a.tiles_w * a.tiles_h);
  const int ih0 = 2 * th * kTH - 3, iw0 = 2 * tw * kTW - 3;
  const float* xn = a.x + (long long)n * 3 * a.in_h * a.in_w;
#pragma unroll
  for (int it = 0; it < (kPatchElems + NT - 1) / NT; ++it) {
    const int i = tid + it * NT;
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
template <int NT>
__device__ __forceinline__ void patch_store(const float (&v)[(kPatchElems + NT - 1) / NT],
                                            __nv_bfloat16* patch, int tid) {
#pragma unroll
  for (int it = 0; it < (kPatchElems + NT - 1) / NT; ++it) {
    const int i = tid + it * NT;
    if (i < kPatchElems) patch[i] = __float2bfloat16_rn(v[it]);
  }
}

// Warp-specialised forward kernel, one CTA per SM:
//   warps 0-3   loader:   fp32 NCHW patch -> registers (two tiles in flight) -> bf16 patch ring
//   warps 4-7   builder:  patch -> im2col rows in the swizzled UMMA layout (double-buffered)
//   warp  8     MMA:      12 tcgen05.mma per tile into one of two TMEM accumulators
//   warps 9-12  epilogue: TMEM -> scale/shift/ReLU -> bf16 -> swizzled staging -> TMA store
// every hand-over is an mbarrier, so the four stages of consecutive tiles overlap.
constexpr int kFwdThreads = 416;
constexpr int kPatchSlots = 3;
constexpr int kPatchBytes = 6144;
constexpr int kLdIters = (kPatchElems + 127) / 128;  // 24
constexpr int kFwdSmemBytes = 3 * kBBytes + 2 * 3 * kABytes + 2 * kABytes /*out staging*/ +
                              kPatchSlots * kPatchBytes + 512 /*scale,shift*/ + 256 /*barriers*/ +
                              1024 /*align*/;

__global__ void __launch_bounds__(kFwdThreads, 1) stem_kernel(const __grid_constant__ StemArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by an integer round trip of the pointer) so the compiler
  // keeps the shared address space and emits LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;
  uint8_t* sA = sB + 3 * kBBytes;        // [2][3 k-blocks][128 rows][128 B]
  uint8_t* sOut = sA + 2 * 3 * kABytes;  // [2][128 rows][128 B]
  uint8_t* sPatchBase = sOut + 2 * kABytes;
  float* s_scale = reinterpret_cast<float*>(sPatchBase + kPatchSlots * kPatchBytes);
  float* s_shift = s_scale + kCout;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + kCout);
  uint64_t* w_bar = bars;
  uint64_t* patch_full = bars + 1;                 // [3] loader  -> builder
  uint64_t* patch_empty = patch_full + kPatchSlots;  // [3] builder -> loader
  uint64_t* a_full = patch_empty + kPatchSlots;    // [2] builder -> MMA
  uint64_t* a_empty = a_full + 2;                  // [2] MMA     -> builder
  uint64_t* tmem_full = a_empty + 2;               // [2] MMA     -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;            // [2] epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tma_prefetch_desc(&a.tmap_w);
    tma_prefetch_desc(&a.tmap_out);
    mbar_init(w_bar, 1);
    for (int i = 0; i < kPatchSlots; ++i) { mbar_init(&patch_full[i], 128); mbar_init(&patch_empty[i], 128); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1);
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr, 128);
    tmem_relinquish();
  }
  // zero K-padding columns 168..191 (16-byte units 5,6,7 of k-block 2) of both A buffers: never change
  for (int i = tid; i < 2 * 128 * 3; i += kFwdThreads) {
    const int buf = i / (128 * 3), row = (i / 3) % 128;
    const uint32_t j = 5 + i % 3;
    *reinterpret_cast<uint4*>(sA + buf * 3 * kABytes + 2 * kABytes + row * 128 +
                              ((j ^ (uint32_t)(row & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  griddep_wait();    // PDL: global memory is touched from here on
  griddep_launch();
  const int step = gridDim.x;

  if (warp < 4) {
    // ------------------------------- loader ---------------------------------------------------
    // A single warp issues only ~0.5 instructions per clock, so the per-element work must be tiny:
    // the patch is fetched as ALIGNED 16-byte chunks (11 per row: global columns 32*tw-4 ...
    // 32*tw+39, one to the left of the patch origin 32*tw-3), whose (row, chunk) -> offset tables
    // do not depend on the tile and live in registers. Needs in_w % 4 == 0 and a 16-byte aligned
    // image base (the 224x224 case); otherwise the element-wise path below is used.
    if (a.x_u8 != nullptr) {
      // uint8 HWC: one image row holds all three channels of the 38 patch pixels in 114 contiguous
      // bytes starting at byte 96*tw - 9 of the row; fetched as 8 aligned 16-byte chunks from byte
      // 96*tw - 16, normalised, converted to bf16 and scattered into the planar patch.
      constexpr int kChunksRow = 8, kChunks = kPatchH * kChunksRow;  // 168
      constexpr int kIt = (kChunks + 127) / 128;                     // 2
      const int row_bytes = 3 * a.in_w;
      uint4 q0[kIt], q1[kIt];
      auto issue = [&](int tile, uint4 (&q)[kIt]) {
        const int tw = tile % a.tiles_w;
        const int th = (tile / a.tiles_w) % a.tiles_h;
        const int n = tile / (a.tiles_w * a.tiles_h);
        const int ih0 = 2 * th * kTH - 3;
        const unsigned char* img = a.x_u8 + (long long)n * a.in_h * row_bytes;
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          const int k = tid + it * 128;
          const int r = k / kChunksRow, j = k - r * kChunksRow;
          const int ih = ih0 + r, b0 = 96 * tw - 16 + 16 * j;   // first byte of the chunk in its row
          q[it] = (k < kChunks && ih >= 0 && ih < a.in_h && b0 >= 0 && b0 < row_bytes)
                      ? __ldg(reinterpret_cast<const uint4*>(img + (long long)ih * row_bytes + b0))
                      : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
      };
      auto park = [&](int tile, const uint4 (&q)[kIt], __nv_bfloat16* patch) {
        const int tw = tile % a.tiles_w;
        const int th = (tile / a.tiles_w) % a.tiles_h;
        const int ih0 = 2 * th * kTH - 3;
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          const int k = tid + it * 128;
          if (k >= kChunks) continue;
          const int r = k / kChunksRow, j = k - r * kChunksRow;
          const int ih = ih0 + r, b0 = 96 * tw - 16 + 16 * j;
          const bool in = ih >= 0 && ih < a.in_h && b0 >= 0 && b0 < row_bytes;  // else: conv zero padding
          const uint32_t w4[4] = {q[it].x, q[it].y, q[it].z, q[it].w};
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            // byte position relative to the first byte of patch column 0 (= byte 96*tw - 9)
            const int pos = 16 * j - 7 + e;
            if (pos < 0 || pos >= 3 * kPatchWUsed) continue;
            const int col = pos / 3, c = pos - 3 * col;
            const float byte = (float)((w4[e >> 2] >> (8 * (e & 3))) & 0xFFu);
            const float v = in ? fmaf(byte, a.u8_scale[c], a.u8_offset[c]) : 0.f;
            patch[(c * kPatchH + r) * kPatchPitch + col] = __float2bfloat16_rn(v);
          }
        }
      };
      int tile = blockIdx.x;
      if (tile < a.total_tiles) issue(tile, q0);
      if (tile + step < a.total_tiles) issue(tile + step, q1);
      for (int i = 0; tile < a.total_tiles; i += 2) {
        {
          const int slot = i % kPatchSlots;
          mbar_wait(&patch_empty[slot], ((i / kPatchSlots) & 1) ^ 1);
          park(tile, q0, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes));
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) issue(tile + 2 * step, q0);
          tile += step;
        }
        if (tile < a.total_tiles) {
          const int slot = (i + 1) % kPatchSlots;
          mbar_wait(&patch_empty[slot], (((i + 1) / kPatchSlots) & 1) ^ 1);
          park(tile, q1, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes));
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) issue(tile + 2 * step, q1);
          tile += step;
        }
      }
    } else if (a.vec_ok) {
      constexpr int kChunksRow = 11, kChunks = 3 * kPatchH * kChunksRow;  // 693
      constexpr int kIt = (kChunks + 127) / 128;                          // 6
      int goff[kIt], poff[kIt], rr[kIt], jj[kIt];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int k = tid + it * 128;
        const int row = k / kChunksRow, j = k - row * kChunksRow;  // row = c*21 + r
        const int c = row / kPatchH, r = row - c * kPatchH;
        rr[it] = (k < kChunks) ? r : -100000;                      // invalid chunk: never in range
        jj[it] = 4 * j - 4;                                        // column relative to 32*tw
        goff[it] = (c * a.in_h + r) * a.in_w + 4 * j - 4;
        poff[it] = row * kPatchPitch + 4 * j - 1;                  // patch column of element 0
      }
      float4 q0[kIt], q1[kIt];
      auto issue = [&](int tile, float4 (&q)[kIt]) {
        const int tw = tile % a.tiles_w;
        const int th = (tile / a.tiles_w) % a.tiles_h;
        const int n = tile / (a.tiles_w * a.tiles_h);
        const int ih0 = 2 * th * kTH - 3, gw0 = 2 * tw * kTW;
        const float* org = a.x + ((long long)n * 3 * a.in_h + ih0) * a.in_w + gw0;
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          const int ih = ih0 + rr[it], gw = gw0 + jj[it];
          q[it] = (ih >= 0 && ih < a.in_h && gw >= 0 && gw < a.in_w)
                      ? __ldg(reinterpret_cast<const float4*>(org + goff[it]))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto park = [&](const float4 (&q)[kIt], __nv_bfloat16* patch) {
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
          if (rr[it] >= 0) {
            __nv_bfloat16* d = patch + poff[it];
            if (jj[it] >= 0) d[0] = __float2bfloat16_rn(q[it].x);  // chunk 0 starts left of the patch
            d[1] = __float2bfloat16_rn(q[it].y);
            d[2] = __float2bfloat16_rn(q[it].z);
            d[3] = __float2bfloat16_rn(q[it].w);
          }
        }
      };
      int tile = blockIdx.x;
      if (tile < a.total_tiles) issue(tile, q0);
      if (tile + step < a.total_tiles) issue(tile + step, q1);
      for (int i = 0; tile < a.total_tiles; i += 2) {
        {
          const int slot = i % kPatchSlots;
          mbar_wait(&patch_empty[slot], ((i / kPatchSlots) & 1) ^ 1);
          park(q0, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes));
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) issue(tile + 2 * step, q0);
          tile += step;
        }
        if (tile < a.total_tiles) {
          const int slot = (i + 1) % kPatchSlots;
          mbar_wait(&patch_empty[slot], (((i + 1) / kPatchSlots) & 1) ^ 1);
          park(q1, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes));
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) issue(tile + 2 * step, q1);
          tile += step;
        }
      }
    } else {
      float p0[kLdIters], p1[kLdIters];
      int tile = blockIdx.x;
      if (tile < a.total_tiles) patch_issue<128>(a, tile, tid, p0);
      if (tile + step < a.total_tiles) patch_issue<128>(a, tile + step, tid, p1);
      for (int i = 0; tile < a.total_tiles; i += 2) {
        {
          const int slot = i % kPatchSlots;
          mbar_wait(&patch_empty[slot], ((i / kPatchSlots) & 1) ^ 1);
          patch_store<128>(p0, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes), tid);
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) patch_issue<128>(a, tile + 2 * step, tid, p0);
          tile += step;
        }
        if (tile < a.total_tiles) {
          const int slot = (i + 1) % kPatchSlots;
          mbar_wait(&patch_empty[slot], (((i + 1) / kPatchSlots) & 1) ^ 1);
          patch_store<128>(p1, reinterpret_cast<__nv_bfloat16*>(sPatchBase + slot * kPatchBytes), tid);
          mbar_arrive(&patch_full[slot]);
          if (tile + 2 * step < a.total_tiles) patch_issue<128>(a, tile + 2 * step, tid, p1);
          tile += step;
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------- im2col builder -------------------------------------------
    const int row = tid - 128;
    const int dy = row >> 4, dx = row & 15;
    const uint32_t sw = (uint32_t)(row & 7);
    int i = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += step, ++i) {
      const int slot = i % kPatchSlots, ab = i & 1;
      const __nv_bfloat16* sPatch =
          reinterpret_cast<const __nv_bfloat16*>(sPatchBase + slot * kPatchBytes);
      uint8_t* dst = sA + ab * 3 * kABytes;
      mbar_wait(&patch_full[slot], (i / kPatchSlots) & 1);
      mbar_wait(&a_empty[ab], ((i >> 1) & 1) ^ 1);
      // 16-byte unit q = (c, kh): 8 consecutive kw taps of input row 2*dy+kh, channel c
#pragma unroll
      for (int q = 0; q < 21; ++q) {
        const int c = q / 7, kh = q - c * 7;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(
            sPatch + (c * kPatchH + 2 * dy + kh) * kPatchPitch + 2 * dx);
        const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
        *reinterpret_cast<uint4*>(dst + (q >> 3) * kABytes + row * 128 +
                                  ((((uint32_t)q & 7) ^ sw) << 4)) = v;
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[ab]);
      mbar_arrive(&patch_empty[slot]);
    }
  } else if (warp == 8) {
    // ------------------------------- MMA issuer -----------------------------------------------
    if (lane == 0) {
      mbar_expect_tx(w_bar, 3 * kBBytes);
      for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * kBBytes, &a.tmap_w, w_bar, kb * 64, 0);
    }
    constexpr uint32_t idesc = umma_idesc_bf16(128, kCout, 0, 0);
    mbar_wait(w_bar, 0);
    int i = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += step, ++i) {
      const int ab = i & 1;
      const uint32_t par = (i >> 1) & 1;
      mbar_wait(&a_full[ab], par);
      mbar_wait(&tmem_empty[ab], par ^ 1);
      tc_fence_after_sync();
      if (lane == 0) {
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + (ab * 3 + kb) * kABytes), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + kb * kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem + ab * kCout, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&a_empty[ab]);
        umma_commit(&tmem_full[ab]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- epilogue (warps 9..12) -----------------------------------
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int erow = quarter * 32 + lane;
    const uint32_t esw = (uint32_t)(erow & 7);
    const int tid_e = tid - 9 * 32;
    if (tid_e < kCout) {
      s_scale[tid_e] = a.scale ? __ldg(a.scale + tid_e) : 1.f;
      s_shift[tid_e] = a.shift ? __ldg(a.shift + tid_e) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    int i = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += step, ++i) {
      const int tb = i & 1;
      uint8_t* obuf = sOut + tb * kABytes;
      mbar_wait(&tmem_full[tb], (i >> 1) & 1);
      tc_fence_after_sync();
      uint32_t v[64];
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + tb * kCout;
      tmem_ld_32x32b_x32(taddr, v);
      tmem_ld_32x32b_x32(taddr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[tb]);
      // staging buffer tb was handed to the TMA store two tiles ago: make sure it has been read
      if (tid_e == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float f[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + q * 8 + h * 4);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + q * 8 + h * 4);
          f[h * 4 + 0] = fmaf(__uint_as_float(v[q * 8 + h * 4 + 0]), sc.x, sh.x);
          f[h * 4 + 1] = fmaf(__uint_as_float(v[q * 8 + h * 4 + 1]), sc.y, sh.y);
          f[h * 4 + 2] = fmaf(__uint_as_float(v[q * 8 + h * 4 + 2]), sc.z, sh.z);
          f[h * 4 + 3] = fmaf(__uint_as_float(v[q * 8 + h * 4 + 3]), sc.w, sh.w);
        }
        if (a.relu) {
#pragma unroll
          for (int t = 0; t < 8; ++t) f[t] = fmaxf(f[t], 0.f);
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(obuf + erow * 128 + (((uint32_t)q ^ esw) << 4)) = o;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid_e == 0) {
        const int tw = tile % a.tiles_w;
        const int th = (tile / a.tiles_w) % a.tiles_h;
        const int n = tile / (a.tiles_w * a.tiles_h);
        stem_tma_store(&a.tmap_out, obuf, 0, tw * kTW, th * kTH, n);
      }
    }
    if (tid_e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 128);
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
    patch_issue<kThreads>(a, tile, tid, pv);
    patch_store<kThreads>(pv, sPatch, tid);
    if (tile + step < w.total_tiles) patch_issue<kThreads>(a, tile + step, tid, pv);
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
    if (next < w.total_tiles) patch_store<kThreads>(pv, sPatch, tid);
    if (next + step < w.total_tiles) patch_issue<kThreads>(a, next + step, tid, pv);
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

static int stem_fwd_launch(const float* x_nchw, const unsigned char* x_u8, const float* mean,
                           const float* stdv, const void* w_packed, const float* scale,
                           const float* shift, void* y_nhwc, int n_img, int in_h, int in_w,
                           int relu, void* stream);

extern "C" int rmv_stem_conv_fwd(const float* x_nchw, const void* w_packed, const float* scale,
                                 const float* shift, void* y_nhwc, int n_img, int in_h, int in_w,
                                 int relu, void* stream) {
  RMV_CHECK_ARG(x_nchw != nullptr, "stem_conv_fwd: null pointer");
  return stem_fwd_launch(x_nchw, nullptr, nullptr, nullptr, w_packed, scale, shift, y_nhwc, n_img,
                         in_h, in_w, relu, stream);
}

extern "C" int rmv_stem_conv_fwd_u8(const unsigned char* x_nhwc_u8, const float* mean3,
                                    const float* std3, const void* w_packed, const float* scale,
                                    const float* shift, void* y_nhwc, int n_img, int in_h, int in_w,
                                    int relu, void* stream) {
  RMV_CHECK_ARG(x_nhwc_u8 && mean3 && std3, "stem_conv_fwd_u8: null pointer");
  RMV_CHECK_ARG((3 * in_w) % 16 == 0 && (reinterpret_cast<uintptr_t>(x_nhwc_u8) & 15) == 0,
                "stem_conv_fwd_u8: 3*in_w must be a multiple of 16 bytes and x 16-byte aligned");
  for (int c = 0; c < 3; ++c)
    RMV_CHECK_ARG(std3[c] > 0.f, "stem_conv_fwd_u8: std must be positive");
  return stem_fwd_launch(nullptr, x_nhwc_u8, mean3, std3, w_packed, scale, shift, y_nhwc, n_img,
                         in_h, in_w, relu, stream);
}

static int stem_fwd_launch(const float* x_nchw, const unsigned char* x_u8, const float* mean,
                           const float* stdv, const void* w_packed, const float* scale,
                           const float* shift, void* y_nhwc, int n_img, int in_h, int in_w,
                           int relu, void* stream) {
  RMV_CHECK_ARG((x_nchw || x_u8) && w_packed && y_nhwc, "stem_conv_fwd: null pointer");
  RMV_CHECK_ARG(in_h >= 7 && in_w >= 7, "stem_conv_fwd: input %dx%d too small", in_h, in_w);
  RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_nhwc) & 15) == 0,
                "stem_conv_fwd: pointers must be 16-byte aligned");
  if (n_img == 0) return 0;
  StemArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x_nchw; a.scale = scale; a.shift = shift;
  a.n_img = n_img; a.in_h = in_h; a.in_w = in_w; a.relu = relu;
  a.vec_ok = (in_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0);
  a.x_u8 = x_u8;
  if (x_u8 != nullptr) {
    for (int c = 0; c < 3; ++c) {
      a.u8_scale[c] = 1.f / (255.f * stdv[c]);
      a.u8_offset[c] = -mean[c] / stdv[c];
    }
  }
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
                                  kFwdSmemBytes));
    attr_set = true;
  }
  const int grid = a.total_tiles < num_sms() ? a.total_tiles : num_sms();
  RMV_CUDA(launch_pdl_tc(stem_kernel, dim3(grid), dim3(kFwdThreads), kFwdSmemBytes,
                         (cudaStream_t)stream, a));
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
