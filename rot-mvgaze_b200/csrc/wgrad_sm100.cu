// Filter gradient of a convolution on the tensor cores (sm_100a):
//
//   dW[k, tap, c] = sum over output pixels p of dY[p, k] * X[p shifted by tap, c]
//
// is a GEMM whose reduction axis is the PIXEL axis. Both operands are read exactly as the forward
// kernel reads its activations -- 4-D tiled TMA boxes of (64 channels, box_w, box_h, box_n) pixels,
// tap shift and zero padding done by the TMA unit, stride-2 layers through parity planes -- and fed
// to tcgen05.mma as MN-MAJOR operands (channels contiguous, pixels along K), so no transposed copy
// of any activation is ever made. Work item = (k_out tile of 128) x (tap) x (c tile) x (pixel
// split); each CTA accumulates its pixel range in TMEM (fp32) and adds its partial tile into an
// fp32 [k_out][tap*C] scratch with a TMA reduce-add store (split-K without atomics in the kernel).
//
// Replaces the autograd weight-gradient of nn.Conv2d / nn.Linear (trainer.py:142).
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace rmv {
namespace {

constexpr int kPix = 128;      // pixels per k-block (one TMA box of box_w*box_h*box_n pixels)
constexpr int kBoxBytes = kPix * 128;  // [128 pixels][64 channels] bf16
constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int kMaxTaps = 49;

struct WgArgs {
  CUtensorMap tmap_x[4];   // input parity planes (64 ch, box_w, box_h, box_n)
  CUtensorMap tmap_dy;     // output gradient (64 ch, box_w, box_h, box_n)
  CUtensorMap tmap_dw;     // fp32 scratch [k_out][taps*C], box (32 cols, 128 rows)
  int box_w, box_h, box_n, tiles_w, tiles_h, tiles_n;
  int m_blocks;            // pixel blocks in total
  int blocks_per_split, splits;
  int k_tiles, c_tiles, num_taps, c_in;
  signed char tap_map[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
};

// MN-major, 128-byte swizzle: 64 channels (128 B) contiguous per pixel row, 8-pixel groups 1024 B
// apart (SBO), 64-channel blocks `lbo` bytes apart (LBO).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo) {
  return umma_desc_sw128(addr, lbo, 1024);
}

__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, const void* src, int c0, int c1) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1)
      : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int N_TILE>  // channels of X per tile: 64 or 128
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgArgs a) {
  constexpr int kNB = N_TILE / 64;
  constexpr int kStageBytes = 2 * kBoxBytes + kNB * kBoxBytes;
  constexpr int kStages = (N_TILE == 128) ? 3 : 4;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by an integer round trip of the pointer) so the compiler
  // keeps the shared address space and emits LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* s_out = smem + kStages * kStageBytes;  // [128 rows][32 fp32], SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + kBoxBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* done_bar = bars + 2 * kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmap_dy);
    tma_prefetch_desc(&a.tmap_x[0]);
    tma_prefetch_desc(&a.tmap_dw);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, N_TILE < 32 ? 32 : N_TILE); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  griddep_wait();    // PDL: the prologue above overlapped the previous kernel's tail
  griddep_launch();

  // work item: tap fastest so the CTAs sharing a pixel range run together (L2 reuse of dY and X)
  int w = blockIdx.x;
  const int tap = w % a.num_taps; w /= a.num_taps;
  const int ct = w % a.c_tiles; w /= a.c_tiles;
  const int kt = w % a.k_tiles; w /= a.k_tiles;
  const int split = w;
  const int pb0 = split * a.blocks_per_split;
  const int pb1 = min(a.m_blocks, pb0 + a.blocks_per_split);
  const int n_blocks = pb1 - pb0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pb = pb0; pb < pb1; ++pb) {
        const int tw = pb % a.tiles_w;
        const int th = (pb / a.tiles_w) % a.tiles_h;
        const int tn = pb / (a.tiles_w * a.tiles_h);
        const int ow0 = tw * a.box_w, oh0 = th * a.box_h, n0 = tn * a.box_n;
        uint8_t* sa = stage_base + stage * kStageBytes;
        uint8_t* sb = sa + 2 * kBoxBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], kStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_4d(sa + j * kBoxBytes, &a.tmap_dy, &full_bar[stage], kt * 128 + j * 64, ow0, oh0, n0);
#pragma unroll
        for (int j = 0; j < kNB; ++j)
          tma_load_4d(sb + j * kBoxBytes, &a.tmap_x[a.tap_map[tap]], &full_bar[stage],
                      ct * N_TILE + j * 64, ow0 + a.tap_dw[tap], oh0 + a.tap_dh[tap], n0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N_TILE, 1, 1);  // both operands MN-major
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < n_blocks; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after_sync();
      if (lane == 0) {
        const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
        const uint32_t sb = sa + 2 * kBoxBytes;
#pragma unroll
        for (int k = 0; k < kPix / 16; ++k) {
          // 16 pixels per MMA = two 8-pixel swizzle atoms = 2048 bytes further down the box
          const uint64_t adesc = desc_mn_sw128(sa + k * 2048, kBoxBytes);
          const uint64_t bdesc = desc_mn_sw128(sb + k * 2048, kBoxBytes);
          umma_f16(tmem, adesc, bdesc, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (i == n_blocks - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (n_blocks > 0) {
    // epilogue: TMEM -> swizzled fp32 staging -> TMA reduce-add into the scratch
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const int tid_e = threadIdx.x - 64;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < N_TILE / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (tid_e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(s_out + row * 128 + (((uint32_t)q ^ sw) << 4)) =
            make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid_e == 0)
        tma_reduce_add_2d(&a.tmap_dw, s_out, tap * a.c_in + ct * N_TILE + c * 32, kt * 128);
    }
    if (tid_e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem, N_TILE < 32 ? 32 : N_TILE); }
}


// ---------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 / 64 -> 64 channels (layer1 conv2: the largest pixel count, the smallest
// GEMM): tap-by-tap the kernel above re-reads dY and X nine times and issues N = 64 MMAs. Here one
// CTA owns a TAP ROW r: per 8x16-pixel tile it loads dY once (16 KB) and ONE (16 wide x 16 high) box
// of X starting at (ow0-1, oh0+r-1), and the three taps (r, 0..2) are the three 64-channel N-blocks
// of a single N = 192 MN-major B operand whose blocks lie 128 B (one pixel) apart -- LBO = 128,
// SBO = 2048 (one 16-pixel patch row per 8-pixel group). 3x fewer, 3x wider MMAs, 9x -> 3x dY
// traffic, 9x -> 6x smaller X traffic. The upper 64 rows of the M = 128 accumulator are fed from a
// shared-memory block of zeros (c_out = 64).
// ---------------------------------------------------------------------------------------------
struct WgRowsArgs {
  CUtensorMap tmap_x;    // (64 ch, in_w, in_h, n), box (64, 16, 16, 1)
  CUtensorMap tmap_dy;   // (64 ch, out_w, out_h, n), box (64, 8, 16, 1)
  CUtensorMap tmap_dw;   // fp32 scratch [64][9*64], box (32 cols, 128 rows)
  int tiles_w, tiles_h, m_blocks, blocks_per_split;
};
constexpr int kRowsStages = 3;
constexpr int kRowsStageBytes = 2 * kBoxBytes /*dY + zeros*/ + 2 * kBoxBytes /*X patch 16x16*/;
constexpr int kRowsSmem = kRowsStages * kRowsStageBytes + kBoxBytes + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_rows_kernel(const __grid_constant__ WgRowsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_out = smem + kRowsStages * kRowsStageBytes;  // [128 rows][32 fp32], SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + kBoxBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kRowsStages;
  uint64_t* done_bar = bars + 2 * kRowsStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmap_dy);
    tma_prefetch_desc(&a.tmap_x);
    tma_prefetch_desc(&a.tmap_dw);
    for (int s = 0; s < kRowsStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 256); tmem_relinquish(); }
  // the zero half of every stage's A operand (channels 64..127 of dY do not exist)
  for (int i = threadIdx.x; i < kRowsStages * (kBoxBytes / 16); i += kThreads) {
    const int st = i / (kBoxBytes / 16), j = i % (kBoxBytes / 16);
    *reinterpret_cast<uint4*>(smem + st * kRowsStageBytes + kBoxBytes + j * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  griddep_wait();
  griddep_launch();

  const int r = blockIdx.x % 3, split = blockIdx.x / 3;
  const int pb0 = split * a.blocks_per_split;
  const int pb1 = min(a.m_blocks, pb0 + a.blocks_per_split);
  const int n_blocks = pb1 - pb0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pb = pb0; pb < pb1; ++pb) {
        const int tw = pb % a.tiles_w;
        const int th = (pb / a.tiles_w) % a.tiles_h;
        const int n = pb / (a.tiles_w * a.tiles_h);
        uint8_t* sa = smem + stage * kRowsStageBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], 3 * kBoxBytes);
        tma_load_4d(sa, &a.tmap_dy, &full_bar[stage], 0, tw * 8, th * 16, n);
        tma_load_4d(sa + 2 * kBoxBytes, &a.tmap_x, &full_bar[stage], 0, tw * 8 - 1, th * 16 + r - 1, n);
        if (++stage == kRowsStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 1, 1);  // both operands MN-major
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < n_blocks; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after_sync();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + stage * kRowsStageBytes);
        const uint32_t sb = sa + 2 * kBoxBytes;
#pragma unroll
        for (int k = 0; k < kPix / 16; ++k) {
          // A: 16 pixels = two 8-pixel atoms of the 8-wide dY box (2048 B), channel block 1 = zeros
          const uint64_t adesc = umma_desc_sw128(sa + k * 2048, kBoxBytes, 1024);
          // B: 16 pixels = two patch rows (2 x 2048 B); N-block j = tap (r, j) = one pixel further
          const uint64_t bdesc = umma_desc_sw128(sb + k * 4096, 128, 2048);
          umma_f16(tmem, adesc, bdesc, idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (i == n_blocks - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == kRowsStages) { stage = 0; phase ^= 1; }
    }
  } else if (n_blocks > 0) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const int tid_e = threadIdx.x - 64;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < 192 / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(quarter * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (tid_e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(s_out + row * 128 + (((uint32_t)q ^ sw) << 4)) =
            make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      // columns of the scratch: tap (r, s) * 64 + c_in = r*192 + accumulator column
      if (tid_e == 0) tma_reduce_add_2d(&a.tmap_dw, s_out, r * 192 + c * 32, 0);
    }
    if (tid_e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem, 256); }
}

int wgrad_rows_enabled() { return tuning("WGRAD_ROWS", 1, 1); }

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int N_TILE>
int launch_wg(const WgArgs& a, int grid, cudaStream_t stream) {
  constexpr int kNB = N_TILE / 64;
  constexpr int kStageBytes = 2 * kBoxBytes + kNB * kBoxBytes;
  constexpr int kStages = (N_TILE == 128) ? 3 : 4;
  constexpr int smem = kStages * kStageBytes + kBoxBytes + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    RMV_CUDA(cudaFuncSetAttribute(wgrad_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  RMV_CUDA(launch_pdl_tc(wgrad_kernel<N_TILE>, dim3(grid), dim3(kThreads), smem, stream, a));
  return 0;
}

}  // namespace

// dw_scratch: fp32 [c_out][kh*kw*c_in], accumulated (+=); the caller zeroes it.
int conv_wgrad_tc(const ConvArgs& p, const void* dy, float* dw_scratch, cudaStream_t stream) {
  RMV_CHECK_ARG(p.x_dtype == RMV_DTYPE_BF16, "tcgen05 wgrad: bf16 activations only");
  RMV_CHECK_ARG(p.c_in % 64 == 0 && p.c_out % 8 == 0, "tcgen05 wgrad: c_in %% 64 / c_out %% 8");
  RMV_CHECK_ARG(p.stride == 1 || p.stride == 2, "tcgen05 wgrad: stride %d unsupported", p.stride);
  RMV_CHECK_ARG(p.kh * p.kw <= kMaxTaps, "tcgen05 wgrad: filter too large");
  RMV_CHECK_ARG(p.x_sw % 8 == 0 && p.x_sh % 8 == 0 && p.x_sn % 8 == 0 && p.y_sw % 8 == 0 &&
                    p.y_sh % 8 == 0 && p.y_sn % 8 == 0,
                "tcgen05 wgrad: pixel strides must be multiples of 8 elements");
  if (wgrad_rows_enabled() && p.kh == 3 && p.kw == 3 && p.stride == 1 && p.pad == 1 && p.c_in == 64 &&
      p.c_out == 64 && p.out_w >= 8 && p.out_h >= 16) {
    WgRowsArgs r;
    memset(&r, 0, sizeof(r));
    r.tiles_w = ceil_div(p.out_w, 8);
    r.tiles_h = ceil_div(p.out_h, 16);
    r.m_blocks = r.tiles_w * r.tiles_h * p.n_img;
    if (r.m_blocks == 0) return 0;
    {
      cuuint64_t dims[4] = {64, (cuuint64_t)p.in_w, (cuuint64_t)p.in_h, (cuuint64_t)p.n_img};
      cuuint64_t strides[3] = {(cuuint64_t)(p.x_sw * 2), (cuuint64_t)(p.x_sh * 2), (cuuint64_t)(p.x_sn * 2)};
      cuuint32_t box[4] = {64, 16, 16, 1};
      if (int rc = encode_map(&r.tmap_x, p.x, 4, dims, strides, box)) return rc;
    }
    {
      cuuint64_t dims[4] = {64, (cuuint64_t)p.out_w, (cuuint64_t)p.out_h, (cuuint64_t)p.n_img};
      cuuint64_t strides[3] = {(cuuint64_t)(p.y_sw * 2), (cuuint64_t)(p.y_sh * 2), (cuuint64_t)(p.y_sn * 2)};
      cuuint32_t box[4] = {64, 8, 16, 1};
      if (int rc = encode_map(&r.tmap_dy, dy, 4, dims, strides, box)) return rc;
    }
    {
      cuuint64_t dims[2] = {9 * 64, 64};
      cuuint64_t strides[1] = {9 * 64 * 4};
      cuuint32_t obox[2] = {32, 128};
      if (int rc = encode_map(&r.tmap_dw, dw_scratch, 2, dims, strides, obox, true)) return rc;
    }
    long splits = num_sms() / 3;   // one wave: 3 tap rows x splits CTAs <= SMs (1 CTA per SM)
    if (splits < 1) splits = 1;
    if (splits > r.m_blocks) splits = r.m_blocks;
    r.blocks_per_split = ceil_div(r.m_blocks, splits);
    splits = ceil_div(r.m_blocks, r.blocks_per_split);
    static bool attr_set = false;
    if (!attr_set) {
      RMV_CUDA(cudaFuncSetAttribute(wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowsSmem));
      attr_set = true;
    }
    RMV_CUDA(launch_pdl_tc(wgrad_rows_kernel, dim3((unsigned)(3 * splits)), dim3(kThreads), kRowsSmem, stream, r));
    return 0;
  }
  WgArgs a;
  memset(&a, 0, sizeof(a));
  int out_w = p.out_w, out_h = p.out_h, n_img = p.n_img, in_w = p.in_w, in_h = p.in_h;
  long long x_sw = p.x_sw, x_sh = p.x_sh, x_sn = p.x_sn;
  long long y_sw = p.y_sw, y_sh = p.y_sh, y_sn = p.y_sn;
  const bool pointwise = (p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad == 0);
  const bool x_dense = (p.x_sh == p.x_sw * p.in_w) && (p.x_sn == p.x_sh * p.in_h);
  const bool y_dense = (p.y_sh == p.y_sw * p.out_w) && (p.y_sn == p.y_sh * p.out_h);
  if (pointwise && ((x_dense && y_dense) || (p.n_img == 1 && p.in_h == 1))) {
    out_w = in_w = p.n_img * p.in_h * p.in_w;
    out_h = in_h = 1; n_img = 1;
    x_sh = x_sw * in_w; x_sn = x_sh; y_sh = y_sw * out_w; y_sn = y_sh;
  }
  int best_w = 128, best_h = 1, best_n = 1;
  double best_eff = -1;
  for (int bw = 128; bw >= 1; bw >>= 1)
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bn = 128 / (bw * bh);
      const double eff = (double)out_w * out_h * n_img /
                         ((double)ceil_div(out_w, bw) * bw * ceil_div(out_h, bh) * bh * ceil_div(n_img, bn) * bn);
      if (eff > best_eff + 1e-9) { best_eff = eff; best_w = bw; best_h = bh; best_n = bn; }
    }
  a.box_w = best_w; a.box_h = best_h; a.box_n = best_n;
  a.tiles_w = ceil_div(out_w, best_w); a.tiles_h = ceil_div(out_h, best_h); a.tiles_n = ceil_div(n_img, best_n);
  a.m_blocks = a.tiles_w * a.tiles_h * a.tiles_n;
  if (a.m_blocks == 0) return 0;
  cuuint32_t box[4] = {64, (cuuint32_t)best_w, (cuuint32_t)best_h, (cuuint32_t)best_n};
  const int s = p.stride;
  int plane_id[2][2] = {{-1, -1}, {-1, -1}};
  int n_planes = 0;
  a.num_taps = p.kh * p.kw;
  for (int r = 0; r < p.kh; ++r)
    for (int q = 0; q < p.kw; ++q) {
      const int t = r * p.kw + q;
      const int ph = ((r - p.pad) % s + s) % s, pw = ((q - p.pad) % s + s) % s;
      if (plane_id[ph][pw] < 0) {
        const int pl_w = (in_w - pw + s - 1) / s, pl_h = (in_h - ph + s - 1) / s;
        cuuint64_t dims[4] = {(cuuint64_t)p.c_in, (cuuint64_t)pl_w, (cuuint64_t)pl_h, (cuuint64_t)n_img};
        cuuint64_t strides[3] = {(cuuint64_t)(x_sw * s * 2), (cuuint64_t)(x_sh * s * 2), (cuuint64_t)(x_sn * 2)};
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(p.x) + ph * x_sh + pw * x_sw;
        if (int rc = encode_map(&a.tmap_x[n_planes], base, 4, dims, strides, box)) return rc;
        plane_id[ph][pw] = n_planes++;
      }
      a.tap_map[t] = (signed char)plane_id[ph][pw];
      a.tap_dh[t] = (signed char)floordiv(r - p.pad, s);
      a.tap_dw[t] = (signed char)floordiv(q - p.pad, s);
    }
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.c_out, (cuuint64_t)out_w, (cuuint64_t)out_h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)(y_sw * 2), (cuuint64_t)(y_sh * 2), (cuuint64_t)(y_sn * 2)};
    if (int rc = encode_map(&a.tmap_dy, dy, 4, dims, strides, box)) return rc;
  }
  const long long k_cols = (long long)a.num_taps * p.c_in;
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_cols, (cuuint64_t)p.c_out};
    cuuint64_t strides[1] = {(cuuint64_t)(k_cols * 4)};
    cuuint32_t obox[2] = {32, 128};
    if (int rc = encode_map(&a.tmap_dw, dw_scratch, 2, dims, strides, obox, true)) return rc;
  }
  const int n_tile = (p.c_in % 128 == 0) ? 128 : 64;
  a.c_in = p.c_in;
  a.k_tiles = ceil_div(p.c_out, 128);
  a.c_tiles = p.c_in / n_tile;
  const long tiles = (long)a.k_tiles * a.c_tiles * a.num_taps;
  // pixel splits: one wave of one-CTA-per-SM blocks (WGRAD_WAVES = 1, default; measured 3.8 -> 3.3 ms
  // per step against two waves: half the split-K reduce traffic, no second-wave ramp)
  const int waves = tuning("WGRAD_WAVES", 1, 4);
  long splits = waves == 1 ? (num_sms() / tiles) : ((long)waves * num_sms() + tiles - 1) / tiles;
  if (splits > a.m_blocks) splits = a.m_blocks;
  if (splits < 1) splits = 1;
  a.blocks_per_split = ceil_div(a.m_blocks, splits);
  a.splits = ceil_div(a.m_blocks, a.blocks_per_split);
  const long grid = tiles * a.splits;
  RMV_CHECK_ARG(grid < (1L << 31), "tcgen05 wgrad: grid too large");
  return n_tile == 128 ? launch_wg<128>(a, (int)grid, stream) : launch_wg<64>(a, (int)grid, stream);
}

}  // namespace rmv

// dw (fp32, [c_out][kh][kw][c_in] = KRSC, +=): tensor-core weight gradient. The caller zeroes dw
// and converts to the parameter layout (rmv_permute_cast).
extern "C" int rmv_conv2d_wgrad_tc(const rmv_conv_args* args, const void* dy, float* dw_krsc,
                                   void* stream) {
  RMV_CHECK_ARG(args && dy && dw_krsc, "conv2d_wgrad_tc: null pointer");
  return rmv::conv_wgrad_tc(*args, dy, dw_krsc, (cudaStream_t)stream);
}
