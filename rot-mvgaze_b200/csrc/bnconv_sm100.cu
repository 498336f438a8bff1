// BatchNorm(train) reductions over a RECOMPUTED 1x1 convolution output (sm_100a).
//
// The expanding 1x1 convolutions of the ResNet bottlenecks (conv3, downsample; models/resnet.py:
// 122-123,229-230) produce the largest tensors of the training step, and their BatchNorm needs two
// grid-wide reductions over them: (sum z, sum z^2) in the forward pass and (sum dy, sum dy*xhat) in
// the backward pass. Reading z back from HBM for that costs four times the bytes of the conv's own
// input (c_out = 4 c_in). These kernels never materialise z: they recompute the GEMM on the (otherwise
// idle) tensor cores and reduce the accumulator on the fly.
//
// The GEMM runs TRANSPOSED -- D[channel, pixel] = sum_k W[channel, k] X[pixel, k]: M = 128 output
// channels (A operand = filter tile), N = N_PX pixels (B operand = activation tile, the same 4-D TMA
// box the forward kernel loads) -- so that a TMEM lane (= an epilogue thread) is ONE channel and the
// per-channel sums over pixels are serial sums inside a thread: no shuffles, no shared-memory
// round trips, no atomics in the loop. A CTA keeps its channel tile for the whole kernel, so the
// sums live in registers (fp32 per tile, fp64 across tiles) and are flushed once with fp64 atomics.
//
//   FWD:  acc[v][c] += (sum_p z, sum_p z^2)                      v = image & 1 (two views)
//   BWD:  acc[v][c] += (sum_p dy, sum_p dy * xhat),  xhat = (z - mean[v][c]) * invstd[v][c]
//         (dy arrives already masked by the ReLU of the block output: rmv_conv_args.mask_bits)
//
// Replaces the statistics half of nn.BatchNorm2d (train mode) forward/backward behind
// models/resnet.py:139-146 (autograd at trainer.py:142).
#include "common.cuh"
#include "bn_finalize.cuh"
#include "ops.h"

namespace rmv {
namespace {

constexpr int kTsBlockK = 64;
constexpr int kTsWBytes = 128 * kTsBlockK * 2;   // filter tile: 128 channels x 64 k
constexpr int kTsEpiThreads = 256;               // 8 epilogue warps

struct TStatArgs {
  CUtensorMap tmap_w;    // [c_out][K] bf16, box {64, 128}
  CUtensorMap tmap_x;    // (K, W, H, N), box {64, bw, bh, bn}: N_PX pixels per tile
  CUtensorMap tmap_dy;   // BWD: (c_out, W, H, N), box {64, bw, bh, bn}
  int box_n, tiles_w, tiles_h, tiles_n;
  int box_w, box_h;
  int k_blocks;          // K / 64
  int n_ct;              // c_out / 128
  int c_out;
  int flat_pix;          // > 0: flattened rows (tiles_h = tiles_n = 1), image of row P is P / flat_pix
  int ppi_shift;         // boxed: image of tile column r is tn * box_n + (r >> ppi_shift)
  double* acc;           // [2][c_out][2]
  const float* mean;     // BWD: [2][c_out]
  const float* invstd;
  BnFinalize fin;        // ticket != null: the last CTA turns the sums into coefficients
};

template <int N_PX, bool BWD>
struct TsCfg {
  static constexpr int kXBytes = N_PX * 128;
  static constexpr int kStageBytes = kTsWBytes + kXBytes;
  static constexpr int kStages = 4;
  static constexpr int kDyBytes = BWD ? 2 * N_PX * 128 : 0;   // two 64-channel boxes
  static constexpr int kDySlots = 2;
  static constexpr int kSmem = kStages * kStageBytes + kDySlots * kDyBytes + 256 + 1024;
  static constexpr int kThreads = BWD ? 352 : 320;
  static_assert(kSmem <= 232448, "shared memory budget exceeded");
};

template <int N_PX, bool BWD>
__global__ void __launch_bounds__(TsCfg<N_PX, BWD>::kThreads, 1)
tstat_kernel(const __grid_constant__ TStatArgs a) {
  using C = TsCfg<N_PX, BWD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_dy = smem + C::kStages * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dy + C::kDySlots * C::kDyBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* dy_full = tmem_empty + 2;
  uint64_t* dy_empty = dy_full + C::kDySlots;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dy_empty + C::kDySlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmap_w);
    tma_prefetch_desc(&a.tmap_x);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kTsEpiThreads); }
    for (int i = 0; i < C::kDySlots; ++i) { mbar_init(&dy_full[i], 1); mbar_init(&dy_empty[i], kTsEpiThreads); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 2 * N_PX); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();
  griddep_launch();

  // the CTA's channel tile is fixed; pixel tiles are strided over the CTAs that share it
  const int ct = blockIdx.x % a.n_ct;
  const int pt0 = blockIdx.x / a.n_ct, pt_step = gridDim.x / a.n_ct;
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pt = pt0; pt < m_tiles; pt += pt_step) {
        const int tw = pt % a.tiles_w;
        const int th = (pt / a.tiles_w) % a.tiles_h;
        const int tn = pt / (a.tiles_w * a.tiles_h);
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          uint8_t* sw = smem + stage * C::kStageBytes;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_2d(sw, &a.tmap_w, &full_bar[stage], kb * kTsBlockK, ct * 128);
          tma_load_4d(sw + kTsWBytes, &a.tmap_x, &full_bar[stage], kb * kTsBlockK, tw * a.box_w,
                      th * a.box_h, tn * a.box_n);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N_PX, 0, 0);
    int stage = 0; uint32_t phase = 0; int local = 0;
    for (int pt = pt0; pt < m_tiles; pt += pt_step, ++local) {
      const int accb = local & 1;
      mbar_wait(&tmem_empty[accb], ((local >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + accb * N_PX;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t sw = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sw, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sw + kTsWBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kTsBlockK / 16; ++k)
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == a.k_blocks - 1) umma_commit(&tmem_full[accb]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    if (BWD && lane == 0) {
      tma_prefetch_desc(&a.tmap_dy);
      int slot = 0; uint32_t phase = 0;
      for (int pt = pt0; pt < m_tiles; pt += pt_step) {
        const int tw = pt % a.tiles_w;
        const int th = (pt / a.tiles_w) % a.tiles_h;
        const int tn = pt / (a.tiles_w * a.tiles_h);
        mbar_wait(&dy_empty[slot], phase ^ 1);
        mbar_expect_tx(&dy_full[slot], C::kDyBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_4d(smem_dy + slot * C::kDyBytes + j * N_PX * 128, &a.tmap_dy, &dy_full[slot],
                      ct * 128 + j * 64, tw * a.box_w, th * a.box_h, tn * a.box_n);
        if (++slot == C::kDySlots) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..9) ----------------------
    const int quarter = warp & 3;            // TMEM lanes [32*quarter, +32) = channels of the tile
    const int half = (warp - 2) >> 2;        // pixel columns [half*N_PX/2, +N_PX/2)
    const int ch_in_tile = quarter * 32 + lane;
    constexpr int kCols = N_PX / 2, kChunks = kCols / 32;
    double t1[2] = {0.0, 0.0}, t2[2] = {0.0, 0.0};   // [view]: sums across tiles
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};    // fp32 sums of up to 4 tiles (<= 512 values each)
    // BWD: this thread's channel inside the two swizzled [N_PX][64 ch] boxes of the dy tile
    const uint32_t dy_box = (uint32_t)(ch_in_tile >> 6) * (uint32_t)(N_PX * 128);
    const uint32_t dy_unit = (uint32_t)(ch_in_tile & 63) >> 3, dy_el = (uint32_t)(ch_in_tile & 7) * 2;
    int local = 0, slot = 0;
    uint32_t dphase = 0;
    for (int pt = pt0; pt < m_tiles; pt += pt_step, ++local) {
      const int accb = local & 1;
      const int tw = pt % a.tiles_w;
      const int tn = pt / (a.tiles_w * a.tiles_h);
      // view-1 masks of this thread's column chunks (bit j = column chunk*32 + j belongs to view 1)
      uint32_t vm[kChunks];
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        const int col = half * kCols + ch * 32 + lane;
        const int img = a.flat_pix > 0 ? (int)((unsigned)(tw * a.box_w + col) / (unsigned)a.flat_pix)
                                       : tn * a.box_n + (col >> a.ppi_shift);
        vm[ch] = __ballot_sync(0xffffffffu, img & 1);
      }
      mbar_wait(&tmem_full[accb], (local >> 1) & 1);
      tc_fence_after_sync();
      const uint8_t* dyt = nullptr;
      if (BWD) {
        mbar_wait(&dy_full[slot], dphase);
        dyt = smem_dy + slot * C::kDyBytes + dy_box + dy_el;
      }
      // all of this thread's accumulator columns in one go (kChunks loads in flight, one wait), then
      // the accumulator goes straight back to the MMA warp
      uint32_t vv[kChunks][32];
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch)
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + accb * N_PX + half * kCols + ch * 32,
                           vv[ch]);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[accb]);
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        const uint32_t* v = vv[ch];
        float d[32];
        if (BWD) {
          const int p0 = half * kCols + ch * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t p = (uint32_t)(p0 + j);
            const uint16_t raw = *reinterpret_cast<const uint16_t*>(
                dyt + p * 128 + ((dy_unit ^ (p & 7u)) << 4));
            d[j] = __uint_as_float((uint32_t)raw << 16);
          }
        }
        const uint32_t m = vm[ch];
        if (m == 0u || m == 0xffffffffu) {   // whole chunk in one view (the common case)
          // four independent partial sums per statistic: the 4-cycle FADD/FFMA latency chains of a
          // single accumulator left the two warps of a scheduler idle half of the time
          float x1[4] = {0.f, 0.f, 0.f, 0.f}, x2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float z = __uint_as_float(v[j]);
            if (BWD) { x1[j & 3] += d[j]; x2[j & 3] = fmaf(d[j], z, x2[j & 3]); }
            else { x1[j & 3] += z; x2[j & 3] = fmaf(z, z, x2[j & 3]); }
          }
          const int vi = m != 0u;
          s1[vi] += (x1[0] + x1[1]) + (x1[2] + x1[3]);
          s2[vi] += (x2[0] + x2[1]) + (x2[2] + x2[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float z = __uint_as_float(v[j]);
            const float w1 = (float)((m >> j) & 1u), w0 = 1.f - w1;
            const float q = BWD ? d[j] : z;
            s1[0] = fmaf(w0, q, s1[0]); s2[0] = fmaf(w0 * q, z, s2[0]);
            s1[1] = fmaf(w1, q, s1[1]); s2[1] = fmaf(w1 * q, z, s2[1]);
          }
        }
      }
      if (BWD) {
        mbar_arrive(&dy_empty[slot]);
        if (++slot == C::kDySlots) { slot = 0; dphase ^= 1; }
      }
      if ((local & 3) == 3) {   // the fp64 pipe is narrow: one hand-over per four tiles
        t1[0] += (double)s1[0]; t1[1] += (double)s1[1];
        t2[0] += (double)s2[0]; t2[1] += (double)s2[1];
        s1[0] = s1[1] = s2[0] = s2[1] = 0.f;
      }
    }
    t1[0] += (double)s1[0]; t1[1] += (double)s1[1];
    t2[0] += (double)s2[0]; t2[1] += (double)s2[1];
    const int chn = ct * 128 + ch_in_tile;
#pragma unroll
    for (int vi = 0; vi < 2; ++vi) {
      double o1 = t1[vi], o2 = t2[vi];
      if (BWD) {   // sum dy*xhat = invstd * (sum dy*z - mean * sum dy)
        const double mu = (double)__ldg(a.mean + vi * a.c_out + chn);
        const double is = (double)__ldg(a.invstd + vi * a.c_out + chn);
        o2 = is * (o2 - mu * o1);
      }
      if (o1 != 0.0) atomicAdd(a.acc + ((long long)vi * a.c_out + chn) * 2, o1);
      if (o2 != 0.0) atomicAdd(a.acc + ((long long)vi * a.c_out + chn) * 2 + 1, o2);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * N_PX);
  }
  __shared__ unsigned int s_ticket;
  bn_last_block_finalize<BWD>(a.acc, a.fin, a.c_out, 2, gridDim.x, &s_ticket, a.mean, a.invstd);
}

template <int N_PX, bool BWD>
int launch_tstat(const TStatArgs& a, cudaStream_t stream) {
  using C = TsCfg<N_PX, BWD>;
  static bool attr_set = false;
  if (!attr_set) {
    RMV_CUDA(cudaFuncSetAttribute(tstat_kernel<N_PX, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::kSmem));
    attr_set = true;
  }
  const long m_tiles = (long)a.tiles_w * a.tiles_h * a.tiles_n;
  long per_ct = num_sms() / a.n_ct;
  if (per_ct < 1) per_ct = 1;
  if (per_ct > m_tiles) per_ct = m_tiles;
  const int grid = (int)(per_ct * a.n_ct);
  RMV_CUDA(launch_pdl_tc(tstat_kernel<N_PX, BWD>, dim3(grid), dim3(C::kThreads), C::kSmem, stream, a));
  return 0;
}

}  // namespace

// p describes the forward 1x1 convolution (x, w, strides, shapes; stride 1 or 2, pad 0). BWD: `dy`
// has the geometry (p.y_sn, p.y_sh, p.y_sw) of the conv output.
int conv_bn_reduce_tc(const ConvArgs& p, bool bwd, const void* dy, const float* mean,
                      const float* invstd, double* acc, const rmv_bn_params* finalize,
                      cudaStream_t stream) {
  RMV_CHECK_ARG(p.kh == 1 && p.kw == 1 && p.pad == 0 && (p.stride == 1 || p.stride == 2),
                "conv_bn_reduce: 1x1 convolutions (stride 1 or 2) only");
  RMV_CHECK_ARG(p.x_dtype == RMV_DTYPE_BF16 && p.c_in % 64 == 0 && p.c_out % 128 == 0,
                "conv_bn_reduce: bf16, c_in %% 64 == 0, c_out %% 128 == 0 (got %d -> %d)", p.c_in, p.c_out);
  RMV_CHECK_ARG(p.c_out / 128 <= num_sms(), "conv_bn_reduce: c_out too large");
  RMV_CHECK_ARG(acc != nullptr && (!bwd || (dy && mean && invstd)), "conv_bn_reduce: null pointer");
  RMV_CHECK_ARG(p.x_sw % 8 == 0 && p.x_sh % 8 == 0 && p.x_sn % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0,
                "conv_bn_reduce: strides must be multiples of 8 elements, pointers 16-byte aligned");
  const int n_px = bwd ? 128 : 256;
  TStatArgs a;
  memset(&a, 0, sizeof(a));
  int out_w = p.out_w, out_h = p.out_h, n_img = p.n_img;
  long long x_sw = p.x_sw * p.stride, x_sh = p.x_sh * p.stride, x_sn = p.x_sn;
  long long y_sw = p.y_sw, y_sh = p.y_sh, y_sn = p.y_sn;
  const bool x_dense = p.stride == 1 && (p.x_sh == p.x_sw * p.in_w) && (p.x_sn == p.x_sh * p.in_h);
  const bool y_dense = !bwd || ((p.y_sh == p.y_sw * p.out_w) && (p.y_sn == p.y_sh * p.out_h));
  if (x_dense && y_dense) {   // plain GEMM over the flattened pixel axis
    a.flat_pix = p.out_h * p.out_w;
    out_w = p.n_img * p.out_h * p.out_w; out_h = 1; n_img = 1;
    x_sh = x_sw * out_w; x_sn = x_sh;
    y_sh = y_sw * out_w; y_sn = y_sh;
  }
  int best_w = n_px, best_h = 1, best_n = 1;
  double best_eff = -1;
  for (int bw = n_px; bw >= 1; bw >>= 1)
    for (int bh = n_px / bw; bh >= 1; bh >>= 1) {
      const int bn = n_px / (bw * bh);
      if (bw > 256 || bh > 256 || bn > 256) continue;
      const double eff = (double)out_w * out_h * n_img /
                         ((double)ceil_div(out_w, bw) * bw * ceil_div(out_h, bh) * bh * ceil_div(n_img, bn) * bn);
      if (eff > best_eff + 1e-9) { best_eff = eff; best_w = bw; best_h = bh; best_n = bn; }
    }
  a.box_w = best_w; a.box_h = best_h; a.box_n = best_n;
  a.tiles_w = ceil_div(out_w, best_w); a.tiles_h = ceil_div(out_h, best_h); a.tiles_n = ceil_div(n_img, best_n);
  int sh = 0;
  while ((1 << sh) < best_w * best_h) ++sh;
  a.ppi_shift = sh;
  a.k_blocks = p.c_in / kTsBlockK;
  a.n_ct = p.c_out / 128;
  a.c_out = p.c_out;
  a.acc = acc; a.mean = mean; a.invstd = invstd;
  RMV_CHECK_ARG(finalize == nullptr || (finalize->ticket != nullptr && p.n_img % 2 == 0 &&
                                        (long long)(p.n_img / 2) * p.out_h * p.out_w > 1),
                "conv_bn_reduce: finalize needs a ticket counter and an even image count");
  a.fin = bn_finalize_args(finalize, bwd, (long long)(p.n_img / 2) * p.out_h * p.out_w);
  if ((long)a.tiles_w * a.tiles_h * a.tiles_n == 0) return 0;
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.c_in, (cuuint64_t)p.c_out};
    cuuint64_t strides[1] = {(cuuint64_t)p.c_in * 2};
    cuuint32_t box[2] = {(cuuint32_t)kTsBlockK, 128};
    if (int rc = encode_map(&a.tmap_w, p.w, 2, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.c_in, (cuuint64_t)out_w, (cuuint64_t)out_h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)(x_sw * 2), (cuuint64_t)(x_sh * 2), (cuuint64_t)(x_sn * 2)};
    cuuint32_t box[4] = {(cuuint32_t)kTsBlockK, (cuuint32_t)best_w, (cuuint32_t)best_h, (cuuint32_t)best_n};
    if (int rc = encode_map(&a.tmap_x, p.x, 4, dims, strides, box)) return rc;
  }
  if (bwd) {
    RMV_CHECK_ARG(y_sw % 8 == 0 && y_sh % 8 == 0 && y_sn % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(dy) & 15) == 0,
                  "conv_bn_reduce: dy strides must be multiples of 8 elements, pointer 16-byte aligned");
    cuuint64_t dims[4] = {(cuuint64_t)p.c_out, (cuuint64_t)out_w, (cuuint64_t)out_h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)(y_sw * 2), (cuuint64_t)(y_sh * 2), (cuuint64_t)(y_sn * 2)};
    cuuint32_t box[4] = {64, (cuuint32_t)best_w, (cuuint32_t)best_h, (cuuint32_t)best_n};
    if (int rc = encode_map(&a.tmap_dy, dy, 4, dims, strides, box)) return rc;
    return launch_tstat<128, true>(a, stream);
  }
  return launch_tstat<256, false>(a, stream);
}

}  // namespace rmv

extern "C" int rmv_conv_bn_stats(const rmv_conv_args* args, double* acc,
                                 const rmv_bn_params* finalize, void* stream) {
  RMV_CHECK_ARG(args != nullptr && args->x && args->w, "conv_bn_stats: null pointer");
  return rmv::conv_bn_reduce_tc(*args, false, nullptr, nullptr, nullptr, acc, finalize,
                                (cudaStream_t)stream);
}

extern "C" int rmv_conv_bn_bwd_reduce(const rmv_conv_args* args, const void* dy, const float* mean,
                                      const float* invstd, double* acc,
                                      const rmv_bn_params* finalize, void* stream) {
  RMV_CHECK_ARG(args != nullptr && args->x && args->w, "conv_bn_bwd_reduce: null pointer");
  return rmv::conv_bn_reduce_tc(*args, true, dy, mean, invstd, acc, finalize, (cudaStream_t)stream);
}
