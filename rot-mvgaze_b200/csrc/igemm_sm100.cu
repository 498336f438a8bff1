// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a).
//
//   D[m, n] = sum_{tap, c} A_tap[m, c] * W[n, tap*C + c]        m = output pixel (n_img, oh, ow)
//
// * A (activations, NHWC bf16) is never materialised as an im2col matrix: for every filter tap the
//   producer warp issues ONE 4-D tiled TMA load whose box is (64 channels, box_w, box_h, box_n)
//   output pixels shifted by the tap offset; out-of-bounds coordinates (the conv zero padding) are
//   zero-filled by the TMA unit. Stride-2 convs read one of four "parity planes" of the input
//   (separate tensor maps with doubled strides), so every tap is still a dense box.
// * W (filters, [c_out][kh][kw][c_in] bf16 = K-major) is a plain 2-D TMA load.
// * tcgen05.mma (cta_group::1, kind::f16, M=128, N=BLOCK_N, K=16) accumulates in TMEM (fp32); two
//   accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
// * The epilogue (4 warps, one TMEM lane quarter each) applies the folded BatchNorm scale/shift or
//   bias, the residual add and ReLU, and writes bf16 / fp32 NHWC with arbitrary pixel strides.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.BatchNorm2d / nn.ReLU / nn.Linear at
// reference models/resnet.py:31-47,128-148 and models/backbones/blocks.py:41-60.
#include "common.cuh"
#include "ops.h"

#include <mutex>

namespace rmv {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 49;
constexpr int kNumThreads = 192;  // warp0 TMA, warp1 MMA, warps2-5 epilogue
constexpr int kABytes = kBlockM * kBlockK * 2;

struct IgemmArgs {
  CUtensorMap tmap_a[4];
  CUtensorMap tmap_b;
  int out_w, out_h, n_img;
  int box_w, box_h, box_n;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles, n_total;
  int c_blocks, num_taps;
  signed char tap_map[kMaxTaps];
  signed char tap_dw[kMaxTaps];
  signed char tap_dh[kMaxTaps];
  void* out;
  long long os_n, os_h, os_w;
  const __nv_bfloat16* residual;
  long long rs_n, rs_h, rs_w;
  const float* scale;
  const float* shift;
  int relu;
  int out_fp32;
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 128 / 256 / 512: powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
igemm_kernel(const __grid_constant__ IgemmArgs args) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                      // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + C::kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * C::kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;           // [2]        epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmap_b);
    tma_prefetch_desc(&args.tmap_a[0]);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int m_tiles = args.tiles_w * args.tiles_h * args.tiles_n;
  const int total_tiles = m_tiles * args.n_tiles;
  const int num_kb = args.num_taps * args.c_blocks;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % args.n_tiles;
        const int m_tile = tile / args.n_tiles;
        const int tw = m_tile % args.tiles_w;
        const int th = (m_tile / args.tiles_w) % args.tiles_h;
        const int tn = m_tile / (args.tiles_w * args.tiles_h);
        const int ow0 = tw * args.box_w, oh0 = th * args.box_h, n0 = tn * args.box_n;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_4d(smem_a + stage * kABytes, &args.tmap_a[args.tap_map[tap]], &full_bar[stage],
                      cb * kBlockK, ow0 + args.tap_dw[tap], oh0 + args.tap_dh[tap], n0);
          tma_load_2d(smem_b + stage * C::kBBytes, &args.tmap_b, &full_bar[stage], kb * kBlockK,
                      n_tile * BLOCK_N);
          if (++cb == args.c_blocks) { cb = 0; ++tap; }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * kABytes), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * C::kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes per K=16 step inside the 128-byte swizzle row (>>4 -> +2)
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    const int row = quarter * 32 + lane;
    const int dw = row % args.box_w;
    const int dh = (row / args.box_w) % args.box_h;
    const int dn = row / (args.box_w * args.box_h);
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_tile = tile % args.n_tiles;
      const int m_tile = tile / args.n_tiles;
      const int tw = m_tile % args.tiles_w;
      const int th = (m_tile / args.tiles_w) % args.tiles_h;
      const int tn = m_tile / (args.tiles_w * args.tiles_h);
      const int ow = tw * args.box_w + dw, oh = th * args.box_h + dh, n = tn * args.box_n + dn;
      const bool valid = (ow < args.out_w) && (oh < args.out_h) && (n < args.n_img);
      const long long o_off = n * args.os_n + oh * args.os_h + ow * args.os_w;
      const long long r_off = n * args.rs_n + oh * args.rs_h + ow * args.rs_w;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + c * 32,
                           v);
        tmem_ld_wait();
        const int col0 = n_tile * BLOCK_N + c * 32;
        if (valid && col0 < args.n_total) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (args.scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= __ldg(args.scale + col0 + j);
          }
          if (args.shift != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(args.shift + col0 + j);
          }
          if (args.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(args.residual + r_off + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 r = __ldg(rp + q);
              const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 p = unpack_bf16x2(w[t]);
                f[q * 8 + t * 2] += p.x;
                f[q * 8 + t * 2 + 1] += p.y;
              }
            }
          }
          if (args.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (args.out_fp32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + o_off + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(args.out) + o_off + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
              o.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
              o.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
              o.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
              op[q] = o;
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  RMV_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RMV_CHECK_ARG(r == CUDA_SUCCESS,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] strides "
                "[%llu %llu %llu] box [%u %u %u %u] base %p",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)strides_bytes[0],
                (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1],
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
  return 0;
}

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int BLOCK_N>
int launch(const IgemmArgs& a, int total_tiles, cudaStream_t stream) {
  using C = Cfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    RMV_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::kSmemBytes));
    attr_set = true;
  }
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  igemm_kernel<BLOCK_N><<<grid, kNumThreads, C::kSmemBytes, stream>>>(a);
  RMV_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int conv_fwd_tc(const ConvArgs& p, cudaStream_t stream) {
  RMV_CHECK_ARG(p.c_in % kBlockK == 0, "tcgen05 conv: c_in=%d must be a multiple of 64", p.c_in);
  RMV_CHECK_ARG(p.c_out % 8 == 0, "tcgen05 conv: c_out=%d must be a multiple of 8", p.c_out);
  RMV_CHECK_ARG(p.stride == 1 || p.stride == 2, "tcgen05 conv: stride %d unsupported", p.stride);
  RMV_CHECK_ARG(p.kh * p.kw <= kMaxTaps, "tcgen05 conv: %dx%d filter too large", p.kh, p.kw);
  RMV_CHECK_ARG(p.x_sw % 8 == 0 && p.x_sh % 8 == 0 && p.x_sn % 8 == 0,
                "tcgen05 conv: input pixel strides must be multiples of 8 elements");
  RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(p.x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.y) & 15) == 0,
                "tcgen05 conv: pointers must be 16-byte aligned");
  RMV_CHECK_ARG(p.y_sw % 8 == 0 && p.y_sh % 8 == 0 && p.y_sn % 8 == 0,
                "tcgen05 conv: output pixel strides must be multiples of 8 elements");

  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  int out_w = p.out_w, out_h = p.out_h, n_img = p.n_img;
  long long x_sw = p.x_sw, x_sh = p.x_sh, x_sn = p.x_sn;
  int in_w = p.in_w, in_h = p.in_h;
  a.os_n = p.y_sn; a.os_h = p.y_sh; a.os_w = p.y_sw;
  a.rs_n = p.r_sn; a.rs_h = p.r_sh; a.rs_w = p.r_sw;

  // 1x1 stride-1 over a dense pixel grid is a plain GEMM: flatten (n, h, w) into one axis.
  const bool pointwise = (p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad == 0);
  const bool x_dense = (p.x_sh == p.x_sw * p.in_w) && (p.x_sn == p.x_sh * p.in_h);
  const bool y_dense = (p.y_sh == p.y_sw * p.out_w) && (p.y_sn == p.y_sh * p.out_h);
  const bool r_dense =
      p.residual == nullptr || ((p.r_sh == p.r_sw * p.out_w) && (p.r_sn == p.r_sh * p.out_h));
  if (pointwise && x_dense && y_dense && r_dense) {
    out_w = in_w = p.n_img * p.in_h * p.in_w;
    out_h = in_h = 1;
    n_img = 1;
    x_sh = x_sw * in_w; x_sn = x_sh;
    a.os_h = a.os_n = 0;
    a.rs_h = a.rs_n = 0;
  }

  // Pick the (box_w, box_h, box_n) factorisation of the 128-row M tile with the least padding.
  int best_w = 128, best_h = 1, best_n = 1;
  double best_eff = -1;
  for (int bw = 128; bw >= 1; bw >>= 1)
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bn = 128 / (bw * bh);
      const double eff = (double)out_w * out_h * n_img /
                         ((double)ceil_div(out_w, bw) * bw * ceil_div(out_h, bh) * bh *
                          ceil_div(n_img, bn) * bn);
      if (eff > best_eff + 1e-9) { best_eff = eff; best_w = bw; best_h = bh; best_n = bn; }
    }
  a.box_w = best_w; a.box_h = best_h; a.box_n = best_n;
  a.out_w = out_w; a.out_h = out_h; a.n_img = n_img;
  a.tiles_w = ceil_div(out_w, a.box_w);
  a.tiles_h = ceil_div(out_h, a.box_h);
  a.tiles_n = ceil_div(n_img, a.box_n);

  // Activation tensor maps: one per input parity plane that some tap touches.
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
        RMV_CHECK_ARG(pl_w > 0 && pl_h > 0, "tcgen05 conv: empty parity plane");
        cuuint64_t dims[4] = {(cuuint64_t)p.c_in, (cuuint64_t)pl_w, (cuuint64_t)pl_h,
                              (cuuint64_t)n_img};
        cuuint64_t strides[3] = {(cuuint64_t)(x_sw * s * 2), (cuuint64_t)(x_sh * s * 2),
                                 (cuuint64_t)(x_sn * 2)};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)a.box_w, (cuuint32_t)a.box_h,
                             (cuuint32_t)a.box_n};
        const __nv_bfloat16* base =
            reinterpret_cast<const __nv_bfloat16*>(p.x) + ph * x_sh + pw * x_sw;
        int rc = encode_map(&a.tmap_a[n_planes], base, 4, dims, strides, box);
        if (rc) return rc;
        plane_id[ph][pw] = n_planes++;
      }
      a.tap_map[t] = (signed char)plane_id[ph][pw];
      a.tap_dh[t] = (signed char)floordiv(r - p.pad, s);
      a.tap_dw[t] = (signed char)floordiv(q - p.pad, s);
    }

  const int block_n = (p.block_n > 0) ? p.block_n
                      : (p.c_out <= 64) ? 64
                      : (p.c_out % 256 == 0 && (long)a.tiles_w * a.tiles_h * a.tiles_n * (p.c_out / 256) >= 2L * num_sms()) ? 256
                      : 128;
  RMV_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "bad block_n %d", block_n);
  {
    const long long k_total = (long long)a.num_taps * p.c_in;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)p.c_out};
    cuuint64_t strides[1] = {(cuuint64_t)(k_total * 2)};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
    int rc = encode_map(&a.tmap_b, p.w, 2, dims, strides, box);
    if (rc) return rc;
  }
  a.n_total = p.c_out;
  a.n_tiles = ceil_div(p.c_out, block_n);
  a.c_blocks = p.c_in / kBlockK;
  a.out = p.y;
  a.residual = reinterpret_cast<const __nv_bfloat16*>(p.residual);
  a.scale = p.scale;
  a.shift = p.shift;
  a.relu = p.relu;
  a.out_fp32 = (p.y_dtype == RMV_DTYPE_F32);
  const int total = a.tiles_w * a.tiles_h * a.tiles_n * a.n_tiles;
  if (total == 0) return 0;
  switch (block_n) {
    case 64: return launch<64>(a, total, stream);
    case 128: return launch<128>(a, total, stream);
    default: return launch<256>(a, total, stream);
  }
}

}  // namespace rmv
