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
#include "bn_finalize.cuh"
#include "ops.h"

#include <mutex>
#include <stdlib.h>

namespace rmv {

namespace {

#ifndef RMV_EPI_PREFETCH
#define RMV_EPI_PREFETCH 1   // 0: round-1 epilogue (one TMEM load per chunk, waited for at once)
#endif
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 49;
constexpr int kEpiThreads = 256;  // 8 epilogue warps
constexpr int kResWarp = 10;
constexpr int kNumThreads = 352;  // warp0 TMA(A,B), warp1 MMA, warps2-9 epilogue, warp10 TMA(residual)
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kChunkBytes = kBlockM * 128;  // one staged output / residual chunk: 128 rows x 128 B

struct IgemmArgs {
  CUtensorMap tmap_a[4];
  CUtensorMap tmap_b;
  CUtensorMap tmap_out;  // store map: (c_out, W, H, N), box (128 B of channels, box_w, box_h, box_n)
  CUtensorMap tmap_res;  // residual load map, same geometry (bf16)
  int box_w, box_h, box_n;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles, n_total;
  int c_blocks, num_taps;
  signed char tap_map[kMaxTaps];
  signed char tap_dw[kMaxTaps];
  signed char tap_dh[kMaxTaps];
  signed char tap_w[kMaxTaps];  // index of the tap inside the filter tensor (B operand K offset)
  const float* scale;
  const float* shift;
  int relu;
  int halo_base_mode;  // HALO kernels: 1 = set the descriptor base offset for row-shifted starts
  // STATS kernels (training): per-(view, channel) sum and sum of squares of the bf16 output, added
  // into stat_acc[2][n_total][2] (fp64). Image n belongs to view n & 1.
  double* stat_acc;
  int stat_pix;        // > 0: flattened pointwise GEMM, image of output row P is P / stat_pix
  long long stat_rows; //      ... and P < stat_rows are the rows of the tensor
  int stat_ppi_shift;  // boxed tiles: image of tile row r is tn*box_n + (r >> stat_ppi_shift),
  int stat_bw_shift;   //   its pixel (th*box_h + ((r >> stat_bw_shift) & (box_h-1)), tw*box_w + (r & (box_w-1)))
  int stat_w, stat_h, stat_n;  // output extent
  // BNM kernels (training, recomputed BatchNorm): per-(view, channel) coefficient tables [2][n_total]
  //   BNM 1 (forward apply):  y = relu?(bn_a*z + bn_b (+ residual)); bn_bits <- sign mask of the pre-ReLU value
  //   BNM 2 (backward apply): y = bn_a*dy + bn_b*z + bn_c with dy = the "residual" tile
  const float* bn_a;
  const float* bn_b;
  const float* bn_c;
  // packed ReLU sign masks, 1 bit per element of a tensor with the geometry of y (element offset =
  // n*m_sn + oh*m_sh + ow*m_sw + channel, + mask_off; bit e of byte off/8 = element off, e = off%8):
  uint32_t* bn_bits;          // BNM 1: written
  const uint32_t* mask_bits;  // any kernel: the value about to be stored is zeroed where the bit is 0
  long long m_sn, m_sh, m_sw, mask_off;
  BnFinalize stat_fin;        // STATS: ticket != null -> the last CTA finalizes the statistics
  // split-K (small-M linear layers): tile = split * (m_tiles * n_tiles) + (m, n) tile; split s owns the
  // k-blocks [s * kb_per_split, (s+1) * kb_per_split) and stores its fp32 partial tile into slice s
  // of a workspace (4th coordinate of tmap_out); splitk_reduce_kernel adds the slices in order.
  int k_splits, kb_per_split;
};

// HALO variant (3x3, stride 1, pad 1, 64 -> 64 channels; BLOCK_N = 64): the producer loads ONE
// (16 wide x 18 high) input patch per 8x16-pixel output tile -- 16-pixel rows so that every image
// row starts 2048 B (a whole number of 1024-byte swizzle atoms) after the previous one -- and the
// nine taps are nine UMMA descriptors into that patch (start shifted by (r*16+s) pixels = rows of
// 128 B, stride 2048 B between the 8-row groups), instead of nine separate TMA boxes: 6x less
// L2->SM traffic for A. The 72 KiB of filters stay resident in shared memory for the whole kernel.
constexpr int kHaloW = 16, kHaloH = 18;
constexpr int kHaloABytes = kHaloW * kHaloH * 128;  // 36864
constexpr int kHaloTaps = 9;
constexpr int kHaloStages = 3;

// RES: 0 = no residual; 1 = residual, deep residual ring (the HBM-bound layers with 1-4 k-blocks per
// tile); 2 = residual, deep A/B ring (K >= 512: 3 stages + 2 residual slots; measured 0.234 -> 0.184 ms
// on the 7x7 512->2048 layers, slower on the shallow-K ones)
template <int BLOCK_N, int RES, bool HALO = false, bool STATS = false, int BNM = 0>
struct Cfg {
  static constexpr bool HAS_RES = RES != 0;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageA = HALO ? kHaloABytes : kABytes;             // A bytes per stage
  static constexpr int kStageB = HALO ? 0 : kBBytes;                        // B bytes per stage
  static constexpr int kResidentB = HALO ? kHaloTaps * kBBytes : 0;         // filters kept in smem
  static constexpr int kStageBytes = kStageA + kStageB;
  // Layers with a residual (the expanding 1x1 convs) are HBM-bound with 1-8 k-blocks per tile: they
  // trade A/B stages for a deeper residual ring (4 x 16 KiB) so the residual loads run well ahead.
  static constexpr int kStages = HALO ? kHaloStages
                                 : HAS_RES ? (BLOCK_N == 256 ? (RES == 2 ? 3 : 2) : (BLOCK_N == 128 ? 3 : 4))
                                           : (BLOCK_N == 256 ? 3 : ((STATS || BNM != 0) && BLOCK_N == 128 ? 4 : 6));
  static constexpr int kResSlots = (BLOCK_N == 256 && RES == 2) ? 2 : 4;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 128 / 256 / 512: powers of two
  static constexpr int kOutBytes = 2 * kChunkBytes;
  static constexpr int kResBytes = HAS_RES ? kResSlots * kChunkBytes : 0;
  // scale + shift of the current N tile; BNM: 2 (a, b) or 3 (a, b, c) tables for each of the 2 views
  static constexpr int kVecBytes = (BNM == 1 ? 4 : (BNM == 2 ? 6 : 2)) * BLOCK_N * 4;
  // BNM 1 / 3: the packed ReLU-mask words of one tile, [128 rows][BLOCK_N / 32], staged so that global
  // memory sees whole 32-byte sectors per row instead of one 4-byte access per thread and chunk
  static constexpr int kBitsBytes = (BNM == 1 || BNM == 3 || BNM == 4) ? 128 * (BLOCK_N / 32) * 4 : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kResidentB + kOutBytes + kResBytes +
                                    kVecBytes + kBitsBytes + 256 /*barriers*/ + 1024 /*align*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync(int id) {
  asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
}

// OUT_F32: 32 fp32 columns per staged chunk; otherwise 64 bf16 columns (both 128 B per row).
template <int BLOCK_N, int RES, bool OUT_F32, bool HALO = false, bool STATS = false, int BNM = 0>
__global__ void __launch_bounds__(kNumThreads, 1)
igemm_kernel(const __grid_constant__ IgemmArgs args) {
  constexpr bool HAS_RES = RES != 0;
  using C = Cfg<BLOCK_N, RES, HALO, STATS, BNM>;
  static_assert(!STATS || (!HAS_RES && !OUT_F32), "STATS: bf16 output, no residual");
  static_assert(BNM == 0 || (!OUT_F32 && !HALO && !STATS), "BNM: bf16 output, plain tiles");
  constexpr bool MASK = BNM == 3 || BNM == 4;   // plain epilogue + packed ReLU mask applied to the stored value
  // BNM 4 (training data gradients of the mid layers): the masked gradient dy is stored AND reduced for
  // the BatchNorm backward of the layer below -- sum dy, sum dy*z per (view, channel), with z (the
  // pre-BatchNorm conv output of that layer) arriving through the residual ring (not added)
  constexpr bool BSTAT = BNM == 4;
  constexpr bool ST = STATS || BSTAT;   // the column-statistics machinery of the epilogue
  static_assert(!BSTAT || HAS_RES, "BNM 4 reads z through the residual ring");
  constexpr bool BNT = BNM == 1 || BNM == 2;   // per-view BatchNorm coefficient tables
  static_assert(BNM != 2 || HAS_RES, "BNM 2 reads dy through the residual ring");
  constexpr int kChunkCols = OUT_F32 ? 32 : 64;
  constexpr int kChunks = BLOCK_N / kChunkCols;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by an integer round trip of the pointer) so the compiler
  // keeps the shared address space and emits LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + C::kStages * C::kStageA;
  uint8_t* smem_out = smem_b + C::kStages * C::kStageB + C::kResidentB;  // [2][128 rows][128 B], SW128
  uint8_t* smem_res = smem_out + C::kOutBytes;           // [2][128 rows][128 B], SW128
  float* s_scale = reinterpret_cast<float*>(smem_res + C::kResBytes);
  float* s_shift = s_scale + (BNT ? 2 : 1) * BLOCK_N;   // BNT: [2 views][BLOCK_N] per table
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_res + C::kResBytes + C::kVecBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_res + C::kResBytes + C::kVecBytes + C::kBitsBytes);
  uint64_t* full_bar = bars;                      // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + C::kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * C::kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;           // [2]        epilogue -> MMA
  uint64_t* res_full = tmem_empty + 2;            // [4]        TMA(residual) -> epilogue
  uint64_t* res_empty = res_full + C::kResSlots;  // [4]        epilogue -> TMA(residual)
  uint64_t* b_bar = res_empty + C::kResSlots;     // [1]        HALO: resident filters landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_bar + 1);
  // STATS: [2 views][n_total][2] fp32 partial sums of this CTA, after the fixed-size regions
  float* s_slot = reinterpret_cast<float*>(smem + C::kSmemBytes - 1024);  // [8][4][64] fp32
  float* s_stat = s_slot + 8 * 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmap_b);
    tma_prefetch_desc(&args.tmap_a[0]);
    tma_prefetch_desc(&args.tmap_out);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiThreads);
    }
    for (int i = 0; i < C::kResSlots; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], kEpiThreads);
    }
    mbar_init(b_bar, 1);
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
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail
  // of the previous kernel; global memory is only touched from here on.
  griddep_wait();
  griddep_launch();

  const int m_tiles = args.tiles_w * args.tiles_h * args.tiles_n;
  const int mn_tiles = m_tiles * args.n_tiles;
  const int total_tiles = mn_tiles * args.k_splits;
  const int num_kb = args.num_taps * args.c_blocks;

  if (warp == 0) {
    // ------------------------------- TMA producer (A, B) ------------------------
    if (HALO && lane == 0) {
      // resident filters: tap t = rows [t*64, t*64+64) of the K axis, [64 c_out][64 c_in] each
      mbar_expect_tx(b_bar, C::kResidentB);
      for (int t = 0; t < kHaloTaps; ++t)
        tma_load_2d(smem_b + t * C::kBBytes, &args.tmap_b, b_bar, t * kBlockK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tw = tile % args.tiles_w;
        const int th = (tile / args.tiles_w) % args.tiles_h;
        const int tn = tile / (args.tiles_w * args.tiles_h);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], kHaloABytes);
        tma_load_4d(smem_a + stage * kHaloABytes, &args.tmap_a[0], &full_bar[stage], 0,
                    tw * args.box_w - 1, th * args.box_h - 1, tn);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int split = 0, mn = tile;
        if (args.k_splits > 1) { split = tile / mn_tiles; mn = tile - split * mn_tiles; }
        const int n_tile = mn % args.n_tiles;
        const int m_tile = mn / args.n_tiles;
        const int tw = m_tile % args.tiles_w;
        const int th = (m_tile / args.tiles_w) % args.tiles_h;
        const int tn = m_tile / (args.tiles_w * args.tiles_h);
        const int ow0 = tw * args.box_w, oh0 = th * args.box_h, n0 = tn * args.box_n;
        const int kb0 = split * args.kb_per_split;
        const int kb1 = (kb0 + args.kb_per_split < num_kb) ? kb0 + args.kb_per_split : num_kb;
        int tap = kb0 / args.c_blocks, cb = kb0 - tap * args.c_blocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_4d(smem_a + stage * kABytes, &args.tmap_a[args.tap_map[tap]], &full_bar[stage],
                      cb * kBlockK, ow0 + args.tap_dw[tap], oh0 + args.tap_dh[tap], n0);
          tma_load_2d(smem_b + stage * C::kBBytes, &args.tmap_b, &full_bar[stage],
                      (args.tap_w[tap] * args.c_blocks + cb) * kBlockK, n_tile * BLOCK_N);
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
      if (HALO) {
        if (local == 0) mbar_wait(b_bar, 0);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t a0 = smem_u32(smem_a + stage * kHaloABytes);
          const uint32_t b0 = smem_u32(smem_b);
#pragma unroll
          for (int tap = 0; tap < kHaloTaps; ++tap) {
            // tap (r, s): the patch shifted by r rows of 16 pixels and s pixels (128 B each)
            const uint32_t aaddr = a0 + ((tap / 3) * kHaloW + (tap % 3)) * 128;
            const uint64_t adesc = umma_desc_sw128(
                aaddr, 16, kHaloW * 128, args.halo_base_mode ? (aaddr >> 7) & 7 : 0);
            const uint64_t bdesc = umma_desc_sw128(b0 + tap * C::kBBytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        continue;
      }
      const int split = args.k_splits > 1 ? tile / mn_tiles : 0;
      const int kb0 = split * args.kb_per_split;
      const int kb1 = (kb0 + args.kb_per_split < num_kb) ? kb0 + args.kb_per_split : num_kb;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * kABytes), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * C::kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 bytes per K=16 step inside the 128-byte swizzle row (>>4 -> +2)
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kResWarp) {
    // ------------------------------- TMA producer (residual) --------------------
    if (HAS_RES && lane == 0) {
      tma_prefetch_desc(&args.tmap_res);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {   // k_splits == 1 with a residual
        const int n_tile = tile % args.n_tiles;
        const int m_tile = tile / args.n_tiles;
        const int tw = m_tile % args.tiles_w;
        const int th = (m_tile / args.tiles_w) % args.tiles_h;
        const int tn = m_tile / (args.tiles_w * args.tiles_h);
        for (int c = 0; c < kChunks; ++c) {
          mbar_wait(&res_empty[slot], phase ^ 1);
          mbar_expect_tx(&res_full[slot], kChunkBytes);
          tma_load_4d(smem_res + slot * kChunkBytes, &args.tmap_res, &res_full[slot],
                      n_tile * BLOCK_N + c * kChunkCols, tw * args.box_w, th * args.box_h,
                      tn * args.box_n);
          if (++slot == C::kResSlots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------- epilogue (warps 2..9) ----------------------
    // Two warps per TMEM lane quarter (= two per SM sub-partition, so one can issue while the
    // other waits on TMEM / shared memory); each owns half of the chunk's columns = 64 B per row.
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int tid_e = threadIdx.x - 64;  // 0..255
    const uint32_t sw = (uint32_t)(row & 7);
    constexpr int kWarpCols = kChunkCols / 2;  // 32 (bf16 out) or 16 (fp32 out)
    int local = 0;
    int cc = 0;  // running chunk counter: staging buffer = cc & 1
    int rslot = 0;
    uint32_t rphase = 0;
    // STATS: thread = (column pair cp, 16-row group rg) of a staged chunk
    const int st_cp = tid_e & 31, st_rg = tid_e >> 5;
    if (ST) {
      for (int i = tid_e; i < 4 * args.n_total; i += kEpiThreads) s_stat[i] = 0.f;
      epi_bar_sync(1);
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      int split = 0, mn = tile;
      if (args.k_splits > 1) { split = tile / mn_tiles; mn = tile - split * mn_tiles; }
      const int n_tile = mn % args.n_tiles;
      const int m_tile = mn / args.n_tiles;
      const int tw = m_tile % args.tiles_w;
      const int th = (m_tile / args.tiles_w) % args.tiles_h;
      const int tn = m_tile / (args.tiles_w * args.tiles_h);
      // STATS: bit i of st_m0 / st_m1 = row st_rg*16 + i of this tile is a pixel of the tensor and
      // belongs to view 0 / view 1 (image & 1). Rows outside the tensor must not be counted: a 3x3
      // conv gives them non-zero values when their receptive field reaches into the image.
      uint32_t st_m0 = 0, st_m1 = 0;
      if (ST) {
        if (args.stat_pix > 0) {   // flattened: rows are consecutive pixels of consecutive images
          const long long p0 = (long long)tw * args.box_w + st_rg * 16;
          const int img0 = (int)(p0 / args.stat_pix);
          const int rem = (int)(p0 - (long long)img0 * args.stat_pix);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t ok = (p0 + i < args.stat_rows) ? 1u : 0u;
            const uint32_t v = (uint32_t)((img0 + ((rem + i) / args.stat_pix)) & 1);
            st_m0 |= (ok & (v ^ 1u)) << i;
            st_m1 |= (ok & v) << i;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = st_rg * 16 + i;
            const int ow = tw * args.box_w + (r & (args.box_w - 1));
            const int oh = th * args.box_h + ((r >> args.stat_bw_shift) & (args.box_h - 1));
            const int n = tn * args.box_n + (r >> args.stat_ppi_shift);
            const uint32_t ok = (ow < args.stat_w && oh < args.stat_h && n < args.stat_n) ? 1u : 0u;
            st_m0 |= (ok & (uint32_t)((n & 1) ^ 1)) << i;
            st_m1 |= (ok & (uint32_t)(n & 1)) << i;
          }
        }
      }
      // per-channel scale/shift of this N tile -> smem (all readers of the previous tile's values
      // are behind the last epi_bar_sync(2) of that tile)
      if (!BNT) {
        for (int i = tid_e; i < BLOCK_N; i += kEpiThreads) {
          const int col = n_tile * BLOCK_N + i;
          const bool ok = col < args.n_total;
          s_scale[i] = (ok && args.scale != nullptr) ? __ldg(args.scale + col) : 1.f;
          s_shift[i] = (ok && args.shift != nullptr) ? __ldg(args.shift + col) : 0.f;
        }
      } else {
        for (int i = tid_e; i < 2 * BLOCK_N; i += kEpiThreads) {
          const int vi = i / BLOCK_N, col = n_tile * BLOCK_N + (i - vi * BLOCK_N);
          const bool ok = col < args.n_total;
          s_scale[i] = ok ? __ldg(args.bn_a + vi * args.n_total + col) : 0.f;
          s_shift[i] = ok ? __ldg(args.bn_b + vi * args.n_total + col) : 0.f;
          if (BNM == 2) s_shift[2 * BLOCK_N + i] = ok ? __ldg(args.bn_c + vi * args.n_total + col) : 0.f;
        }
      }
      // this thread's output row: validity, image (-> view) and element offset inside a tensor of
      // y's geometry (for the packed ReLU masks)
      bool row_ok = true;
      int row_view = 0;
      long long row_off = 0;
      // (validity, image, element offset) of tile row r
      auto row_info = [&](int r, bool& ok, int& img, long long& off) {
        if (args.stat_pix > 0) {
          const long long p = (long long)tw * args.box_w + r;
          ok = p < args.stat_rows;
          img = (int)(p / args.stat_pix);
          off = p * args.m_sw + args.mask_off;
        } else {
          const int ow = tw * args.box_w + (r & (args.box_w - 1));
          const int oh = th * args.box_h + ((r >> args.stat_bw_shift) & (args.box_h - 1));
          img = tn * args.box_n + (r >> args.stat_ppi_shift);
          ok = ow < args.stat_w && oh < args.stat_h && img < args.stat_n;
          off = (long long)img * args.m_sn + (long long)oh * args.m_sh + (long long)ow * args.m_sw +
                args.mask_off;
        }
      };
      if (BNT) {
        int img;
        row_info(row, row_ok, img, row_off);
        row_view = img & 1;
      }
      // mask words of this tile <-> global memory: thread = (row fr, half fq of its BLOCK_N/32 words)
      constexpr int kBitW = BLOCK_N / 32;        // words per row: 2 / 4 / 8
      constexpr int kBitH = kBitW / 2;           // words per thread: 1 / 2 / 4
      const int fr = tid_e >> 1, fq = tid_e & 1;
      bool f_ok = false;
      long long f_word = 0;
      if (BNM == 1 || MASK) {
        int img; long long off;
        row_info(fr, f_ok, img, off);
        f_ok = f_ok && (n_tile * BLOCK_N + fq * kBitH * 32 < args.n_total);
        f_word = (off + n_tile * BLOCK_N) / 32 + fq * kBitH;
      }
      if (MASK) {
        // (the previous tile's readers are behind its last epi_bar_sync(2); the first use below is
        // behind this tile's first epi_bar_sync(1))
        uint32_t* dst = s_bits + fr * kBitW + fq * kBitH;
        if (kBitH == 4) {
          *reinterpret_cast<uint4*>(dst) = f_ok ? __ldg(reinterpret_cast<const uint4*>(args.mask_bits + f_word))
                                                : make_uint4(0, 0, 0, 0);
        } else if (kBitH == 2) {
          *reinterpret_cast<uint2*>(dst) = f_ok ? __ldg(reinterpret_cast<const uint2*>(args.mask_bits + f_word))
                                                : make_uint2(0, 0);
        } else {
          *dst = f_ok ? __ldg(args.mask_bits + f_word) : 0u;
        }
      }
      const float* t_a = s_scale + row_view * BLOCK_N;
      const float* t_b = s_shift + row_view * BLOCK_N;
      const float* t_c = s_shift + (2 + row_view) * BLOCK_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      // bf16 output: the chunk loop is unrolled and the TMEM load of chunk c+1 is issued as soon as
      // chunk c has landed, so its latency hides behind chunk c's arithmetic, staging and barriers
      // (the epilogue of the HBM-bound layers is a latency chain with two warps per scheduler)
      constexpr bool kPrefetch = !OUT_F32 && RMV_EPI_PREFETCH;
      uint32_t vbuf[kPrefetch ? 2 : 1][32];
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * kWarpCols;
      if (kPrefetch) tmem_ld_32x32b_x32(taddr0, vbuf[0]);
#pragma unroll(kPrefetch ? kChunks : 1)
      for (int c = 0; c < kChunks; ++c, ++cc) {
        uint8_t* obuf = smem_out + (cc & 1) * kChunkBytes;
        // (1) staging buffer free: the store issued two chunks ago has finished reading it
        if (tid_e == 0) tma_store_wait_read<1>();
        epi_bar_sync(1);
        if (HAS_RES) mbar_wait(&res_full[rslot], rphase);
        const uint8_t* rbuf = smem_res + rslot * kChunkBytes;
        // (2) TMEM -> registers -> scale/shift(/residual)/ReLU -> swizzled smem
        const int col_in_tile = c * kChunkCols + half * kWarpCols;
        uint32_t* v = vbuf[kPrefetch ? (c & 1) : 0];
        if (!kPrefetch) {
          if (OUT_F32) tmem_ld_32x32b_x16(taddr0 + c * kChunkCols, v);
          else tmem_ld_32x32b_x32(taddr0 + c * kChunkCols, v);
        }
        tmem_ld_wait();
        if (kPrefetch && c + 1 < kChunks)
          tmem_ld_32x32b_x32(taddr0 + (c + 1) * kChunkCols, vbuf[(c + 1) & 1]);
        if (c == kChunks - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp early
          tc_fence_before_sync();
          mbar_arrive(&tmem_empty[acc]);
        }
        float f[32];
#pragma unroll
        for (int q = 0; q < kWarpCols / 4; ++q) {
          const float4 sc = *reinterpret_cast<const float4*>((BNT ? t_a : s_scale) + col_in_tile + q * 4);
          const float4 sh = *reinterpret_cast<const float4*>((BNT ? t_b : s_shift) + col_in_tile + q * 4);
          if (BNM == 2) {   // dz = a*dy + b*z + c: the b*z + c part (dy comes from the residual tile)
            const float4 sc2 = *reinterpret_cast<const float4*>(t_c + col_in_tile + q * 4);
            f[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]), sh.x, sc2.x);
            f[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]), sh.y, sc2.y);
            f[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]), sh.z, sc2.z);
            f[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]), sh.w, sc2.w);
          } else {
            f[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]), sc.x, sh.x);
            f[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]), sc.y, sh.y);
            f[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]), sc.z, sh.z);
            f[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]), sc.w, sh.w);
          }
        }
        if (HAS_RES && !BSTAT) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t j = (uint32_t)(half * 4 + q);  // 16-byte unit within the 128-byte row
            const uint4 r = *reinterpret_cast<const uint4*>(rbuf + row * 128 + ((j ^ sw) << 4));
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
            float ka[8];
            if (BNM == 2) {
              *reinterpret_cast<float4*>(ka) = *reinterpret_cast<const float4*>(t_a + col_in_tile + q * 8);
              *reinterpret_cast<float4*>(ka + 4) = *reinterpret_cast<const float4*>(t_a + col_in_tile + q * 8 + 4);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 p = unpack_bf16x2(w[t]);
              if (BNM == 2) {
                f[q * 8 + t * 2] = fmaf(p.x, ka[t * 2], f[q * 8 + t * 2]);
                f[q * 8 + t * 2 + 1] = fmaf(p.y, ka[t * 2 + 1], f[q * 8 + t * 2 + 1]);
              } else {
                f[q * 8 + t * 2] += p.x;
                f[q * 8 + t * 2 + 1] += p.y;
              }
            }
          }
        }
        if (BNM == 1 || MASK) {
          // packed ReLU masks: 32 consecutive channels of this row = one 32-bit word, staged in s_bits
          uint32_t* wp = s_bits + row * kBitW + (col_in_tile >> 5);
          if (BNM == 1) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) word |= (f[j] > 0.f ? 1u : 0u) << j;
            *wp = word;
          }
          if (MASK) {
            const uint32_t word = *wp;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = (word >> j) & 1u ? f[j] : 0.f;
          }
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < kWarpCols; ++j) f[j] = fmaxf(f[j], 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          if (OUT_F32) {
            o = make_uint4(__float_as_uint(f[q * 4]), __float_as_uint(f[q * 4 + 1]),
                           __float_as_uint(f[q * 4 + 2]), __float_as_uint(f[q * 4 + 3]));
          } else {
            o.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
            o.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
            o.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
            o.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
          }
          const uint32_t j = (uint32_t)(half * 4 + q);
          *reinterpret_cast<uint4*>(obuf + row * 128 + ((j ^ sw) << 4)) = o;
        }
        if (HAS_RES && !BSTAT) {
          mbar_arrive(&res_empty[rslot]);
          if (++rslot == C::kResSlots) { rslot = 0; rphase ^= 1; }
        }
        // (3) make the generic-proxy smem writes visible to the TMA store, then (4) store
        fence_proxy_async_smem();
        epi_bar_sync(2);
        if (tid_e == 0) {
          tma_store_4d(&args.tmap_out, obuf, n_tile * BLOCK_N + c * kChunkCols, tw * args.box_w,
                       th * args.box_h, args.k_splits > 1 ? split : tn * args.box_n);
          tma_store_commit();
        }
        if (ST) {
          // BatchNorm batch statistics of the chunk just staged (the bf16 values that go to HBM;
          // rows outside the tensor are exact zeros). Thread = (column pair, 16-row group):
          // conflict-free LDS.32 (a warp reads one whole 128-byte row per step). The eight row
          // groups are combined through a slot array and ONE owner thread per (view, stat, column)
          // adds into the CTA-resident sums -- no shared-memory atomics (fp32 atomicAdd on shared
          // memory is a CAS loop and tripled the time of the HBM-bound layers).
          // STATS: (sum y, sum y^2). BSTAT: (sum dy, sum dy*z) with z from the residual-ring tile.
          float a1x = 0.f, a1y = 0.f, a2x = 0.f, a2y = 0.f;  // view 0
          float b1x = 0.f, b1y = 0.f, b2x = 0.f, b2y = 0.f;  // view 1
          const uint8_t* src = obuf + st_rg * 16 * 128 + (st_cp & 3) * 4;
          const uint8_t* zsrc = rbuf + st_rg * 16 * 128 + (st_cp & 3) * 4;
          const uint32_t unit = (uint32_t)st_cp >> 2;
          if ((st_m0 ^ st_m1) == 0xFFFFu && (st_m0 == 0u || st_m1 == 0u)) {
            // warp-uniform fast path: all 16 rows are valid pixels of one view
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t o16 = (unit ^ (uint32_t)(i & 7)) << 4;
              const float2 z = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + i * 128 + o16));
              float2 w = z;
              if (BSTAT) w = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(zsrc + i * 128 + o16));
              a1x += z.x; a1y += z.y;
              a2x = fmaf(z.x, w.x, a2x); a2y = fmaf(z.y, w.y, a2y);
            }
            if (st_m1 != 0u) {
              b1x = a1x; b1y = a1y; b2x = a2x; b2y = a2y;
              a1x = a1y = a2x = a2y = 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t o16 = (unit ^ (uint32_t)(i & 7)) << 4;
              const float2 z = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(src + i * 128 + o16));
              float2 w = z;
              if (BSTAT) w = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(zsrc + i * 128 + o16));
              const float m0 = (float)((st_m0 >> i) & 1u), m1 = (float)((st_m1 >> i) & 1u);
              a1x = fmaf(m0, z.x, a1x); a1y = fmaf(m0, z.y, a1y);
              a2x = fmaf(m0 * z.x, w.x, a2x); a2y = fmaf(m0 * z.y, w.y, a2y);
              b1x = fmaf(m1, z.x, b1x); b1y = fmaf(m1, z.y, b1y);
              b2x = fmaf(m1 * z.x, w.x, b2x); b2y = fmaf(m1 * z.y, w.y, b2y);
            }
          }
          if (BSTAT) {   // the z tile has been read: hand its ring slot back
            mbar_arrive(&res_empty[rslot]);
            if (++rslot == C::kResSlots) { rslot = 0; rphase ^= 1; }
          }
          // slot[rg][q][64 columns], q = view*2 + stat
          float* slot = s_slot + st_rg * 256 + 2 * st_cp;
          *reinterpret_cast<float2*>(slot + 0) = make_float2(a1x, a1y);
          *reinterpret_cast<float2*>(slot + 64) = make_float2(a2x, a2y);
          *reinterpret_cast<float2*>(slot + 128) = make_float2(b1x, b1y);
          *reinterpret_cast<float2*>(slot + 192) = make_float2(b2x, b2y);
          epi_bar_sync(3);
          {
            const int q = tid_e >> 6, cl = tid_e & 63;   // owner of (view q>>1, stat q&1, column cl)
            float sum = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) sum += s_slot[g * 256 + q * 64 + cl];
            const int col = n_tile * BLOCK_N + c * kChunkCols + cl;
            if (col < args.n_total)
              s_stat[((size_t)(q >> 1) * args.n_total + col) * 2 + (q & 1)] += sum;
          }
          // the slots are rewritten only after the next chunk's epi_bar_sync(1)/(2): ordered
        }
      }
      if (BNM == 1 && args.bn_bits != nullptr) {
        // the tile's sign-mask words leave as whole sectors: 16 / 8 / 4 bytes per thread, rows contiguous
        epi_bar_sync(3);
        if (f_ok) {
          const uint32_t* src = s_bits + fr * kBitW + fq * kBitH;
          if (kBitH == 4) *reinterpret_cast<uint4*>(args.bn_bits + f_word) = *reinterpret_cast<const uint4*>(src);
          else if (kBitH == 2) *reinterpret_cast<uint2*>(args.bn_bits + f_word) = *reinterpret_cast<const uint2*>(src);
          else args.bn_bits[f_word] = *src;
        }
      }
    }
    if (tid_e == 0) tma_store_wait_all();
    if (ST) {
      epi_bar_sync(1);  // every thread's additions into s_stat are done
      if (BSTAT) {
        // sum dy*xhat = invstd * (sum dy*z - mean * sum dy); bn_a / bn_b carry mean / invstd [2][n_total]
        for (int i = tid_e; i < 2 * args.n_total; i += kEpiThreads) {
          const double s1 = (double)s_stat[2 * i], t = (double)s_stat[2 * i + 1];
          if (s1 != 0.0 || t != 0.0) {
            const double mu = (double)__ldg(args.bn_a + i), is = (double)__ldg(args.bn_b + i);
            atomicAdd(args.stat_acc + 2 * i, s1);
            atomicAdd(args.stat_acc + 2 * i + 1, is * (t - mu * s1));
          }
        }
      } else {
        for (int i = tid_e; i < 4 * args.n_total; i += kEpiThreads) {
          const float v = s_stat[i];
          if (v != 0.f) atomicAdd(args.stat_acc + i, (double)v);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
  if (ST) {
    __shared__ unsigned int s_ticket;
    bn_last_block_finalize<BSTAT>(args.stat_acc, args.stat_fin, args.n_total, 2, gridDim.x, &s_ticket,
                                  args.bn_a, args.bn_b);
  }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the tensor-bound layers with c_out % 256 == 0: two CTAs of a
// cluster (the two SMs of a TPC) own two consecutive 128-row M tiles and ONE 256-wide N tile; each
// loads its own A tile and HALF of the B tile (128 filter rows), the leader issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads both halves of B from both CTAs' shared
// memory and accumulates 128 x 256 in EACH CTA's TMEM. Per k-block a CTA now moves 32 KB through
// L2->SM instead of 48 KB -- these layers run at the L2->SM rate with one CTA per tile (ncu:
// l1tex__m_xbar2l1tex_read_bytes 13-14.6 TB/s), so that is a 1.5x lighter main loop.
//   producer (warp 0, both CTAs): TMA with .cta_group::2, completion bytes signalled on the
//       LEADER's full barrier (2 arrivals + 64 KB per stage);
//   MMA (warp 1, leader): waits the leader's full barrier, commits with multicast to the empty /
//       tmem_full barriers of both CTAs;
//   epilogue (warps 2-9, both CTAs): as in igemm_kernel on the CTA's own accumulator; releases it by
//       arriving on the leader's tmem_empty barrier (2 x 256 arrivals).
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct PairCfg {
  static constexpr int kHalfN = BLOCK_N / 2;                     // filter rows each CTA loads
  static constexpr int kBBytes = kHalfN * kBlockK * 2;           // 16 / 8 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;          // 32 / 24 KiB
  static constexpr int kStages = BLOCK_N == 256 ? 5 : 7;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmem = kStages * kStageBytes + 2 * kChunkBytes + 2 * BLOCK_N * 4 + 256 + 1024;
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
igemm_pair_kernel(const __grid_constant__ IgemmArgs args) {
  using PC = PairCfg<BLOCK_N>;
  constexpr int kPairStages = PC::kStages;
  constexpr int kPairHalfN = PC::kHalfN;
  constexpr int kPairBBytes = PC::kBBytes;
  constexpr int kPairStageBytes = PC::kStageBytes;
  constexpr int kChunkCols = 64, kChunks = BLOCK_N / kChunkCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kPairStages * kABytes;
  uint8_t* smem_out = smem_b + kPairStages * kPairBBytes;
  float* s_scale = reinterpret_cast<float*>(smem_out + 2 * kChunkBytes);
  float* s_shift = s_scale + BLOCK_N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + BLOCK_N);
  uint64_t* full_bar = bars;                     // [stages] used in the leader
  uint64_t* empty_bar = bars + kPairStages;      // [stages] both CTAs
  uint64_t* tmem_full = bars + 2 * kPairStages;  // [2] both CTAs
  uint64_t* tmem_empty = tmem_full + 2;          // [2] used in the leader
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.tmap_b);
    tma_prefetch_desc(&args.tmap_a[0]);
    tma_prefetch_desc(&args.tmap_out);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 2);   // one arrive.expect_tx per CTA of the pair
      mbar_init(&empty_bar[s], 1);  // the leader's multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, PC::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();
  griddep_launch();

  const int m_tiles = args.tiles_w * args.tiles_h * args.tiles_n;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_pairs = m_pairs * args.n_tiles;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int num_kb = args.num_taps * args.c_blocks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
        const int n_tile = pt % args.n_tiles;
        const int m_tile = 2 * (pt / args.n_tiles) + (int)rank;
        const int tw = m_tile % args.tiles_w;
        const int th = (m_tile / args.tiles_w) % args.tiles_h;
        const int tn = m_tile / (args.tiles_w * args.tiles_h);  // >= tiles_n for the odd tail: OOB
        const int ow0 = tw * args.box_w, oh0 = th * args.box_h, n0 = tn * args.box_n;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
          const uint32_t lead_full = mapa_u32(&full_bar[stage], 0);
          mbar_expect_tx_cluster(lead_full, kPairStageBytes);
          tma_load_4d_pair(smem_a + stage * kABytes, &args.tmap_a[args.tap_map[tap]], lead_full,
                           cb * kBlockK, ow0 + args.tap_dw[tap], oh0 + args.tap_dh[tap], n0);
          tma_load_2d_pair(smem_b + stage * kPairBBytes, &args.tmap_b, lead_full,
                           (args.tap_w[tap] * args.c_blocks + cb) * kBlockK,
                           n_tile * BLOCK_N + (int)rank * kPairHalfN);
          if (++cb == args.c_blocks) { cb = 0; ++tap; }
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (lane == 0) {
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * kABytes), 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * kPairBBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp != kResWarp) {
    // ------------------------------- epilogue (warps 2..9) ----------------------
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int tid_e = threadIdx.x - 64;
    const uint32_t sw = (uint32_t)(row & 7);
    constexpr int kWarpCols = kChunkCols / 2;  // 32
    int local = 0, cc = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int n_tile = pt % args.n_tiles;
      const int m_tile = 2 * (pt / args.n_tiles) + (int)rank;
      const int tw = m_tile % args.tiles_w;
      const int th = (m_tile / args.tiles_w) % args.tiles_h;
      const int tn = m_tile / (args.tiles_w * args.tiles_h);
      const bool tile_ok = m_tile < m_tiles;   // the odd tail tile of the last pair does not exist
      for (int i = tid_e; i < BLOCK_N; i += kEpiThreads) {
        const int col = n_tile * BLOCK_N + i;
        const bool ok = col < args.n_total;
        s_scale[i] = (ok && args.scale != nullptr) ? __ldg(args.scale + col) : 1.f;
        s_shift[i] = (ok && args.shift != nullptr) ? __ldg(args.shift + col) : 0.f;
      }
      mbar_wait_cluster(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c, ++cc) {
        uint8_t* obuf = smem_out + (cc & 1) * kChunkBytes;
        if (tid_e == 0) tma_store_wait_read<1>();
        epi_bar_sync(1);
        const int col_in_tile = c * kChunkCols + half * kWarpCols;
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + col_in_tile;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (c == kChunks - 1) {
          tc_fence_before_sync();
          mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));  // hand the accumulator back (leader)
        }
        float f[32];
#pragma unroll
        for (int q = 0; q < kWarpCols / 4; ++q) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + col_in_tile + q * 4);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + col_in_tile + q * 4);
          f[q * 4 + 0] = fmaf(__uint_as_float(v[q * 4 + 0]), sc.x, sh.x);
          f[q * 4 + 1] = fmaf(__uint_as_float(v[q * 4 + 1]), sc.y, sh.y);
          f[q * 4 + 2] = fmaf(__uint_as_float(v[q * 4 + 2]), sc.z, sh.z);
          f[q * 4 + 3] = fmaf(__uint_as_float(v[q * 4 + 3]), sc.w, sh.w);
        }
        if (args.relu) {
#pragma unroll
          for (int j = 0; j < kWarpCols; ++j) f[j] = fmaxf(f[j], 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
          o.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
          o.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
          o.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
          const uint32_t j = (uint32_t)(half * 4 + q);
          *reinterpret_cast<uint4*>(obuf + row * 128 + ((j ^ sw) << 4)) = o;
        }
        fence_proxy_async_smem();
        epi_bar_sync(2);
        if (tid_e == 0 && tile_ok) {
          tma_store_4d(&args.tmap_out, obuf, n_tile * BLOCK_N + c * kChunkCols, tw * args.box_w,
                       th * args.box_h, tn * args.box_n);
          tma_store_commit();
        }
      }
    }
    if (tid_e == 0) tma_store_wait_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer may still read this CTA's B half / signal its barriers
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, PC::kTmemCols);
  }
}

// Split-K fix-up: out[m, n] = act(scale[n] * sum_s ws[s][m][n] + shift[n]); the slices are added in
// the fixed order s = 0, 1, ... so the result does not depend on which CTA finished first.
template <typename TO>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long rows, int n_total,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                     TO* __restrict__ out, long long ld_out) {
  griddep_wait();
  griddep_launch();
  const int ng = n_total / 8;
  const long long total = rows * ng;
  const long long slice = rows * (long long)n_total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ng;
    const int c0 = (int)(i - r * ng) * 8;
    const float* src = ws + r * n_total + c0;
    float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
    for (int s = 1; s < splits; ++s) {
      const float4 b0 = *reinterpret_cast<const float4*>(src + s * slice);
      const float4 b1 = *reinterpret_cast<const float4*>(src + s * slice + 4);
      a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
      a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
    }
    float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sc = scale ? __ldg(scale + c0 + e) : 1.f, sh = shift ? __ldg(shift + c0 + e) : 0.f;
      f[e] = fmaf(f[e], sc, sh);
      if (relu) f[e] = fmaxf(f[e], 0.f);
    }
    TO* dst = out + r * ld_out + c0;
    if (sizeof(TO) == 2) {
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

// RMV_SPLITK: 1 (default) = split the reduction of small-M pointwise GEMMs when a workspace is given;
// 0 = never
int splitk_mode() { return tuning("SPLITK", 1, 1); }

// RMV_CTA2: 1 (default) = use the CTA-pair kernel where it applies; 0 = never; 2 = also force 256-wide N
// tiles for every c_out % 256 == 0 layer, however small (exercises the pair kernel in the tests)
int cta2_mode() { return tuning("CTA2", 1, 2); }

template <int BLOCK_N>
int launch_pair(const IgemmArgs& a, cudaStream_t stream) {
  constexpr int kPairSmem = PairCfg<BLOCK_N>::kSmem;
  static bool attr_set = false;
  if (!attr_set) {
    RMV_CUDA(cudaFuncSetAttribute(igemm_pair_kernel<BLOCK_N>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    attr_set = true;
  }
  // persistent kernel: exactly as many clusters as can be co-resident. That is NOT num_sms / 2:
  // a CTA pair needs both SMs of one TPC, and floor-swept parts have TPCs with a single SM -- with
  // 74 clusters on this B200 the leftover clusters ran as a second wave and doubled the time.
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / 2 * 2);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = kPairSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    RMV_CUDA(cudaOccupancyMaxActiveClusters(&n, igemm_pair_kernel<BLOCK_N>, &cfg));
    RMV_CHECK_ARG(n >= 1, "cta_group::2 kernel: no co-resident cluster possible");
    max_clusters = n;
  }
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_n;
  const int pairs = ((m_tiles + 1) / 2) * a.n_tiles;
  int clusters = max_clusters < pairs ? max_clusters : pairs;
  igemm_pair_kernel<BLOCK_N><<<2 * clusters, kNumThreads, kPairSmem, stream>>>(a);
  RMV_LAUNCH_CHECK();
  return 0;
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

}  // namespace

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  RMV_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

namespace {

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int BLOCK_N, int RES, bool OUT_F32, bool HALO = false, bool STATS = false, int BNM = 0>
int launch(const IgemmArgs& a, int total_tiles, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, RES, HALO, STATS, BNM>;
  // STATS: [2][n_total][2] floats behind the fixed regions (they replace the 1 KiB alignment slack
  // at the end of kSmemBytes, which the 1024-byte aligned base may consume: keep it as well)
  const int stat_bytes = (STATS || BNM == 4) ? (4 * a.n_total + 8 * 256) * (int)sizeof(float) + 1024 : 0;
  const int smem = C::kSmemBytes + stat_bytes;
  RMV_CHECK_ARG(smem <= 232448, "tcgen05 conv: %d bytes of shared memory needed (c_out=%d)", smem,
                a.n_total);
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    RMV_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N, RES, OUT_F32, HALO, STATS, BNM>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_bytes = smem;
  }
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  RMV_CUDA(launch_pdl_tc(igemm_kernel<BLOCK_N, RES, OUT_F32, HALO, STATS, BNM>, dim3(grid),
                         dim3(kNumThreads), smem, stream, a));
  return 0;
}

// RMV_HALO: 0 = never use the halo-patch 3x3 kernel; 1 (default) = use it. The row-shifted
// descriptors keep base offset 0: measured on B200, tcgen05 applies the 128-byte swizzle to the
// absolute shared-memory address (the same function TMA wrote the patch with), so a start address
// that is a multiple of 128 B but not of 1024 B needs no base offset. 2 = set the base offset to
// (addr >> 7) & 7 anyway (diagnostic: gives wrong results, kept to document the finding).
int halo_mode() { return tuning("HALO", 1, 2); }

template <int BLOCK_N>
int dispatch(const IgemmArgs& a, int total, bool has_res, bool out_f32, int bn_mode, cudaStream_t stream) {
  if (bn_mode == 1) {   // recomputed-BatchNorm forward apply (+ residual): HBM-bound, deep residual ring
    if (has_res) return launch<BLOCK_N, 1, false, false, false, 1>(a, total, stream);
    return launch<BLOCK_N, 0, false, false, false, 1>(a, total, stream);
  }
  if (bn_mode == 2) return launch<BLOCK_N, 1, false, false, false, 2>(a, total, stream);
  if (bn_mode == 4) return launch<BLOCK_N, 1, false, false, false, 4>(a, total, stream);
  if (bn_mode == 3) {   // plain epilogue (+ residual) with the packed ReLU mask applied: training data gradients
    if (has_res) return launch<BLOCK_N, 1, false, false, false, 3>(a, total, stream);
    return launch<BLOCK_N, 0, false, false, false, 3>(a, total, stream);
  }
  if (a.stat_acc != nullptr) return launch<BLOCK_N, 0, false, false, true>(a, total, stream);
  if (out_f32) return launch<BLOCK_N, 0, true>(a, total, stream);
  if (has_res) {
    // K >= 512 (8+ k-blocks per tile): the main loop needs the shared memory more than the residual ring
    // (RMV_RES_DEEPK: 0 = never, 1 = K >= 512, 2 = K >= 256)
    const int deepk = tuning("RES_DEEPK", 1, 2);
    if (BLOCK_N == 256 && deepk != 0 && a.num_taps * a.c_blocks >= (deepk == 2 ? 4 : 8))
      return launch<BLOCK_N, 2, false>(a, total, stream);
    return launch<BLOCK_N, 1, false>(a, total, stream);
  }
  return launch<BLOCK_N, 0, false>(a, total, stream);
}

}  // namespace

int conv_fwd_tc(const ConvArgs& p, cudaStream_t stream) { return conv_taps_tc(p, nullptr, stream); }

// `taps` == nullptr: the dense kh x kw filter of `p`. Otherwise (stride 1 only) an explicit list of
// taps: A is the dy/x box shifted by (dh, dw), B the filter tap `widx` of a [c_out][w_taps][c_in]
// tensor; p.kh/p.kw/p.pad are ignored (used by the stride-2 data gradient, see conv_dgrad_tc).
int conv_taps_tc(const ConvArgs& p, const TapList* taps, cudaStream_t stream) {
  const bool out_f32 = (p.y_dtype == RMV_DTYPE_F32);
  const int y_es = out_f32 ? 4 : 2;
  RMV_CHECK_ARG(p.c_in % kBlockK == 0, "tcgen05 conv: c_in=%d must be a multiple of 64", p.c_in);
  RMV_CHECK_ARG(p.c_out % 8 == 0, "tcgen05 conv: c_out=%d must be a multiple of 8", p.c_out);
  RMV_CHECK_ARG(p.stride == 1 || p.stride == 2, "tcgen05 conv: stride %d unsupported", p.stride);
  RMV_CHECK_ARG(p.kh * p.kw <= kMaxTaps, "tcgen05 conv: %dx%d filter too large", p.kh, p.kw);
  RMV_CHECK_ARG(taps == nullptr || (p.stride == 1 && taps->n >= 1 && taps->n <= kMaxTaps &&
                                    taps->w_taps >= 1),
                "tcgen05 conv: explicit tap lists need stride 1 and 1..%d taps", kMaxTaps);
  RMV_CHECK_ARG(p.x_sw % 8 == 0 && p.x_sh % 8 == 0 && p.x_sn % 8 == 0,
                "tcgen05 conv: input pixel strides must be multiples of 8 elements");
  RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(p.x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0,
                "tcgen05 conv: pointers must be 16-byte aligned");
  RMV_CHECK_ARG(p.y_sw % 8 == 0 && p.y_sh % 8 == 0 && p.y_sn % 8 == 0,
                "tcgen05 conv: output pixel strides must be multiples of 8 elements");
  RMV_CHECK_ARG(p.residual == nullptr || (p.r_sw % 8 == 0 && p.r_sh % 8 == 0 && p.r_sn % 8 == 0),
                "tcgen05 conv: residual pixel strides must be multiples of 8 elements");
  RMV_CHECK_ARG(!(out_f32 && p.residual != nullptr),
                "tcgen05 conv: residual add with fp32 output is not supported");

  IgemmArgs a;
  memset(&a, 0, sizeof(a));
  int out_w = p.out_w, out_h = p.out_h, n_img = p.n_img;
  long long x_sw = p.x_sw, x_sh = p.x_sh, x_sn = p.x_sn;
  long long y_sw = p.y_sw, y_sh = p.y_sh, y_sn = p.y_sn;
  long long r_sw = p.r_sw, r_sh = p.r_sh, r_sn = p.r_sn;
  int in_w = p.in_w, in_h = p.in_h;

  // 1x1 stride-1 over a dense pixel grid is a plain GEMM: flatten (n, h, w) into one axis.
  const bool pointwise = (taps == nullptr && p.kh == 1 && p.kw == 1 && p.stride == 1 && p.pad == 0);
  const bool x_dense = (p.x_sh == p.x_sw * p.in_w) && (p.x_sn == p.x_sh * p.in_h);
  const bool y_dense = (p.y_sh == p.y_sw * p.out_w) && (p.y_sn == p.y_sh * p.out_h);
  const bool r_dense =
      p.residual == nullptr || ((p.r_sh == p.r_sw * p.out_w) && (p.r_sn == p.r_sh * p.out_h));
  if (pointwise && x_dense && y_dense && r_dense) {
    out_w = in_w = p.n_img * p.in_h * p.in_w;
    out_h = in_h = 1;
    n_img = 1;
    x_sh = x_sw * in_w; x_sn = x_sh;
    y_sh = y_sw * out_w; y_sn = y_sh;
    r_sh = r_sw * out_w; r_sn = r_sh;
  }

  // 3x3 / stride 1 / pad 1 / 64 -> 64 channels (layer1 conv2 and its data gradient): halo-patch
  // kernel, 8 wide x 16 high tiles of one image, filters resident in shared memory.
  const bool halo = taps == nullptr && halo_mode() != 0 && p.kh == 3 && p.kw == 3 && p.stride == 1 &&
                    p.pad == 1 && p.c_in == 64 && p.c_out == 64 && !out_f32 &&
                    p.residual == nullptr && p.block_n == 0 && out_w >= 8 && out_h >= 16;
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
  if (halo) { best_w = 8; best_h = 16; best_n = 1; }
  a.box_w = best_w; a.box_h = best_h; a.box_n = best_n;
  a.tiles_w = ceil_div(out_w, a.box_w);
  a.tiles_h = ceil_div(out_h, a.box_h);
  a.tiles_n = ceil_div(n_img, a.box_n);

  // Activation tensor maps: one per input parity plane that some tap touches.
  const int s = p.stride;
  int plane_id[2][2] = {{-1, -1}, {-1, -1}};
  int n_planes = 0;
  a.num_taps = taps ? taps->n : p.kh * p.kw;
  const int loop_h = taps ? 1 : p.kh, loop_w = taps ? taps->n : p.kw;
  for (int r = 0; r < loop_h; ++r)
    for (int q = 0; q < loop_w; ++q) {
      const int t = r * loop_w + q;
      // dense filter: tap (r,q) reads input pixel (oh*s - pad + r, ow*s - pad + q)
      const int off_h = taps ? taps->dh[t] : r - p.pad, off_w = taps ? taps->dw[t] : q - p.pad;
      const int ph = (off_h % s + s) % s, pw = (off_w % s + s) % s;
      if (plane_id[ph][pw] < 0) {
        const int pl_w = (in_w - pw + s - 1) / s, pl_h = (in_h - ph + s - 1) / s;
        RMV_CHECK_ARG(pl_w > 0 && pl_h > 0, "tcgen05 conv: empty parity plane");
        cuuint64_t dims[4] = {(cuuint64_t)p.c_in, (cuuint64_t)pl_w, (cuuint64_t)pl_h,
                              (cuuint64_t)n_img};
        cuuint64_t strides[3] = {(cuuint64_t)(x_sw * s * 2), (cuuint64_t)(x_sh * s * 2),
                                 (cuuint64_t)(x_sn * 2)};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)a.box_w, (cuuint32_t)a.box_h,
                             (cuuint32_t)a.box_n};
        if (halo) { box[1] = kHaloW; box[2] = kHaloH; box[3] = 1; }  // one patch per tile
        const __nv_bfloat16* base =
            reinterpret_cast<const __nv_bfloat16*>(p.x) + ph * x_sh + pw * x_sw;
        int rc = encode_map(&a.tmap_a[n_planes], base, 4, dims, strides, box);
        if (rc) return rc;
        plane_id[ph][pw] = n_planes++;
      }
      a.tap_map[t] = (signed char)plane_id[ph][pw];
      a.tap_dh[t] = (signed char)floordiv(off_h, s);
      a.tap_dw[t] = (signed char)floordiv(off_w, s);
      a.tap_w[t] = (signed char)(taps ? taps->widx[t] : t);
    }

  const long m_tiles = (long)a.tiles_w * a.tiles_h * a.tiles_n;
  const int block_n = (p.block_n > 0) ? p.block_n
                      : (cta2_mode() == 2 && p.c_out % 256 == 0) ? 256   // test mode: force pairs
                      : (p.c_out <= 64) ? 64
                      : (p.c_out % 256 == 0 && m_tiles * (p.c_out / 256) >= 2L * num_sms()) ? 256
                      : 128;
  RMV_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "bad block_n %d", block_n);
  // CTA pairs (cta_group::2): 256-wide N tiles, no residual / fp32 / statistics / halo variant
  // measured: the 128-wide pair variant is 5-10 % SLOWER than the single-CTA kernel (N = 128 MMAs
  // are too short to amortise the pair hand-shakes) -- it only runs in the forced test mode
  const bool pair = cta2_mode() >= 1 && (block_n == 256 || (block_n == 128 && cta2_mode() == 2)) &&
                    p.block_n == 0 && !halo &&
                    !out_f32 && p.residual == nullptr && p.stat_acc == nullptr &&
                    p.bn_mode == 0 && p.mask_bits == nullptr &&
                    p.c_out % block_n == 0 && m_tiles >= 2;
  // Split-K for the small-M pointwise GEMMs (the lifter / fuser / head Linear layers at M = B*V rows and
  // their data gradients): with m_tiles * n_tiles well below the SM count every CTA would stream the
  // whole K axis alone; S splits fill the machine, their fp32 partial tiles go to the caller's
  // workspace and splitk_reduce_kernel adds them in a fixed order (bit-reproducible).
  int k_splits = 1, kb_per_split = a.num_taps * (p.c_in / kBlockK);
  {
    const int num_kb = a.num_taps * (p.c_in / kBlockK);
    const long mn = m_tiles * ceil_div(p.c_out, block_n);
    if (splitk_mode() != 0 && p.workspace != nullptr && taps == nullptr && !halo && !pair &&
        out_h == 1 && n_img == 1 && p.residual == nullptr && p.stat_acc == nullptr &&
        p.bn_mode == 0 && p.mask_bits == nullptr && num_kb >= 8 && mn > 0 && 2 * mn <= num_sms() &&
        p.c_out % 8 == 0 && y_sw % 8 == 0) {
      int s_want = (int)(num_sms() / mn);
      if (s_want > num_kb / 4) s_want = num_kb / 4;
      if (s_want > 16) s_want = 16;
      if (s_want >= 2) {
        const int per = ceil_div(num_kb, s_want);
        const int s_eff = ceil_div(num_kb, per);
        const size_t need = (size_t)s_eff * (size_t)out_w * (size_t)p.c_out * sizeof(float);
        if (s_eff >= 2 && need <= p.workspace_bytes &&
            (reinterpret_cast<uintptr_t>(p.workspace) & 15) == 0) {
          k_splits = s_eff; kb_per_split = per;
        }
      }
    }
  }
  const bool splitk = k_splits > 1;
  {
    const long long k_total = (long long)(taps ? taps->w_taps : a.num_taps) * p.c_in;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)p.c_out};
    cuuint64_t strides[1] = {(cuuint64_t)(k_total * 2)};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(pair ? block_n / 2 : block_n)};
    int rc = encode_map(&a.tmap_b, p.w, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    // Output (store) and residual (load) maps share the M-tile box; 128 bytes of channels per row.
    cuuint64_t dims[4] = {(cuuint64_t)p.c_out, (cuuint64_t)out_w, (cuuint64_t)out_h,
                          (cuuint64_t)n_img};
    cuuint32_t box[4] = {(cuuint32_t)(128 / y_es), (cuuint32_t)a.box_w, (cuuint32_t)a.box_h,
                         (cuuint32_t)a.box_n};
    cuuint64_t ystr[3] = {(cuuint64_t)(y_sw * y_es), (cuuint64_t)(y_sh * y_es),
                          (cuuint64_t)(y_sn * y_es)};
    int rc;
    if (splitk) {   // fp32 partial tiles: slice s of the workspace = [out_w rows][c_out]
      cuuint64_t wdims[4] = {(cuuint64_t)p.c_out, (cuuint64_t)out_w, 1, (cuuint64_t)k_splits};
      const cuuint64_t slice = (cuuint64_t)out_w * p.c_out * 4;
      cuuint64_t wstr[3] = {(cuuint64_t)p.c_out * 4, slice, slice};
      cuuint32_t wbox[4] = {32, (cuuint32_t)a.box_w, 1, 1};
      rc = encode_map(&a.tmap_out, p.workspace, 4, wdims, wstr, wbox, true);
    } else {
      rc = encode_map(&a.tmap_out, p.y, 4, dims, ystr, box, out_f32);
    }
    if (rc) return rc;
    if (p.residual != nullptr) {
      cuuint64_t rstr[3] = {(cuuint64_t)(r_sw * 2), (cuuint64_t)(r_sh * 2), (cuuint64_t)(r_sn * 2)};
      rc = encode_map(&a.tmap_res, p.residual, 4, dims, rstr, box);
      if (rc) return rc;
    }
  }
  a.n_total = p.c_out;
  a.n_tiles = ceil_div(p.c_out, block_n);
  a.c_blocks = p.c_in / kBlockK;
  a.scale = splitk ? nullptr : p.scale;
  a.shift = splitk ? nullptr : p.shift;
  a.relu = splitk ? 0 : p.relu;
  a.k_splits = k_splits;
  a.kb_per_split = kb_per_split;
  const int total = (int)(m_tiles * a.n_tiles) * k_splits;
  if (total == 0) return 0;
  if (splitk) {
    int rc = (block_n == 64)    ? launch<64, 0, true>(a, total, stream)
             : (block_n == 128) ? launch<128, 0, true>(a, total, stream)
                                : launch<256, 0, true>(a, total, stream);
    if (rc) return rc;
    const long long groups = (long long)out_w * (p.c_out / 8);
    long blocks = (long)((groups + 255) / 256);
    if (blocks > 8L * num_sms()) blocks = 8L * num_sms();
    if (out_f32) {
      RMV_CUDA(launch_pdl(splitk_reduce_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, stream,
                          (const float*)p.workspace, k_splits, (long long)out_w, p.c_out, p.scale,
                          p.shift, p.relu, (float*)p.y, y_sw));
    } else {
      RMV_CUDA(launch_pdl(splitk_reduce_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0,
                          stream, (const float*)p.workspace, k_splits, (long long)out_w, p.c_out,
                          p.scale, p.shift, p.relu, (__nv_bfloat16*)p.y, y_sw));
    }
    return 0;
  }
  if (pair) return block_n == 256 ? launch_pair<256>(a, stream) : launch_pair<128>(a, stream);
  if (p.stat_acc != nullptr && p.bn_mode != 4)
    RMV_CHECK_ARG(!out_f32 && p.residual == nullptr && p.stat_views == 2 && p.bn_mode == 0,
                  "tcgen05 conv: fused BatchNorm statistics need bf16 output, no residual, 2 views");
  if (p.stat_acc != nullptr || p.bn_mode != 0 || p.mask_bits != nullptr) {
    // row bookkeeping of the epilogue: image (view) and pixel of every tile row
    a.stat_acc = p.stat_acc;
    const bool flattened = (out_h == 1 && n_img == 1 && (p.out_h != 1 || p.n_img != 1));
    a.stat_pix = flattened ? p.out_h * p.out_w : 0;
    a.stat_rows = (long long)p.n_img * p.out_h * p.out_w;
    int sh = 0;
    while ((1 << sh) < a.box_w * a.box_h) ++sh;
    a.stat_ppi_shift = sh;
    sh = 0;
    while ((1 << sh) < a.box_w) ++sh;
    a.stat_bw_shift = sh;
    a.stat_w = out_w; a.stat_h = out_h; a.stat_n = n_img;
    a.m_sn = y_sn; a.m_sh = y_sh; a.m_sw = y_sw; a.mask_off = p.mask_off;
    if (p.stat_acc != nullptr && p.stat_finalize != nullptr) {
      const rmv_bn_params* fp = reinterpret_cast<const rmv_bn_params*>(p.stat_finalize);
      RMV_CHECK_ARG(fp->ticket != nullptr && p.n_img % 2 == 0 &&
                        (long long)(p.n_img / 2) * p.out_h * p.out_w > 1,
                    "tcgen05 conv: stat_finalize needs a ticket counter and an even image count");
      a.stat_fin = bn_finalize_args(fp, p.bn_mode == 4, (long long)(p.n_img / 2) * p.out_h * p.out_w);
    }
  }
  if (p.bn_mode != 0 || p.mask_bits != nullptr) {
    // mask words are moved in units of block_n / 64 words per thread: whole units per row
    const int unit = 32 * (block_n / 64);
    RMV_CHECK_ARG(!out_f32 && !halo && p.c_out % unit == 0 && p.y_sw % unit == 0 && p.y_sh % unit == 0 &&
                      p.y_sn % unit == 0 && p.mask_off % unit == 0,
                  "tcgen05 conv: BatchNorm-apply modes / ReLU masks need bf16 output and channel counts, "
                  "strides and mask offsets that are multiples of %d (whole mask units per row)", unit);
    RMV_CHECK_ARG(p.bn_mode == 0 || (p.bn_a != nullptr && p.bn_b != nullptr && p.scale == nullptr &&
                                     p.shift == nullptr && (p.mask_bits == nullptr || p.bn_mode == 4)),
                  "tcgen05 conv: bn_mode needs bn_a/bn_b, no scale/shift (mask_bits only with bn_mode 4)");
    RMV_CHECK_ARG(p.stat_acc == nullptr || p.bn_mode == 4,
                  "tcgen05 conv: statistics cannot be combined with bn_mode 1-3 / mask_bits");
    RMV_CHECK_ARG(p.bn_mode != 2 || (p.bn_c != nullptr && p.residual != nullptr && p.relu == 0),
                  "tcgen05 conv: bn_mode 2 needs bn_c and dy in `residual`, no ReLU");
    RMV_CHECK_ARG(p.bn_mode != 4 || (p.stat_acc != nullptr && p.residual != nullptr && p.mask_bits != nullptr &&
                                     p.relu == 0 && p.stat_views == 2 && block_n <= 256),
                  "tcgen05 conv: bn_mode 4 needs stat_acc (2 views), z in `residual`, mask_bits, no ReLU");
    RMV_CHECK_ARG(p.bn_mode >= 0 && p.bn_mode <= 4 && p.bn_mode != 3, "tcgen05 conv: bad bn_mode %d", p.bn_mode);
    a.bn_a = p.bn_a; a.bn_b = p.bn_b; a.bn_c = p.bn_c;
    a.bn_bits = reinterpret_cast<uint32_t*>(p.bn_bits);
    a.mask_bits = reinterpret_cast<const uint32_t*>(p.mask_bits);
    RMV_CHECK_ARG((reinterpret_cast<uintptr_t>(p.bn_bits) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.mask_bits) & 15) == 0,
                  "tcgen05 conv: mask pointers must be 16-byte aligned");
  }
  if (halo) {
    a.halo_base_mode = halo_mode() == 2;
    if (a.stat_acc != nullptr) return launch<64, 0, false, true, true>(a, total, stream);
    return launch<64, 0, false, true>(a, total, stream);
  }
  const bool has_res = p.residual != nullptr;
  const int epi_mode = p.bn_mode == 4 ? 4 : (p.mask_bits != nullptr ? 3 : p.bn_mode);
  switch (block_n) {
    case 64: return dispatch<64>(a, total, has_res, out_f32, epi_mode, stream);
    case 128: return dispatch<128>(a, total, has_res, out_f32, epi_mode, stream);
    default: return dispatch<256>(a, total, has_res, out_f32, epi_mode, stream);
  }
}

// Data gradient of y = conv(x, w, stride, pad) through the forward kernel (argument convention:
// rmv_conv2d_dgrad in the header -- p.x = dy, p.w = reversed/transposed filters, p.y = dx).
// stride 1: one conv with pad' = k-1-pad. stride 2: the four output parity classes (ih%2, iw%2)
// are independent stride-1 convolutions over the UNDILATED dy with 1/2/2/4 of the 9 taps each,
// written through TMA store maps with doubled pixel strides -- no zero-dilated copy of dy and a
// quarter of the MMA work of the dilated formulation.
int conv_dgrad_tc(const ConvArgs& p, cudaStream_t stream) {
  RMV_CHECK_ARG(p.kh == p.kw, "dgrad: square filters only");
  if (p.stride == 1) {
    ConvArgs q = p;
    q.pad = p.kh - 1 - p.pad;
    RMV_CHECK_ARG(q.pad >= 0 && p.out_h == p.in_h + 2 * q.pad - p.kh + 1 &&
                      p.out_w == p.in_w + 2 * q.pad - p.kw + 1,
                  "dgrad: dx size %dx%d inconsistent with dy %dx%d k%d p%d", p.out_h, p.out_w,
                  p.in_h, p.in_w, p.kh, p.pad);
    return conv_fwd_tc(q, stream);
  }
  RMV_CHECK_ARG(p.stride == 2, "dgrad: stride %d unsupported", p.stride);
  RMV_CHECK_ARG(p.bn_mode != 4, "dgrad: bn_mode 4 (fused BatchNorm backward reduction) needs stride 1");
  RMV_CHECK_ARG(p.kh * p.kw <= kMaxTaps, "dgrad: filter too large");
  const int y_es = p.y_dtype == RMV_DTYPE_F32 ? 4 : 2;
  TapList tl[2][2];
  bool any_empty = false;
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      TapList& t = tl[pa][pb];
      t.n = 0;
      t.w_taps = p.kh * p.kw;
      // forward tap (r, s) contributes to dx row ih = 2*oh - pad + r: for ih = 2i + pa the dy row is
      // oh = i + (pa + pad - r)/2 when that is an integer. The reversed filter tensor holds
      // w[., ., r, s] at tap (kh-1-r, kw-1-s).
      for (int r = 0; r < p.kh; ++r)
        for (int q = 0; q < p.kw; ++q) {
          if (((pa + p.pad - r) & 1) || ((pb + p.pad - q) & 1)) continue;
          t.dh[t.n] = (signed char)floordiv(pa + p.pad - r, 2);
          t.dw[t.n] = (signed char)floordiv(pb + p.pad - q, 2);
          t.widx[t.n] = (signed char)((p.kh - 1 - r) * p.kw + (p.kw - 1 - q));
          ++t.n;
        }
      const int sub_h = (p.out_h - pa + 1) / 2, sub_w = (p.out_w - pb + 1) / 2;
      if (t.n == 0 && sub_h > 0 && sub_w > 0) any_empty = true;
    }
  if (any_empty) {
    // some parity class receives no gradient (1x1 stride-2 filters): dx must read as zero there
    RMV_CHECK_ARG(p.residual == nullptr, "dgrad: residual with an empty parity class");
    const bool dense =
        p.y_sw == p.c_out && p.y_sh == p.y_sw * p.out_w && p.y_sn == p.y_sh * p.out_h;
    RMV_CHECK_ARG(dense, "dgrad: dx must be dense when a parity class is empty");
    RMV_CUDA(cudaMemsetAsync(p.y, 0, (size_t)p.n_img * p.out_h * p.out_w * p.c_out * y_es, stream));
  }
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      const int sub_h = (p.out_h - pa + 1) / 2, sub_w = (p.out_w - pb + 1) / 2;
      if (tl[pa][pb].n == 0 || sub_h <= 0 || sub_w <= 0) continue;
      ConvArgs q = p;
      q.stride = 1; q.pad = 0;
      q.out_h = sub_h; q.out_w = sub_w;
      q.y = (char*)p.y + ((long long)pa * p.y_sh + (long long)pb * p.y_sw) * y_es;
      q.y_sh = 2 * p.y_sh; q.y_sw = 2 * p.y_sw;
      if (p.residual) {
        q.residual = (const char*)p.residual + ((long long)pa * p.r_sh + (long long)pb * p.r_sw) * 2;
        q.r_sh = 2 * p.r_sh; q.r_sw = 2 * p.r_sw;
      }
      q.mask_off = p.mask_off + (long long)pa * p.y_sh + (long long)pb * p.y_sw;
      if (int rc = conv_taps_tc(q, &tl[pa][pb], stream)) return rc;
    }
  return 0;
}

}  // namespace rmv
