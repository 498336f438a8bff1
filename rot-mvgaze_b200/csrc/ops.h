// Internal C++ declarations shared by the .cu translation units of librotmv_sm100.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/rotmv_sm100.h"

namespace rmv {

typedef rmv_conv_args ConvArgs;

// Explicit tap list of an implicit GEMM: tap t reads the input box shifted by (dh, dw) pixels and
// multiplies it with filter tap widx[t] of a [c_out][w_taps][c_in] tensor.
struct TapList {
  int n, w_taps;
  signed char dh[49], dw[49], widx[49];
};

// tcgen05 / TMEM / TMA implicit GEMM (igemm_sm100.cu). bf16 in, bf16 or fp32 out.
int conv_fwd_tc(const ConvArgs& p, cudaStream_t stream);
int conv_taps_tc(const ConvArgs& p, const TapList* taps, cudaStream_t stream);
// data gradient (stride 1 or 2) through the same kernel; see rmv_conv2d_dgrad
int conv_dgrad_tc(const ConvArgs& p, cudaStream_t stream);
// FFMA tiled implicit GEMM (simt_conv.cu). fp32 or bf16 storage, fp32 accumulation.
int conv_fwd_simt(const ConvArgs& p, cudaStream_t stream);

// cuTensorMapEncodeTiled wrapper (128-byte swizzle, zero OOB fill); bf16 or fp32 elements.
int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, bool f32 = false);

}  // namespace rmv
