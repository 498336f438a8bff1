// Internal C++ declarations shared by the .cu translation units of librotmv_sm100.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/rotmv_sm100.h"

namespace rmv {

typedef rmv_conv_args ConvArgs;

// tcgen05 / TMEM / TMA implicit GEMM (igemm_sm100.cu). bf16 in, bf16 or fp32 out.
int conv_fwd_tc(const ConvArgs& p, cudaStream_t stream);
// FFMA tiled implicit GEMM (simt_conv.cu). fp32 or bf16 storage, fp32 accumulation.
int conv_fwd_simt(const ConvArgs& p, cudaStream_t stream);

// cuTensorMapEncodeTiled wrapper (128-byte swizzle, zero OOB fill); bf16 or fp32 elements.
int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, bool f32 = false);

}  // namespace rmv
