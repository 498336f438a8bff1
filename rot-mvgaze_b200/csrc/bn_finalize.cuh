// BatchNorm(train) finalize steps shared by every reduction kernel of librotmv_sm100: the LAST block
// of a reduction launch (ticket counter) turns the fp64 per-(view, channel) sums into the coefficients
// of the apply pass, so statistics + finalize are one launch (train_kernels.cu: bn_reduce_kernel;
// bnconv_sm100.cu: tstat_kernel; igemm_sm100.cu: STATS epilogue).
#pragma once
#include <string.h>

#include "common.cuh"
#include "ops.h"

namespace rmv {

// Arguments of the finalize step the LAST block of a reduction launch runs (null ticket = none).
struct BnFinalize {
  unsigned int* ticket;  // zero on entry; the block that draws the last ticket finalizes + resets
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;
  float* mean; float* invstd; float* a; float* b;            // forward outputs
  float* dgamma; float* dbeta; float* k0; float* k1; float* k2;  // backward outputs
  double count; float eps, momentum;
};

// Finalize, written for a SINGLE block with plenty of memory-level parallelism (it is the serial
// tail of the reduction launch): phase 1 is one item per (view, channel), four items in flight per
// thread; phase 2 (running statistics in VIEW ORDER / dgamma, dbeta) is one item per channel.
// Forward: mean / invstd, fused affine (a = gamma*invstd, b = beta - mean*a), running stats
// (rm <- (1-m) rm + m mean_v for v = 0..V-1; unbiased variance), accumulator reset.
__device__ __forceinline__ void bn_finalize_block(double* acc, const BnFinalize& f, int c, int views,
                                                  int tid, int nthreads) {
  const int items = views * c;
  const double inv_count = 1.0 / f.count;
  for (int i0 = tid; i0 < items; i0 += nthreads * 4) {
    double s1[4], s2[4];
    float ga[4], be[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * nthreads;
      if (i < items) {
        s1[j] = __ldcg(acc + 2 * (long long)i);
        s2[j] = __ldcg(acc + 2 * (long long)i + 1);
        ga[j] = __ldg(f.gamma + i % c);
        be[j] = __ldg(f.beta + i % c);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * nthreads;
      if (i < items) {
        const double m = s1[j] * inv_count;
        double var = s2[j] * inv_count - m * m;
        if (var < 0.0) var = 0.0;
        const float is = (float)(1.0 / sqrt(var + (double)f.eps));
        f.mean[i] = (float)m;
        f.invstd[i] = is;
        const float av = ga[j] * is;
        f.a[i] = av;
        f.b[i] = be[j] - (float)m * av;
        acc[2 * (long long)i] = m;       // parked for phase 2
        acc[2 * (long long)i + 1] = var;
      }
    }
  }
  __syncthreads();
  const double unb = f.count / (f.count - 1.0);
  for (int ch = tid; ch < c; ch += nthreads) {
    float rm = f.running_mean ? f.running_mean[ch] : 0.f, rv = f.running_var ? f.running_var[ch] : 0.f;
    for (int v = 0; v < views; ++v) {
      double* p = acc + ((long long)v * c + ch) * 2;
      const float m = (float)p[0], unbiased = (float)(p[1] * unb);
      p[0] = 0.0; p[1] = 0.0;
      rm = (1.f - f.momentum) * rm + f.momentum * m;
      rv = (1.f - f.momentum) * rv + f.momentum * unbiased;
    }
    if (f.running_mean) f.running_mean[ch] = rm;
    if (f.running_var) f.running_var[ch] = rv;
  }
}

// Backward: dgamma/dbeta (=), per-(v,c) coefficients for the apply pass
//   dz = k0 * dyr + k1 * z + k2  with  k0 = gamma*invstd, k1 = -k0*invstd*s2/cnt,
//   k2 = -k0*s1/cnt - k1*mean   (from dz = gamma*invstd*(dyr - s1/cnt - xhat*s2/cnt))
__device__ __forceinline__ void bn_bwd_finalize_block(double* acc, const BnFinalize& f,
                                                      const float* mean, const float* invstd, int c,
                                                      int views, int tid, int nthreads) {
  const int items = views * c;
  const double inv_count = 1.0 / f.count;
  for (int i0 = tid; i0 < items; i0 += nthreads * 4) {
    double s1[4], s2[4];
    float ga[4], is[4], mu[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * nthreads;
      if (i < items) {
        s1[j] = __ldcg(acc + 2 * (long long)i);
        s2[j] = __ldcg(acc + 2 * (long long)i + 1);
        ga[j] = __ldg(f.gamma + i % c);
        is[j] = invstd[i];
        mu[j] = mean[i];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * nthreads;
      if (i < items) {
        const double c0 = (double)ga[j] * (double)is[j];
        const double c1 = -c0 * (double)is[j] * s2[j] * inv_count;
        f.k0[i] = (float)c0;
        f.k1[i] = (float)c1;
        f.k2[i] = (float)(-c0 * s1[j] * inv_count - c1 * (double)mu[j]);
      }
    }
  }
  __syncthreads();
  for (int ch = tid; ch < c; ch += nthreads) {
    double dg = 0.0, db = 0.0;
    for (int v = 0; v < views; ++v) {
      double* p = acc + ((long long)v * c + ch) * 2;
      db += __ldcg(p); dg += __ldcg(p + 1);
      p[0] = 0.0; p[1] = 0.0;
    }
    f.dgamma[ch] = (float)dg;
    f.dbeta[ch] = (float)db;
  }
}


// Host: finalize arguments from the public parameter block (include/rotmv_sm100.h).
inline BnFinalize bn_finalize_args(const rmv_bn_params* p, bool bwd, long long count_per_view) {
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  if (p == nullptr) return fin;
  fin.ticket = p->ticket; fin.gamma = p->gamma; fin.beta = p->beta;
  fin.running_mean = p->running_mean; fin.running_var = p->running_var; fin.nbt = p->num_batches;
  fin.mean = p->mean; fin.invstd = p->invstd; fin.a = p->a; fin.b = p->b;
  fin.dgamma = p->dgamma; fin.dbeta = p->dbeta; fin.k0 = p->k0; fin.k1 = p->k1; fin.k2 = p->k2;
  fin.count = (double)count_per_view; fin.eps = p->eps; fin.momentum = p->momentum;
  (void)bwd;
  return fin;
}

#ifdef __CUDACC__
// Called by EVERY thread of the block after its contributions to `acc` have been issued. Returns
// after the finalize when this block drew the last of `n_blocks` tickets.
template <bool BWD>
__device__ __forceinline__ void bn_last_block_finalize(double* acc, const BnFinalize& fin, int c,
                                                       int views, unsigned n_blocks,
                                                       unsigned int* s_ticket,
                                                       const float* mean = nullptr,
                                                       const float* invstd = nullptr) {
  if (fin.ticket == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_ticket = atomicAdd(fin.ticket, 1u);
  __syncthreads();
  if (*s_ticket != n_blocks - 1) return;
  __threadfence();
  if (BWD) bn_bwd_finalize_block(acc, fin, mean, invstd, c, views, threadIdx.x, blockDim.x);
  else bn_finalize_block(acc, fin, c, views, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) {
    *fin.ticket = 0;
    if (!BWD && fin.nbt != nullptr) *fin.nbt += views;
  }
}
#endif

}  // namespace rmv
