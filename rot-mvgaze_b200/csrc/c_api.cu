// extern "C" surface of librotmv_sm100.so that is not tied to one kernel file: version, error
// string, device check and the convolution dispatcher.
#include "common.cuh"
#include "ops.h"

#include <stdarg.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>

namespace rmv {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Tuning switches: environment variable RMV_<KEY> (one digit), overridable at run time through
// rmv_set_tuning (tests flip kernel variants inside one process).
static std::mutex g_tune_mu;
static std::map<std::string, int> g_tune;

int tuning(const char* key, int dflt, int max_value) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  auto it = g_tune.find(key);
  if (it != g_tune.end()) return it->second;
  const std::string env = std::string("RMV_") + key;
  const char* e = getenv(env.c_str());
  const int v = (e != nullptr && e[0] >= '0' && e[0] <= '0' + max_value && e[1] == 0) ? e[0] - '0' : dflt;
  g_tune[key] = v;
  return v;
}

int pdl_level() { return tuning("PDL", 1, 2); }

}  // namespace rmv

using namespace rmv;

extern "C" int rmv_version(void) { return RMV_VERSION; }
extern "C" const char* rmv_last_error(void) { return last_error(); }

extern "C" int rmv_device_check(int device) {
  int major = 0, minor = 0;
  RMV_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  RMV_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  RMV_CHECK_ARG(major == 10, "device %d is sm_%d%d; librotmv_sm100 contains sm_100a code only",
                device, major, minor);
  return 0;
}

extern "C" int rmv_conv2d_fwd(const rmv_conv_args* args, void* stream) {
  RMV_CHECK_ARG(args != nullptr, "conv2d_fwd: null args");
  const ConvArgs& p = *args;
  RMV_CHECK_ARG(p.x && p.w && p.y, "conv2d_fwd: null tensor pointer");
  RMV_CHECK_ARG(p.n_img >= 0 && p.in_h > 0 && p.in_w > 0 && p.c_in > 0 && p.c_out > 0,
                "conv2d_fwd: bad shape");
  RMV_CHECK_ARG(p.out_h == (p.in_h + 2 * p.pad - p.kh) / p.stride + 1 &&
                    p.out_w == (p.in_w + 2 * p.pad - p.kw) / p.stride + 1,
                "conv2d_fwd: out size %dx%d inconsistent with in %dx%d k%dx%d s%d p%d", p.out_h,
                p.out_w, p.in_h, p.in_w, p.kh, p.kw, p.stride, p.pad);
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc_ok = p.x_dtype == RMV_DTYPE_BF16 && p.c_in % 64 == 0 && p.c_out % 8 == 0 &&
                     (p.x_sc == 0 || p.x_sc == 1) && (p.stride == 1 || p.stride == 2);
  if (p.engine == RMV_ENGINE_TC) {
    RMV_CHECK_ARG(tc_ok, "conv2d_fwd: shape/dtype not supported by the tcgen05 engine");
    return conv_fwd_tc(p, s);
  }
  if (p.engine == RMV_ENGINE_AUTO && tc_ok) return conv_fwd_tc(p, s);
  // the FFMA engine has no statistics epilogue: refuse instead of leaving the accumulator empty
  RMV_CHECK_ARG(p.stat_acc == nullptr,
                "conv2d_fwd: fused BatchNorm statistics (stat_acc) need the tcgen05 engine "
                "(bf16, c_in %% 64 == 0)");
  return conv_fwd_simt(p, s);
}

extern "C" int rmv_conv2d_dgrad(const rmv_conv_args* args, void* stream) {
  RMV_CHECK_ARG(args != nullptr, "conv2d_dgrad: null args");
  const ConvArgs& p = *args;
  RMV_CHECK_ARG(p.x && p.w && p.y, "conv2d_dgrad: null tensor pointer");
  RMV_CHECK_ARG(p.n_img >= 0 && p.in_h > 0 && p.in_w > 0 && p.c_in > 0 && p.c_out > 0,
                "conv2d_dgrad: bad shape");
  RMV_CHECK_ARG(p.in_h == (p.out_h + 2 * p.pad - p.kh) / p.stride + 1 &&
                    p.in_w == (p.out_w + 2 * p.pad - p.kw) / p.stride + 1,
                "conv2d_dgrad: dy %dx%d is not the output of a k%d s%d p%d conv over %dx%d", p.in_h,
                p.in_w, p.kh, p.stride, p.pad, p.out_h, p.out_w);
  RMV_CHECK_ARG(p.x_dtype == RMV_DTYPE_BF16 && p.c_in % 64 == 0 && p.c_out % 8 == 0 &&
                    (p.x_sc == 0 || p.x_sc == 1),
                "conv2d_dgrad: tcgen05 engine needs bf16, c(dy) %% 64 == 0, c(dx) %% 8 == 0");
  RMV_CHECK_ARG(p.scale == nullptr && p.shift == nullptr && p.relu == 0,
                "conv2d_dgrad: no scale/shift/relu epilogue");
  return conv_dgrad_tc(p, (cudaStream_t)stream);
}

extern "C" size_t rmv_stem_wgrad_workspace_bytes(void) { return (size_t)192 * 64 * sizeof(float); }

// split-K only runs when (m, n) tiles x splits <= SMs: at most one 128 x 256 fp32 tile per SM
extern "C" size_t rmv_splitk_workspace_bytes(void) {
  return (size_t)rmv::num_sms() * 128 * 256 * sizeof(float);
}

extern "C" size_t rmv_bn_workspace_bytes(int max_channels, int views) {
  if (max_channels <= 0 || views <= 0) return 0;
  return (size_t)views * (size_t)max_channels * 2 * sizeof(double);
}

extern "C" size_t rmv_conv2d_wgrad_tc_workspace_bytes(const rmv_conv_args* args) {
  if (args == nullptr || args->c_out <= 0 || args->c_in <= 0 || args->kh <= 0 || args->kw <= 0) return 0;
  return (size_t)args->c_out * args->kh * args->kw * args->c_in * sizeof(float);
}

extern "C" int rmv_set_tuning(const char* key, int value) {
  RMV_CHECK_ARG(key != nullptr && value >= 0 && value <= 9, "set_tuning: bad key/value");
  std::lock_guard<std::mutex> lk(rmv::g_tune_mu);
  rmv::g_tune[key] = value;
  return 0;
}
