// Shared helpers for libicf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "icf.h"

namespace icf {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define ICF_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::icf::set_error(__VA_ARGS__);           \
      return 1;                                \
    }                                          \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// SM count of the CURRENT device (queried once per device; 148 on B200) — grid sizing never hard-codes it
int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize for one kernel: the attribute is per device and the library is re-entrant
// (Python thread + torch's autograd worker thread, icf.h), so the high-water mark is kept per device under a mutex
struct SmemGuard {
  static constexpr int MAX_DEV = 64;
  size_t configured[MAX_DEV];
  SmemGuard();
  int ensure(const void* kernel, size_t smem, const char* what);
};
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- typed load/store with fp32 math ------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// runtime-dtype element access (dtype: ICF_F32 / ICF_BF16)
__device__ __forceinline__ float ld_any(const void* base, int dtype, int64_t idx) {
  return dtype == ICF_F32 ? reinterpret_cast<const float*>(base)[idx]
                          : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}
__device__ __forceinline__ void st_any(void* base, int dtype, int64_t idx, float v) {
  if (dtype == ICF_F32) reinterpret_cast<float*>(base)[idx] = v;
  else reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ICF_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == ICF_ACT_TANH) return tanhf(v);
  return v;
}
// derivative of the activation expressed through the saved *output* y (LeakyReLU keeps the sign)
__device__ __forceinline__ float act_grad_from_output(float y, int act, float slope) {
  if (act == ICF_ACT_LRELU) return y > 0.f ? 1.f : slope;
  if (act == ICF_ACT_TANH) return 1.f - y * y;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum (blockDim.x <= 1024); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

}  // namespace icf

// implemented in icf_conv_tc.cu; returns 0 = done, -1 = shape not supported by the tensor-core path
// (caller falls through to the SIMT kernels), >0 = error.
int icf_tc_conv_forward(const icf_conv_args* a, cudaStream_t stream);
// implemented in icf_conv_ws.cu (weight-stationary row-streaming kernel); same return convention
int icf_ws_conv_forward(const icf_conv_args* a, cudaStream_t stream);
// implemented in icf_conv_sc.cu (scatter-form transposed conv, few output channels); same return convention
int icf_sc_conv_forward(const icf_conv_args* a, cudaStream_t stream);
// implemented in icf_conv_cm.cu (unit-stride first layer, channel-major accumulator); same return convention
int icf_cm_conv_forward(const icf_conv_args* a, cudaStream_t stream);
int icf_launch_col_stats(const void* y, int ydt, int ypitch, int64_t pixels, int C, float* stats, cudaStream_t st);
int icf_tc_conv_wgrad(const icf_wgrad_args* a, cudaStream_t stream);
// per-tap kernel with split-K into a zeroed fp32 scratch (+ finishing pass); -1 when splitting would not help
int icf_tc_conv_forward_splitk(const icf_conv_args* a, cudaStream_t stream, float* partial, int64_t partial_elems);
