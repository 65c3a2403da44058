// Branch-light epilogue shared by the tcgen05 conv kernels: one 16-channel group of one output pixel.
//   f = act(acc + bias) * mask  ->  bf16 / fp32 store
// The epilogue warps run one warp per SM sub-partition, so every instruction is paid at full dependent-issue
// latency: bias comes from shared memory (broadcast LDS), the Dropout2d mask from registers, the activation is
// chosen once per group (uniform branch), and all loops are unrolled over independent elements.
#pragma once
#include "icf_tc_ptx.cuh"

namespace icf_tc {

// `o` points at channel `kbase` of the destination pixel; nvalid = channels of this group that exist (1..16)
__device__ __forceinline__ void epi16(const uint32_t (&v)[16], const float* sbias, const float (&mk)[16], int act,
                                      float slope, int nvalid, int out_f32, void* o) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + sbias[j];
  if (act == ICF_ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
  } else if (act == ICF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) f[j] = tanhf(f[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] *= mk[j];
  if (out_f32) {
    float* of = reinterpret_cast<float*>(o);
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(of) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(of + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) of[j] = f[j];
    }
  } else {
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(o);
    const bool al = (reinterpret_cast<uintptr_t>(ob) & 15) == 0;
    if (al && nvalid == 16) {
      uint4 a, b;
      a.x = pack_bf16(f[0], f[1]); a.y = pack_bf16(f[2], f[3]); a.z = pack_bf16(f[4], f[5]); a.w = pack_bf16(f[6], f[7]);
      b.x = pack_bf16(f[8], f[9]); b.y = pack_bf16(f[10], f[11]); b.z = pack_bf16(f[12], f[13]); b.w = pack_bf16(f[14], f[15]);
      *reinterpret_cast<uint4*>(ob) = a;
      *reinterpret_cast<uint4*>(ob + 8) = b;
    } else if (al && nvalid >= 8) {
      uint4 a;
      a.x = pack_bf16(f[0], f[1]); a.y = pack_bf16(f[2], f[3]); a.z = pack_bf16(f[4], f[5]); a.w = pack_bf16(f[6], f[7]);
      *reinterpret_cast<uint4*>(ob) = a;
#pragma unroll
      for (int j = 8; j < 16; ++j)
        if (j < nvalid) ob[j] = __float2bfloat16_rn(f[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) ob[j] = __float2bfloat16_rn(f[j]);
    }
  }
}

}  // namespace icf_tc
