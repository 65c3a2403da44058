// Branch-light epilogue shared by the tcgen05 conv kernels: one 16-channel group of one output pixel.
//   f = act(acc + bias) * mask  ->  bf16 / fp32 store
// The epilogue warps run one warp per SM sub-partition, so every instruction is paid at full dependent-issue
// latency: bias comes from shared memory (broadcast LDS), the Dropout2d mask from registers, the activation is
// chosen once per group (uniform branch), and all loops are unrolled over independent elements.
#pragma once
#include "icf_tc_ptx.cuh"

namespace icf_tc {

// LeakyReLU in two instructions (FMUL + FMNMX) for the slopes the models use (0 <= slope <= 1: max(x, slope*x));
// the generic select otherwise.  NaN propagates through both.
__device__ __forceinline__ void lrelu16(float (&f)[16], float slope) {
  if (slope >= 0.f && slope <= 1.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], f[j] * slope);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
  }
}

// `o` points at channel `kbase` of the destination pixel; nvalid = channels of this group that exist (1..16);
// npad >= nvalid = channels that may be WRITTEN (the pitch padding up to the next multiple of 8 is stored as zeros
// so that a ragged channel count still leaves as 16-byte stores)
__device__ __forceinline__ void epi16(const uint32_t (&v)[16], const float* sbias, const float (&mk)[16], int act,
                                      float slope, int nvalid, int npad, int out_f32, void* o) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + sbias[j];
  if (act == ICF_ACT_LRELU) {
    lrelu16(f, slope);
  } else if (act == ICF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) f[j] = tanhf(f[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = j < nvalid ? f[j] * mk[j] : 0.f;
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
    if (al && npad == 16) {
      uint4 a, b;
      a.x = pack_bf16(f[0], f[1]); a.y = pack_bf16(f[2], f[3]); a.z = pack_bf16(f[4], f[5]); a.w = pack_bf16(f[6], f[7]);
      b.x = pack_bf16(f[8], f[9]); b.y = pack_bf16(f[10], f[11]); b.z = pack_bf16(f[12], f[13]); b.w = pack_bf16(f[14], f[15]);
      *reinterpret_cast<uint4*>(ob) = a;
      *reinterpret_cast<uint4*>(ob + 8) = b;
    } else if (al && npad >= 8) {
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

// Same math, result packed as bf16 into two 16-byte chunks (channels 0-7, 8-15; zeros beyond nvalid) for the
// shared-memory staged TMA-store epilogue.
__device__ __forceinline__ void epi16_pack(const uint32_t (&v)[16], const float* sbias, const float (&mk)[16], int act,
                                           float slope, int nvalid, uint4& lo, uint4& hi) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + sbias[j];
  if (act == ICF_ACT_LRELU) {
    lrelu16(f, slope);
  } else if (act == ICF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) f[j] = tanhf(f[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = j < nvalid ? f[j] * mk[j] : 0.f;
  lo.x = pack_bf16(f[0], f[1]); lo.y = pack_bf16(f[2], f[3]); lo.z = pack_bf16(f[4], f[5]); lo.w = pack_bf16(f[6], f[7]);
  hi.x = pack_bf16(f[8], f[9]); hi.y = pack_bf16(f[10], f[11]); hi.z = pack_bf16(f[12], f[13]); hi.w = pack_bf16(f[14], f[15]);
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace icf_tc
