// Probe: how tcgen05.mma resolves a 128B-swizzled K-major A descriptor whose start address is NOT 1024-aligned
// (row-shifted start, arbitrary stride between 8-row groups).  Decides the halo-reuse layout of the conv kernel.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); } }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int AROWS = 512, NB = 32;

struct Cfg { int shift_rows; int sbo_bytes; int mode; };   // mode 0: base_offset 0; 1: (addr>>7)&7

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Cfg cfg, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // AROWS * 128 B
  uint8_t* sB = smem + AROWS * 128;         // NB * 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NB * 128);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars), bar1 = bar0 + 8;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar0, AROWS * 128 + NB * 128);
    for (int r = 0; r < AROWS; r += 256) tma_load_2d(smem_u32(sA) + r * 128, &map_a, bar0, 0, r);
    tma_load_2d(smem_u32(sB), &map_b, bar0, 0, 0);
    mbar_wait(bar0, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_addr = smem_u32(sA) + cfg.shift_rows * 128;
    const uint32_t boff = cfg.mode == 1 ? ((a_addr >> 7) & 7) : 0;
    const uint64_t adesc = make_desc(a_addr, 16, cfg.sbo_bytes, boff), bdesc = make_desc(smem_u32(sB), 16, 1024, 0);
    constexpr uint32_t idesc = make_idesc(128, NB);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(bar1);
  }
  mbar_wait(bar1, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < NB; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[m * NB + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  std::vector<__nv_bfloat16> hA(AROWS * 64), hB(NB * 64);
  std::vector<float> fA(AROWS * 64), fB(NB * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (float)((rand() % 13) - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 128 * NB * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ma, mb;
  { cuuint64_t dims[2] = {64, AROWS}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; } }
  { cuuint64_t dims[2] = {64, NB}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, NB}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; } }
  const size_t smem = AROWS * 128 + NB * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> hO(128 * NB);
  const int shifts[] = {0, 1, 3, 5, 8, 11, 19};
  const int sbos[] = {1024, 1152, 1280, 1408, 2048, 640, 896};
  for (int sbo : sbos) for (int sh : shifts) for (int mode = 0; mode < 2; ++mode) {
    Cfg c{sh, sbo, mode};
    CK(cudaMemset(dO, 0, 128 * NB * 4));
    probe<<<1, 128, smem>>>(ma, mb, c, dO);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    // expected under "rows are plain 128-B smem rows": row(m) = shift + (m/8)*(sbo/128) + m%8
    double err = 0; int bad_rows = 0;
    for (int m = 0; m < 128; ++m) {
      const int row = sh + (m / 8) * (sbo / 128) + (m % 8);
      double rerr = 0;
      for (int n = 0; n < NB; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += (double)fA[row * 64 + k] * fB[n * 64 + k];
        rerr = fmax(rerr, fabs(acc - hO[m * NB + n]));
      }
      if (rerr > 1e-3) ++bad_rows;
      err = fmax(err, rerr);
    }
    printf("sbo %4d shift %2d base_offset_mode %d : max_err %.4f bad_rows %d %s\n", sbo, sh, mode, err, bad_rows, bad_rows ? "MISMATCH" : "ok");
  }
  return 0;
}
