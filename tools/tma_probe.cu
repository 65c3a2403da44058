// Probe: TMA streaming rate of 4-D boxes of 128-byte rows as a function of how the rows lie in global memory
// (contiguous runs vs one row per image), ring depth and box size.  One persistent CTA per SM, a producer thread
// issuing cp.async.bulk.tensor and a consumer thread that only releases the slots.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 4000000000LL) { printf("timeout\n"); __trap(); } }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Job {
  int n_slots; uint32_t box_bytes; int boxes_per_slot;
  int t1, t2, t3;          // tiles along dims 1..3 (box-sized steps), total work items = t1*t2*t3
  int b1, b2, b3;          // box extents (for coordinate steps)
};

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, Job j, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t slot_bytes = j.box_bytes * j.boxes_per_slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)j.n_slots * slot_bytes);
  const uint32_t bb = smem_u32(bars);
  if (threadIdx.x == 0) { for (int s = 0; s < 2 * j.n_slots; ++s) mbar_init(bb + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int total = j.t1 * j.t2 * j.t3 / j.boxes_per_slot;
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x) {
      mbar_wait(bb + 8 * (j.n_slots + s), ph ^ 1);
      mbar_expect_tx(bb + 8 * s, slot_bytes);
      for (int b = 0; b < j.boxes_per_slot; ++b) {
        int w = it * j.boxes_per_slot + b;
        const int i1 = w % j.t1; w /= j.t1;
        const int i2 = w % j.t2; const int i3 = w / j.t2;
        tma_load_4d(smem_u32(smem) + s * slot_bytes + b * j.box_bytes, &map, bb + 8 * s, 0, i1 * j.b1, i2 * j.b2, i3 * j.b3);
      }
      if (++s == j.n_slots) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0; unsigned long long acc = 0;
    for (int it = blockIdx.x; it < total; it += gridDim.x) {
      mbar_wait(bb + 8 * s, ph);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + s * slot_bytes);
      mbar_arrive(bb + 8 * (j.n_slots + s));
      if (++s == j.n_slots) { s = 0; ph ^= 1; }
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  // activation tensor [N=4096][H=24][W=24][C=64] bf16 = 302 MB (larger than L2)
  const int N = 4096, H = 24, W = 24, C = 64;
  const size_t bytes = (size_t)N * H * W * C * 2;
  void* d; CK(cudaMalloc(&d, bytes)); CK(cudaMemset(d, 1, bytes));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  struct Case { const char* name; int order; int b1, b2, b3; int slots; int boxes; int prom; int swz; };
  // order 0: dims (C, W, H, N)  [NHWC natural]; order 1: dims (C, N, W, H) [image-innermost box rows]
  const Case cases[] = {
    {"NHWC  box 64x 8w x 8h x 2n  (16KB, 1KB runs)   slots 4", 0, 8, 8, 2, 4, 1, 2, 3},
    {"NHWC  box 64x 8w x 8h x 2n  (16KB, 1KB runs)   slots 8", 0, 8, 8, 2, 8, 1, 2, 3},
    {"NHWC  box 64x24w x 5h x 1n  (15KB, 3KB runs)   slots 8", 0, 24, 5, 1, 8, 1, 2, 3},
    {"NHWC  box 64x24w x 1h x 4n  (12KB, 3KB runs)   slots 8", 0, 24, 1, 4, 8, 1, 2, 3},
    {"NHWC  box 64x 1w x 1h x128n (16KB, 128B rows)  slots 8", 0, 1, 1, 128, 8, 1, 2, 3},
    {"CNWH  box 64x16n x 8w x 1h  (16KB, 128B rows)  slots 4", 1, 16, 8, 1, 4, 1, 2, 3},
    {"CNWH  box 64x16n x 8w x 1h  (16KB, 128B rows)  slots 8", 1, 16, 8, 1, 8, 1, 2, 3},
    {"CNWH  box 64x16n x 8w x 1h  x2 boxes/slot      slots 6", 1, 16, 8, 1, 6, 2, 2, 3},
    {"CNWH  box 64x16n x 8w x 1h  no L2 promotion    slots 8", 1, 16, 8, 1, 8, 1, 0, 3},
    {"CNWH  box 64x16n x 8w x 1h  128B promotion     slots 8", 1, 16, 8, 1, 8, 1, 1, 3},
    {"CNWH  box 64x16n x 8w x 1h  no swizzle         slots 8", 1, 16, 8, 1, 8, 1, 2, 0},
    {"CNWH  box 64x128n x 1w x 1h (16KB, 128B rows)  slots 8", 1, 128, 1, 1, 8, 1, 2, 3},
    {"CNWH  box 64x 8n x 8w x 2h  (16KB)             slots 8", 1, 8, 8, 2, 8, 1, 2, 3},
    {"CNWH  box 64x 4n x 8w x 1h  (4KB) x4 boxes     slots 8", 1, 4, 8, 1, 8, 4, 2, 3},
  };
  for (const Case& c : cases) {
    CUtensorMap m;
    cuuint64_t dims[4], str[3]; cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    int ext1, ext2, ext3;
    if (c.order == 0) {
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N;
      str[0] = C * 2; str[1] = (cuuint64_t)W * C * 2; str[2] = (cuuint64_t)H * W * C * 2;
      ext1 = W; ext2 = H; ext3 = N;
    } else {
      dims[0] = C; dims[1] = N; dims[2] = W; dims[3] = H;
      str[0] = (cuuint64_t)H * W * C * 2; str[1] = C * 2; str[2] = (cuuint64_t)W * C * 2;
      ext1 = N; ext2 = W; ext3 = H;
    }
    box[0] = 64; box[1] = c.b1; box[2] = c.b2; box[3] = c.b3;
    const CUtensorMapL2promotion prom = c.prom == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (c.prom == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    const CUtensorMapSwizzle swz = c.swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    Job j;
    j.n_slots = c.slots; j.box_bytes = 128u * c.b1 * c.b2 * c.b3; j.boxes_per_slot = c.boxes;
    j.t1 = ext1 / c.b1; j.t2 = ext2 / c.b2; j.t3 = ext3 / c.b3; j.b1 = c.b1; j.b2 = c.b2; j.b3 = c.b3;
    const size_t smem = (size_t)j.n_slots * j.box_bytes * j.boxes_per_slot + 1024 + 256;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double moved = (double)j.t1 * j.t2 * j.t3 * j.box_bytes;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid : {148, 296}) {
      if (grid == 296 && smem > 110 * 1024) continue;
      stream_kernel<<<grid, 64, smem>>>(m, j, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      stream_kernel<<<grid, 64, smem>>>(m, j, sink);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("%-60s grid %3d: %7.3f ms  %7.1f GB/s  (%.0f MB)\n", c.name, grid, ms, moved / ms / 1e6, moved / 1e6);
    }
  }
  return 0;
}
