// Weight gradient of a unit-stride FIRST conv layer whose input is the 8-channel-pitch feature tensor (sm_100a):
//   dw[a][r][s*8 + c] += sum_{n,p,q} dY[n,p,q,a] * X[n, p+r, q+s, c]        (folded "win" packing, c < 8)
// as tcgen05 MMAs D[a][(s,c)] += A^T[pix][a] . B[pix][(s,c)] with the PIXEL index as the reduction dimension and both
// operands MN-major in the un-swizzled ("interleaved") canonical layout, whose 8x8 core matrix keeps consecutive
// reduction rows 16 bytes apart.  With 16 bytes per pixel that is exactly an NHWC row of X: reduction row q, column
// group s  ->  pixel q+s, i.e. the overlapping filter windows are expressed by the DESCRIPTOR strides (SBO = 16 B
// between column groups, LBO = 128 B between 8-pixel groups).  X is fetched once per image for all R*S taps and no
// im2col copy exists anywhere; dY arrives as [row][channel group][q][8 ch] through a 5-D tensor map.  One accumulator
// per filter row r lives in TMEM for the whole kernel; each CTA adds its partial sums to dw once, at the end.
//
// The per-tap kernel (wgrad_tc_kernel) re-fetches both operands once per filter row with 128-byte rows that are
// 5/8 padding; for D.dx.1 (5->32, 5x5, 28x28) that was the most expensive launch of the train step.
#include "icf_tc_ptx.cuh"

#include <cstring>

namespace {

using namespace icf_tc;

constexpr int PX_ISSUERS = 4;            // filter row r is issued by warp 1 + r % PX_ISSUERS (one accumulator, one thread)
constexpr int PX_THREADS = 32 * (1 + PX_ISSUERS + 4);   // warp 0 TMA, issuers (warp 1 owns TMEM), last four warps flush
constexpr int PX_SLOTS = 3;

struct PxParams {
  int N, P, Q, A, R, B;                  // B = win*8 valid columns
  int PB, pblocks;                       // rows of dY per work item, work items per image
  int ksteps;                            // 16-pixel reduction steps per image row
  int QP, WP;                            // q extent of the dY tile, x extent of the X tile (elements of 16 B)
  uint32_t dy_bytes, x_bytes, slot_bytes;
  uint32_t n_cols;                       // MMA N (multiple of 16 >= B)
  uint32_t tmem_cols;
  float* dw;
};

__device__ __forceinline__ void tma_load_5d_if(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                               int c3, int c4, uint32_t leader) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %8, 0;\n"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
      "}\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(leader)
      : "memory");
}

// un-swizzled MN-major operand: `sbo` = bytes between 8-element column groups, `lbo` = bytes between 8-row reduction groups
__device__ __forceinline__ uint64_t make_desc_interleaved(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(PX_THREADS, 1) wgrad_px8_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                  const __grid_constant__ CUtensorMap map_x,
                                                                  const __grid_constant__ PxParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)PX_SLOTS * p.slot_bytes + 8192);   // 8 KB read-past guard
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PX_SLOTS + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PX_SLOTS + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * PX_SLOTS);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
    for (int s = 0; s < PX_SLOTS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), PX_ISSUERS);
    }
    mbar_init(done_bar, PX_ISSUERS);
    fence_barrier_init();
  }
  // the A operand is read 128 columns wide although only p.A channels exist: make sure whatever lies behind the
  // tiles is finite (garbage accumulator rows are never written back, but NaN patterns must not trap anything)
  for (int i = threadIdx.x; i < 2048; i += PX_THREADS) reinterpret_cast<uint32_t*>(smem + (size_t)PX_SLOTS * p.slot_bytes)[i] = 0u;
  if (warp == 1) tmem_alloc_rt(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();             // warp-uniform producer loop, the elected lane issues
      int s = 0;
      uint32_t ph = 0;
      const int items = p.N * p.pblocks;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int n = it / p.pblocks, p0 = (it - n * p.pblocks) * p.PB;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx_if(full_bar(s), p.dy_bytes + p.x_bytes, leader);
        const uint32_t base = smem_base + (uint32_t)s * p.slot_bytes;
        tma_load_5d_if(base, &map_dy, full_bar(s), 0, 0, 0, p0, n, leader);        // rows beyond P arrive as zeros
        tma_load_4d_if(base + p.dy_bytes, &map_x, full_bar(s), 0, 0, p0, n, leader);
        if (++s == PX_SLOTS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp <= PX_ISSUERS) {
    // whole warp runs the loop, the elected lane issues
    const int wi = warp - 1;
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(128, (int)p.n_cols, 1, 1);
    const uint64_t dA = make_desc_interleaved(0, 128, (uint32_t)p.QP * 16u);   // column groups = channel groups of dY
    const uint64_t dB = make_desc_interleaved(0, 128, 16);                     // column groups = consecutive pixels of X
    const uint32_t a_hi = (uint32_t)(dA >> 32), a_lo0 = (uint32_t)dA;
    const uint32_t b_hi = (uint32_t)(dB >> 32), b_lo0 = (uint32_t)dB;
    const uint32_t row_a16 = (uint32_t)((p.A >> 3) * p.QP);                    // 16-byte units per dY row
    int s = 0;
    uint32_t ph = 0;
    uint32_t first = 0;                                                         // accumulate flag (0 only for the very first step)
    const int items = p.N * p.pblocks;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_img = ((smem_base + (uint32_t)s * p.slot_bytes) >> 4) & 0x3FFFu;
      const uint32_t b_img = ((smem_base + (uint32_t)s * p.slot_bytes + p.dy_bytes) >> 4) & 0x3FFFu;
      const int p0 = (it % p.pblocks) * p.PB;
      const int rows = p.P - p0 < p.PB ? p.P - p0 : p.PB;
      for (int pr = 0; pr < rows; ++pr) {
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t a_lo = a_lo0 | (a_img + (uint32_t)pr * row_a16 + (uint32_t)ks * 16u);
          for (int r = wi; r < p.R; r += PX_ISSUERS) {
            const uint32_t b_lo = b_lo0 | (b_img + (uint32_t)((pr + r) * p.WP + ks * 16));
            umma_bf16_lo2(tmem_base + (uint32_t)r * p.n_cols, a_lo, a_hi, b_lo, b_hi, idesc, first, leader);
          }
          first = 1u;
        }
      }
      umma_commit_if(empty_bar(s), leader);
      if (++s == PX_SLOTS) { s = 0; ph ^= 1; }
    }
    umma_commit_if(done_bar, leader);
  } else {
    // final flush: TMEM lane a = output channel
    const int q4 = warp & 3;
    const int a = q4 * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if ((int)blockIdx.x < p.N * p.pblocks) {
      for (int r = 0; r < p.R; ++r)
        for (int c0 = 0; c0 < (int)p.n_cols; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(r * p.n_cols + c0), v);
          tmem_ld_wait();
          if (a < p.A) {
            float* o = p.dw + ((int64_t)a * p.R + r) * p.B + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.B) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_rt(tmem_base, p.tmem_cols);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_plain(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* str, const cuuint32_t* box) {
  EncodeFn fn = reinterpret_cast<EncodeFn>(get_encode());
  ICF_REQUIRE(fn, "first-layer wgrad: cuTensorMapEncodeTiled is unavailable");
  cuuint32_t est[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, str, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICF_REQUIRE(r == CUDA_SUCCESS, "first-layer wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------------
// Row-stacked variant (A = 32 or 64 channels of dY): JB = 128/A rows of dY share the 128 MMA rows,
//   D_delta[(j,a)][(s,c)] += sum_q dY[n][p0+j][q][a] * X[n][p0+delta][q+s][c],     delta = 0 .. JB+R-2,   filter row r = delta - j
// so one item (image, block of JB rows) is (JB+R-1) x ksteps MMAs instead of JB x R x ksteps, none of them 3/4 padding.
// dY arrives in its NATIVE layout — [q][A channels] rows of 64 or 128 bytes, MN-major with the matching swizzle, the JB row
// tiles as the descriptor's M groups — instead of the [channel group][q][8 ch] transposition the kernel above asks of TMA:
// that 5-D box is moved as one request per 16 bytes (3072 per image, plus 1064 for the X tile with its 16-byte inner box:
// 114 k requests per SM and launch, 3.7 cycles each = the whole 222 us).  Here a dY row costs one 64/128-byte request and an X
// row (pixels and channels merged into one box dimension) one request of W*16 bytes.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int PS_ISSUERS = 4;
constexpr int PS_THREADS = 32 * (1 + PS_ISSUERS + 4);
constexpr int PS_MAX_SLOTS = 16;         // ring depth is chosen at launch (as many ~10 KB items as fit ~190 KB).  Depth 6 -> 16 changed nothing
                                         // (104 us): the issuers wait on `full` because the dY boxes arrive as one TMA request per 64-byte row
                                         // (16 k requests per SM and launch, ~8 cycles each) — the kernel is TMA-request bound, not latency bound

struct PsParams {
  int N, P, A, R, B;
  int JB, D, pblocks, ksteps, QP, W;
  int slots;                             // ring depth
  uint32_t swb;                          // bytes per dY row = A*2 (64 or 128) = swizzle span
  uint32_t dy_bytes, x_row_bytes, slot_bytes;
  uint32_t n_cols, tmem_cols;
  float* dw;
};

__global__ void __launch_bounds__(PS_THREADS, 1) wgrad_px8s_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                   const __grid_constant__ CUtensorMap map_x,
                                                                   const __grid_constant__ PsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.slots * p.slot_bytes + 4096);   // 4 KB read-past guard
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PS_MAX_SLOTS + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PS_MAX_SLOTS + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * PS_MAX_SLOTS);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
    for (int s = 0; s < p.slots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), PS_ISSUERS);
    }
    mbar_init(done_bar, PS_ISSUERS);
    fence_barrier_init();
  }
  // the X windows of the last reduction rows run past their row (into the next one, the next slot or the guard): whatever
  // lies there is multiplied by the zero-filled q >= Q part of dY, but it has to be finite from the first MMA on
  for (uint32_t i = threadIdx.x; i < ((uint32_t)p.slots * p.slot_bytes + 4096u) / 16u; i += PS_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc_rt(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const int items = p.N * p.pblocks;

  if (warp == 0) {
    const uint32_t leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int n = it / p.pblocks, p0 = (it - n * p.pblocks) * p.JB;
      mbar_wait(empty_bar(s), ph ^ 1);
      mbar_expect_tx_if(full_bar(s), p.dy_bytes + (uint32_t)p.D * p.x_row_bytes, leader);
      const uint32_t base = smem_base + (uint32_t)s * p.slot_bytes;
      tma_load_4d_if(base, &map_dy, full_bar(s), 0, 0, p0, n, leader);                 // rows beyond P / q beyond Q arrive as zeros
      tma_load_3d_if(base + p.dy_bytes, &map_x, full_bar(s), 0, p0, n, leader);
      if (++s == p.slots) { s = 0; ph ^= 1; }
    }
  } else if (warp <= PS_ISSUERS) {
    const int wi = warp - 1;
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(128, (int)p.n_cols, 1, 1);
    // A: MN-major, swizzle = row width; M groups (the JB row tiles) QP*swb apart (LBO), 8-pixel reduction groups 8*swb apart (SBO)
    const uint64_t dA = make_desc_sw(0, (uint32_t)p.QP * p.swb, 8u * p.swb, p.swb == 128u ? 2u : 4u);
    const uint64_t dB = make_desc_interleaved(0, 128, 16);                             // column groups = consecutive pixels of X
    const uint32_t a_hi = (uint32_t)(dA >> 32), a_lo0 = (uint32_t)dA;
    const uint32_t b_hi = (uint32_t)(dB >> 32), b_lo0 = (uint32_t)dB;
    int s = 0;
    uint32_t ph = 0, first = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_img = ((smem_base + (uint32_t)s * p.slot_bytes) >> 4) & 0x3FFFu;
      const uint32_t b_img = ((smem_base + (uint32_t)s * p.slot_bytes + p.dy_bytes) >> 4) & 0x3FFFu;
      for (int d = wi; d < p.D; d += PS_ISSUERS) {                                     // accumulator delta belongs to ONE issuer
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t a_lo = a_lo0 | (a_img + (uint32_t)ks * p.swb);                // 16 reduction rows = 16*swb bytes
          const uint32_t b_lo = b_lo0 | (b_img + (uint32_t)d * (p.x_row_bytes >> 4) + (uint32_t)ks * 16u);
          umma_bf16_lo2(tmem_base + (uint32_t)d * p.n_cols, a_lo, a_hi, b_lo, b_hi, idesc, first | (uint32_t)ks, leader);
        }
      }
      first = 1u;
      umma_commit_if(empty_bar(s), leader);
      if (++s == p.slots) { s = 0; ph ^= 1; }
    }
    umma_commit_if(done_bar, leader);
  } else {
    // final flush: TMEM lane = (stacked row j, channel a); accumulator delta holds filter row r = delta - j
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int j = m / p.A, a = m - j * p.A;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if ((int)blockIdx.x < items) {
      // Every CTA adds the same A*R*B sums at the same moment.  In accumulator order each address received a burst of one
      // atomic per CTA (ncu: the flush took as long as the main loop); each CTA therefore starts at its own (accumulator,
      // column group) and the sums leave as 16-byte vector reductions.
      const int groups = (int)p.n_cols / 16, units = p.D * groups;
      const bool vec = (p.B & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
      for (int uu = 0; uu < units; ++uu) {
        const int u = (uu + (int)blockIdx.x) % units;
        const int d = u / groups, c0 = (u - d * groups) * 16;
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(d * p.n_cols + c0), v);
        tmem_ld_wait();
        const int r = d - j;
        if (r >= 0 && r < p.R) {
          float* o = p.dw + ((int64_t)a * p.R + r) * p.B + c0;
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            if (vec && c0 + t + 3 < p.B) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + t), "f"(__uint_as_float(v[t])),
                           "f"(__uint_as_float(v[t + 1])), "f"(__uint_as_float(v[t + 2])), "f"(__uint_as_float(v[t + 3])) : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c0 + t + e < p.B) atomicAdd(o + t + e, __uint_as_float(v[t + e]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_rt(tmem_base, p.tmem_cols);
}

int encode_swz(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* str, const cuuint32_t* box,
               CUtensorMapSwizzle swz) {
  EncodeFn fn = reinterpret_cast<EncodeFn>(get_encode());
  ICF_REQUIRE(fn, "first-layer wgrad: cuTensorMapEncodeTiled is unavailable");
  cuuint32_t est[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, str, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICF_REQUIRE(r == CUDA_SUCCESS, "first-layer wgrad: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// returns 0 = launched, -1 = not this variant's case, >0 = error
int launch_px8s(const icf_wgrad_args* a, cudaStream_t st) {
  if ((a->A != 32 && a->A != 64) || a->a_pitch != a->A || a->W * 8 > 256) return -1;
  static const bool off = []() { const char* e = getenv("ICF_DISABLE_PX8S"); return e && e[0] && e[0] != '0'; }();
  if (off) return -1;
  PsParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.P = a->P; p.A = a->A; p.R = a->R; p.B = a->B; p.W = a->W;
  p.JB = 128 / a->A;
  p.D = p.JB + a->R - 1;
  p.pblocks = icf::cdiv(a->P, p.JB);
  p.ksteps = icf::cdiv(a->Q, 16);
  p.QP = 16 * p.ksteps;
  p.swb = (uint32_t)a->A * 2u;
  p.n_cols = (uint32_t)((a->B + 15) & ~15);
  if (p.QP > 256 || (uint32_t)p.D * p.n_cols > 512) return -1;
  p.dy_bytes = (uint32_t)p.JB * (uint32_t)p.QP * p.swb;
  p.x_row_bytes = (uint32_t)a->W * 16u;
  p.slot_bytes = (p.dy_bytes + (uint32_t)p.D * p.x_row_bytes + 1023u) & ~1023u;
  const uint32_t cols = (uint32_t)p.D * p.n_cols;
  p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  // the widest read past an X row: reduction row QP-1, column group n_cols/8-1
  if ((uint32_t)(p.QP + (int)p.n_cols / 8) * 16u > p.x_row_bytes + 4096u) return -1;
  p.slots = (int)((190u * 1024u) / p.slot_bytes);
  if (p.slots > PS_MAX_SLOTS) p.slots = PS_MAX_SLOTS;
  if (p.slots < 2) return -1;
  const size_t smem = (size_t)p.slots * p.slot_bytes + 4096 + 1024 + 512;
  if (smem > 225 * 1024) return -1;
  p.dw = a->dw;
  CUtensorMap mdy, mx;
  {
    // dY [N][P][Q][A]: box = (A channels, QP pixels, JB rows, 1 image) -> [j][q][A] with 64/128-byte swizzled rows
    cuuint64_t dims[4] = {(cuuint64_t)a->A, (cuuint64_t)a->Q, (cuuint64_t)a->P, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->a_pitch * 2, (cuuint64_t)a->Q * a->a_pitch * 2, (cuuint64_t)a->P * a->Q * a->a_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)a->A, (cuuint32_t)p.QP, (cuuint32_t)p.JB, 1};
    if (int r = encode_swz(&mdy, a->small_t, 4, dims, str, box, p.swb == 128u ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
      return r;
  }
  {
    // X [N][H][W][8]: pixels and channels merged into the innermost dimension, box = (W*8, D rows, 1 image)
    cuuint64_t dims[3] = {(cuuint64_t)a->W * 8, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[2] = {(cuuint64_t)a->W * 16, (cuuint64_t)a->H * a->W * 16};
    cuuint32_t box[3] = {(cuuint32_t)a->W * 8, (cuuint32_t)p.D, 1};
    if (int r = encode_swz(&mx, a->big_t, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return r;
  }
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(wgrad_px8s_kernel), smem, "first-layer wgrad (stacked)")) return r;
  const int sms = icf::sm_count();
  const int64_t items = (int64_t)a->N * p.pblocks;
  const int grid = items < sms ? (int)items : sms;
  wgrad_px8s_kernel<<<grid, PS_THREADS, smem, st>>>(mdy, mx, p);
  return icf::check_launch("wgrad_px8s");
}

}  // namespace

// returns 0 = launched, -1 = not this kernel's case, >0 = error
int icf_px8_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st) {
  if (a->dtype != ICF_BF16 || a->win < 2 || a->S != 1 || a->stride != 1 || a->pad != 0 || a->b_pitch != 8) return -1;
  if (a->B != a->win * 8 || (a->A & 7) || a->A > 128 || (a->a_pitch & 7) || a->R > 8) return -1;
  if ((reinterpret_cast<uintptr_t>(a->small_t) & 15) || (reinterpret_cast<uintptr_t>(a->big_t) & 15)) return -1;
  if (a->P + a->R - 1 > a->H || a->Q + a->win - 1 > a->W) return -1;
  static const bool off = []() { const char* e = getenv("ICF_DISABLE_PX8"); return e && e[0] && e[0] != '0'; }();
  if (off) return -1;
  {
    int r = launch_px8s(a, st);
    if (r >= 0) return r;
  }
  PxParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.P = a->P; p.Q = a->Q; p.A = a->A; p.R = a->R; p.B = a->B;
  p.ksteps = icf::cdiv(a->Q, 16);
  p.QP = 16 * p.ksteps;
  p.n_cols = (uint32_t)((a->B + 15) & ~15);
  p.WP = p.QP + (int)p.n_cols / 8;                 // widest pixel a column group of the last reduction row touches
  if (p.QP > 256 || p.WP > 256) return -1;
  const int row_bytes = (a->A / 8) * p.QP * 16;
  p.PB = (56 * 1024) / row_bytes;                  // rows of dY per work item: keep a ring slot near 64 KB
  if (p.PB < 1) return -1;
  if (p.PB > a->P) p.PB = a->P;
  if (p.PB + a->R - 1 > 256) p.PB = 256 - (a->R - 1);
  p.pblocks = icf::cdiv(a->P, p.PB);
  p.dy_bytes = (uint32_t)(p.PB * row_bytes);
  p.x_bytes = (uint32_t)((p.PB + a->R - 1) * p.WP) * 16u;
  p.slot_bytes = (p.dy_bytes + p.x_bytes + 1023u) & ~1023u;
  const uint32_t cols = (uint32_t)a->R * p.n_cols;
  if (cols > 512) return -1;
  p.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : (cols <= 256 ? 256 : 512)));
  const size_t smem = (size_t)PX_SLOTS * p.slot_bytes + 8192 + 1024 + 256;
  if (smem > 225 * 1024) return -1;
  p.dw = a->dw;

  CUtensorMap mdy, mx;
  {
    // dY [N][P][Q][a_pitch] viewed as (8 ch, q, channel group, p, n): the tile lands as [p][group][q][8 ch]
    cuuint64_t dims[5] = {8, (cuuint64_t)a->Q, (cuuint64_t)(a->A / 8), (cuuint64_t)a->P, (cuuint64_t)a->N};
    cuuint64_t str[4] = {(cuuint64_t)a->a_pitch * 2, 16, (cuuint64_t)a->Q * a->a_pitch * 2,
                         (cuuint64_t)a->P * a->Q * a->a_pitch * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)p.QP, (cuuint32_t)(a->A / 8), (cuuint32_t)p.PB, 1};
    if (int r = encode_plain(&mdy, a->small_t, 5, dims, str, box)) return r;
  }
  {
    cuuint64_t dims[4] = {8, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[3] = {16, (cuuint64_t)a->W * 16, (cuuint64_t)a->H * a->W * 16};
    cuuint32_t box[4] = {8, (cuuint32_t)p.WP, (cuuint32_t)(p.PB + a->R - 1), 1};
    if (int r = encode_plain(&mx, a->big_t, 4, dims, str, box)) return r;
  }
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(wgrad_px8_kernel), smem, "first-layer wgrad")) return r;
  const int sms = icf::sm_count();
  const int64_t items = (int64_t)a->N * p.pblocks;
  const int grid = items < sms ? (int)items : sms;
  wgrad_px8_kernel<<<grid, PX_THREADS, smem, st>>>(mdy, mx, p);
  return icf::check_launch("wgrad_px8");
}
