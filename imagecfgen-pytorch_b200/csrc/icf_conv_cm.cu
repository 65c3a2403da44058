// Unit-stride FIRST conv layer on the 8-channel-pitch feature tensor, few output channels (D.dx.1: 5 -> 32, 5x5, 28x28 -> 24x24),
// with the roles of the MMA operands swapped (sm_100a):  ACCUMULATOR LANE = (output row j, output channel k), COLUMN = pixel.
//
// The row-streaming kernel computes this layer as out[pixel][k] with the 128 MMA rows = pixels and N = 32 channels: 15
// tcgen05.mma of N = 32 per 128 output pixels (5 filter rows x 3 K-steps of the folded 40-channel window).  Every MMA costs the
// tensor pipe >= 32 cycles for reading its 128-row A operand from shared memory and the issuing warp ~70 cycles whatever N is,
// and the x-im2col is built by TMA (a 96-byte window per pixel: L2->SM 445 MB for a 52 MB input, ncu) — 139-152 us per launch
// against 28 us of HBM time and 16 us of tensor time.
//
// Here the weights are the A operand and JB = 128/K output rows share one MMA:
//   D[(j,k)][(i,x)] += sum_kk A_d[(j,k)][kk] * B_d[(i,x)][kk],    d = 0 .. JB+R-2 source rows of the block
//   A_d[(j,k)][kk] = w[k][r = d - j][kk]  (zero rows where r is outside the filter),  kk = s*8 + c (folded window)
//   B_d[(i,x)][kk] = in[image i][row y0 + d][pixel x + kk/8][channel kk%8]
// B is the RAW NHWC row of the 16-byte pixels: in the un-swizzled canonical K-major layout a core matrix is 8 rows 16 bytes
// apart, so "row n starts 16 bytes after row n-1" (SBO = 128 B per 8 rows, LBO = 16 B per 8-element K chunk) IS the overlapping
// window — no im2col anywhere, every input byte is fetched from L2 once per row block ((JB+R-1)/JB = 2x).  One work item =
// (block of JB output rows, NI images): JB+R-1 source rows x ceil(win*8/16) K-steps = 24 MMAs of M = 128, N = NI*W = 224 per
// 768 output pixels (was 120 of N = 32), 112 cycles of tensor time each, so the issue cost disappears behind the tensor pipe.
// Accumulator columns whose x lies in the last W-Q slots of an image row are garbage and never read.
//
// Epilogue: a TMEM lane is one (row, channel), so bias / Dropout2d mask are per-thread registers, the BatchNorm statistics are two
// registers per thread for the whole kernel (one atomicAdd pair per thread at the end), and a warp stores the 32 channels of
// one pixel as one 64-byte run.  16 epilogue warps (four per TMEM lane quarter, split over the images), double-buffered
// accumulator (2 x 256 columns), 3-slot TMA ring, persistent grid of one CTA per SM.
#include "icf_tc_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace {

using namespace icf_tc;

constexpr int CM_STAGES = 3;
constexpr int CM_EPI_WARPS = 16;
constexpr int CM_MAX_IMG = 4;                           // 32-column units per epilogue warp and item
constexpr int CM_THREADS = 32 * (3 + CM_EPI_WARPS);     // warp 0 TMA, warp 1 MMA (owns TMEM), 16 epilogue warps, last warp = second MMA issuer
constexpr int CM_ACC_COLS = 256;

struct CmParams {
  int N, H, W, P, Q, K, R;
  int JB, D, KS;                          // output rows per block, source rows per block, 16-element K-steps of the window
  int stride;
  int NI, ncols;                          // images per item, MMA N = NI*W
  int row_blocks, img_groups;
  int w_pitch, out_pitch, mask_pitch;
  int act;
  float slope;
  uint32_t a_bytes;                       // one A_d matrix
  uint32_t d_bytes, stage_bytes;          // one source row of NI images, one ring slot
  uint32_t swap_desc;                     // debug: exchange LBO / SBO
  uint32_t dbg;                           // debug: 1 = no epilogue work, 2 = no MMAs
  uint32_t a_sbo;                         // bytes between 8-row groups of A
  const __nv_bfloat16* w;
  const float* bias;
  const float* mask;
  float* stats;
  __nv_bfloat16* dst;
};

__device__ __forceinline__ uint64_t desc_interleaved(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// FAST: K = out_pitch in {32, 64, 128}, LeakyReLU with 0 <= slope <= 1 (or no activation = slope 1), 16-byte aligned destination
// (D.dx.1 forward; the data gradient of the generator's one-channel tail)
// PLAIN (with FAST): no bias, activation, mask or statistics — a data gradient: the tile is only converted and transposed
// SX: conv stride (1 or 2).  Stride 2 runs the x direction as a unit-stride conv over ALL input pixels (the window of pixel n starts 16 bytes
// after that of pixel n-1 whatever the stride is) and the epilogue reads every second accumulator column; the source rows step by 2.
template <bool FAST, bool PLAIN, int SX>
__global__ void __launch_bounds__(CM_THREADS, 1) conv_cm_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                const __grid_constant__ CmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_total = (uint32_t)p.D * p.a_bytes;
  const uint32_t ring_off = (a_total + 1023u) & ~1023u;
  const uint32_t ring_bytes = CM_STAGES * p.stage_bytes + 1024u;                      // + read-past slack behind the last slot
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_off + ring_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CM_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CM_STAGES + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * CM_STAGES + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * CM_STAGES + 2 + b); };
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x);
    for (int s = 0; s < CM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), CM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  // A_d[(j,k)][kk] in the canonical un-swizzled K-major layout: 8-row groups 2*KS*128 B apart, 16-byte K chunks 128 B apart.
  {
    const int chunks = 2 * p.KS;
    const int total = p.D * 128 * chunks;
    for (int idx = threadIdx.x; idx < total; idx += CM_THREADS) {
      const int c = idx % chunks, m = (idx / chunks) & 127, d = idx / (chunks * 128);
      const int j = m / p.K, k = m - j * p.K, r = d - p.stride * j;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r >= 0 && r < p.R && c * 8 < p.w_pitch)
        v = *reinterpret_cast<const uint4*>(p.w + ((int64_t)k * p.R + r) * p.w_pitch + c * 8);
      *reinterpret_cast<uint4*>(smem + (uint32_t)d * p.a_bytes + (uint32_t)(m >> 3) * p.a_sbo + (uint32_t)c * 128u +
                                (uint32_t)(m & 7) * 16u) = v;
    }
    // the windows of the last pixels of a slot run into whatever follows it: keep the ring finite from the first MMA on
    for (uint32_t i = threadIdx.x; i < ring_bytes / 16; i += CM_THREADS)
      reinterpret_cast<uint4*>(smem + ring_off)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const int items = p.img_groups * p.row_blocks;

  if (warp == 0) {
    const uint32_t leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int ig = it / p.row_blocks, y0 = (it - ig * p.row_blocks) * p.JB * p.stride;     // first source row
      mbar_wait(empty_bar(s), ph ^ 1);
      mbar_expect_tx_if(full_bar(s), (uint32_t)p.D * p.d_bytes, leader);
      tma_load_4d_if(smem_base + ring_off + (uint32_t)s * p.stage_bytes, &map_x, full_bar(s), 0, 0, ig * p.NI, y0, leader);
      if (++s == CM_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 || warp == CM_EPI_WARPS + 2) {
    // two issuing warps, one accumulator each: even items from warp 1, odd items from the last warp
    const uint32_t wi = warp == 1 ? 0u : 1u;
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(128, p.ncols, 0, 0);
    const uint32_t chunks = 2u * (uint32_t)p.KS;
    const uint64_t dA = p.swap_desc ? desc_interleaved(0, p.a_sbo, 128u) : desc_interleaved(0, 128u, p.a_sbo);
    const uint64_t dB = p.swap_desc ? desc_interleaved(0, 128u, 16u) : desc_interleaved(0, 16u, 128u);
    const uint32_t a_hi = (uint32_t)(dA >> 32), a_lo0 = (uint32_t)dA;
    const uint32_t b_hi = (uint32_t)(dB >> 32), b_lo0 = (uint32_t)dB;
    const uint32_t a16 = (smem_base >> 4) & 0x3FFFu;
    int s = 0;
    uint32_t ph = 0, aph = 0, buf = 0;                        // aph: phase of this warp's accumulator
    for (int it = blockIdx.x; it < items; it += gridDim.x, buf ^= 1u) {
      if (buf != wi) {
        if (++s == CM_STAGES) { s = 0; ph ^= 1; }
        continue;
      }
      mbar_wait(full_bar(s), ph);
      mbar_wait(acc_empty(buf), aph ^ 1u);
      tc_fence_after();
      const uint32_t b16 = ((smem_base + ring_off + (uint32_t)s * p.stage_bytes) >> 4) & 0x3FFFu;
      const uint32_t acc = tmem_base + (uint32_t)buf * CM_ACC_COLS;
      for (int d = 0; d < p.D; ++d) {
        for (int ks = 0; ks < p.KS; ++ks) {
          const uint32_t a_lo = a_lo0 | (a16 + (uint32_t)d * (p.a_bytes >> 4) + (uint32_t)ks * 16u);      // two 128-byte K chunks
          const uint32_t b_lo = b_lo0 | (b16 + (uint32_t)d * (p.d_bytes >> 4) + (uint32_t)ks * 2u);       // 32 bytes of the window
          if (p.dbg != 2) umma_bf16_lo2(acc, a_lo, a_hi, b_lo, b_hi, idesc, (d | ks) ? 1u : 0u, leader);
        }
      }
      umma_commit_if(empty_bar(s), leader);
      umma_commit_if(acc_full(buf), leader);
      aph ^= 1u;
      if (++s == CM_STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    // CM_EPI_WARPS / 4 warps per TMEM lane quarter, each owning a share of the item's images.  History of this role (MMAs
    // switched off, us per launch): one image after the other with 2-byte global stores 227 (a chain of latencies); masks
    // loaded ahead + 16 warps 141; tile transposed through shared memory, 16-byte stores 128 — by then issue bound (ncu: 1300
    // instructions per warp and item, half of them bounds checks, per-element activation branches and 16-bit extraction);
    // groups of 8 pixels with one uniform validity branch and no work for the unused columns 75.  A version with the unit
    // body inlined four times (two images in flight, both activation paths) went back to 171: instruction-cache misses
    // (`no_inst` stalls) — the body below exists ONCE per template instance, inside a rolled loop over the units.
    const int e = warp - 2;
    const int q4 = warp & 3;                            // TMEM lane quarter this warp may read
    const int sub = e >> 2, nsub = CM_EPI_WARPS / 4;
    const int m = q4 * 32 + lane;
    const int j = m / p.K, k = m - j * p.K;
    const int kb = k - lane;                            // FAST (K a multiple of 32): first channel of this warp's 32
    const float bias = p.bias ? p.bias[k] : 0.f;
    constexpr int UO = 32 / SX;                         // output pixels per 32-column unit
    const int cpi = (p.Q + UO - 1) / UO;                // units per image
    const int units = p.NI * cpi;                       // units of an item: dealt round-robin to the nsub warps of a lane quarter
    int il_of[CM_MAX_IMG], c0_of[CM_MAX_IMG];           // (a wide image is ONE image per item: its row is split over the warps)
#pragma unroll
    for (int t = 0; t < CM_MAX_IMG; ++t) {
      const int u = sub + t * nsub;
      il_of[t] = u < units ? u / cpi : -1;
      c0_of[t] = u < units ? (u - (u / cpi) * cpi) * UO : 0;
    }
    const int act = p.act;
    const float slope = p.slope;
    float ssum = 0.f, ssq = 0.f;
    const uint32_t tile = smem_u32(tmem_slot + 4) + 2u * (CM_EPI_WARPS / 4) * 128u * 4u + (uint32_t)e * 2048u + (uint32_t)lane * 2u;   // [pixel][32 ch]
    const uint32_t tile_rd = tile - (uint32_t)lane * 2u + (uint32_t)lane * 16u;
    int buf = 0;
    uint32_t aph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int ig = it / p.row_blocks, y = (it - ig * p.row_blocks) * p.JB + j;
      const int n_lo = ig * p.NI;
      float mk[CM_MAX_IMG];
#pragma unroll
      for (int t = 0; t < CM_MAX_IMG; ++t)
        mk[t] = (p.mask && il_of[t] >= 0 && n_lo + il_of[t] < p.N) ? p.mask[(int64_t)(n_lo + il_of[t]) * p.mask_pitch + k] : 1.f;
      mbar_wait(acc_full(buf), (aph >> buf) & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)buf * CM_ACC_COLS;
#pragma unroll 1
      for (int ut = 0; ut < CM_MAX_IMG; ++ut) {
        int il = il_of[0], c0 = c0_of[0];
        float mkv = mk[0];
#pragma unroll
        for (int t = 1; t < CM_MAX_IMG; ++t) {
          il = ut == t ? il_of[t] : il;
          c0 = ut == t ? c0_of[t] : c0;
          mkv = ut == t ? mk[t] : mkv;
        }
        if (il < 0 || n_lo + il >= p.N || p.dbg == 1) break;                     // warp-uniform (units are dealt in image order)
        uint32_t v[32];
        const int nv = p.Q - c0 < UO ? p.Q - c0 : UO;                            // valid output pixels of this unit
        tmem_ld16(tacc + (uint32_t)(il * p.W + SX * c0), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        if (SX * nv > 16) tmem_ld16(tacc + (uint32_t)(il * p.W + SX * c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        const int n = n_lo + il;
        tmem_ld_wait();
        if (FAST) __syncwarp();                                                  // the previous unit's tile has been read
#pragma unroll
        for (int g8 = 0; g8 < UO; g8 += 8) {
          if (g8 >= nv) break;                                                   // warp-uniform: unused accumulator columns
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            float f0 = __uint_as_float(v[SX * (g8 + t)]), f1 = __uint_as_float(v[SX * (g8 + t + 1)]);
            if (!PLAIN) {
              f0 += bias;
              f1 += bias;
              if (FAST) {                                                        // LeakyReLU, 0 <= slope <= 1: max(x, slope*x)
                f0 = fmaxf(f0, f0 * slope);
                f1 = fmaxf(f1, f1 * slope);
              } else {
                f0 = icf::apply_act(f0, act, slope);
                f1 = icf::apply_act(f1, act, slope);
              }
              f0 *= mkv;
              f1 *= mkv;
            }
            pk[t >> 1] = pack_bf16(f0, f1);
          }
          if (y < p.P) {
            const bool whole = g8 + 8 <= nv;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              if (!PLAIN) {
                const float r = __uint_as_float((t & 1) ? (pk[t >> 1] & 0xFFFF0000u) : (pk[t >> 1] << 16));   // the value AS STORED
                if (whole || g8 + t < nv) {
                  ssum += r;
                  ssq = fmaf(r, r, ssq);
                }
              }
              const uint16_t h = (uint16_t)((t & 1) ? (pk[t >> 1] >> 16) : (pk[t >> 1] & 0xFFFFu));
              if (FAST) {
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(tile + (uint32_t)((g8 + t) * 64)), "h"(h) : "memory");
              } else if (g8 + t < nv) {
                *reinterpret_cast<uint16_t*>(p.dst + (((int64_t)n * p.P + y) * p.Q + c0 + g8 + t) * p.out_pitch + k) = h;
              }
            }
          }
        }
        if (FAST) {
          // K = out_pitch, a multiple of 32: the warp's lanes are 32 consecutive channels of ONE output row; the [pixel][32 ch]
          // tile leaves as 16-byte stores (K = 32: 512 contiguous bytes per instruction)
          __syncwarp();
          if (y < p.P) {
            uint8_t* og = reinterpret_cast<uint8_t*>(p.dst + (((int64_t)n * p.P + y) * p.Q + c0) * p.K + kb);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int px = (lane >> 2) + 8 * r;
              if (px < nv) {
                uint4 w;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                             : "r"(tile_rd + (uint32_t)(r * 512)) : "memory");
                *reinterpret_cast<uint4*>(og + (size_t)px * (size_t)(2 * p.K) + (size_t)(lane & 3) * 16) = w;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      aph ^= 1u << buf;
      buf ^= 1;
    }
    // BatchNorm statistics: per-thread partial sums -> one pair of atomics per channel and CTA
    float* red = reinterpret_cast<float*>(tmem_slot + 4);                  // [2][nsub][128]
    red[sub * 128 + m] = ssum;
    red[(nsub + sub) * 128 + m] = ssq;
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats && (int)threadIdx.x < 2 * p.K) {
    const int which = (int)threadIdx.x / p.K, k = (int)threadIdx.x - which * p.K;
    const float* red = reinterpret_cast<const float*>(tmem_slot + 4) + which * (CM_EPI_WARPS / 4) * 128;
    float t = 0.f;
    for (int sub = 0; sub < CM_EPI_WARPS / 4; ++sub)
      for (int mm = k; mm < 128; mm += p.K) t += red[sub * 128 + mm];
    atomicAdd(p.stats + which * p.K + k, t);
  }
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// returns 0 = launched, -1 = not this kernel's case, >0 = error
int icf_cm_conv_forward(const icf_conv_args* a, cudaStream_t st) {
  if (a->dtype != ICF_BF16 || a->form != ICF_FORM_GATHER || a->win < 2 || a->S != 1 || (a->stride != 1 && a->stride != 2) || a->pad != 0) return -1;
  if (a->in_pitch != 8 || a->out_f32 || a->accumulate) return -1;
  if (a->K < 16 || a->K > 128 || (128 % a->K) || a->out_pitch < a->K) return -1;
  if (a->w_pitch & 7) return -1;
  if ((reinterpret_cast<uintptr_t>(a->src) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15)) return -1;
  if ((a->P - 1) * a->stride + a->R > a->H || (a->Q - 1) * a->stride + a->win > a->W) return -1;
  static const int mode = []() { const char* e = getenv("ICF_CM"); return e && e[0] ? atoi(e) : 1; }();   // 0 off, 2 = swapped strides (debug)
  if (mode == 0) return -1;
  CmParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.H = a->H; p.W = a->W; p.P = a->P; p.Q = a->Q; p.K = a->K; p.R = a->R;
  p.JB = 128 / a->K;
  p.stride = a->stride;
  p.D = a->stride * (p.JB - 1) + a->R;
  p.KS = (a->win * 8 + 15) / 16;
  if (p.D > 12 || p.KS > 4) return -1;
  // images per item: MMA N = NI*W <= 256, multiple of 16
  p.NI = 0;
  const int cpi = (a->stride * a->Q + 31) / 32;                                // 32-column epilogue units per image
  for (int ni = 256 / a->W; ni >= 1; --ni)                                     // (the epilogue reads whole 32-column units)
    if ((ni * a->W) % 16 == 0 && (ni - 1) * a->W + cpi * 32 <= 256 && ni * cpi <= CM_MAX_IMG * (CM_EPI_WARPS / 4)) { p.NI = ni; break; }
  p.ncols = p.NI * a->W;
  if (p.NI == 0 && cpi * 32 <= 256 && cpi <= CM_MAX_IMG * (CM_EPI_WARPS / 4)) {
    // a wide image (the 130- and 258-pixel rows of the spectrogram families): ONE image per item, the accumulator columns are the
    // pixels of its row up to the last window start; the windows of the last columns run into the next source row (finite data)
    p.NI = 1;
    p.ncols = ((a->stride * (a->Q - 1) + 1) + 15) & ~15;
  }
  if (p.NI == 0 || p.ncols < 64 || p.ncols > 256) return -1;
  p.row_blocks = icf::cdiv(a->P, p.JB);
  p.img_groups = icf::cdiv(a->N, p.NI);
  p.w_pitch = a->w_pitch; p.out_pitch = a->out_pitch; p.mask_pitch = a->mask_pitch;
  p.act = a->act; p.slope = a->slope;
  static const int apad = []() { const char* e = getenv("ICF_CM_APAD"); return e && e[0] ? atoi(e) : 0; }();
  static const int dbg = []() { const char* e = getenv("ICF_CM_DBG"); return e && e[0] ? atoi(e) : 0; }();
  p.dbg = (uint32_t)dbg;
  p.a_sbo = 2u * (uint32_t)p.KS * 128u + (uint32_t)apad;
  p.a_bytes = 16u * p.a_sbo;
  p.d_bytes = (uint32_t)p.NI * (uint32_t)a->W * 16u;
  p.stage_bytes = ((uint32_t)p.D * p.d_bytes + 1023u) & ~1023u;
  p.swap_desc = mode == 2;
  p.w = reinterpret_cast<const __nv_bfloat16*>(a->w);
  p.bias = a->bias; p.mask = a->out_mask; p.stats = a->stats;
  p.dst = reinterpret_cast<__nv_bfloat16*>(a->dst);
  const size_t smem = (((size_t)p.D * p.a_bytes + 1023) & ~size_t(1023)) + (size_t)CM_STAGES * p.stage_bytes + 1024 + 256 + 1024 + 2 * (CM_EPI_WARPS / 4) * 128 * 4 + CM_EPI_WARPS * 2048;
  if (smem > 227 * 1024) return -1;

  CUtensorMap mx;
  {
    // src [N][H][W][8] viewed as (xw pixels x 8 ch, W/xw, image, row): the box lands as [source row d][image i][x][8 ch].
    // The innermost box dimension is a whole run of pixels (<= 256 elements), NOT the 16 bytes of one pixel: TMA moves a box
    // as one request per innermost row, and 1792 requests of 16 bytes per work item made this kernel 2x slower than the
    // one it replaces (0.244 ms) while the tensor pipe waited for data.
    EncodeFn fn = reinterpret_cast<EncodeFn>(get_encode());
    ICF_REQUIRE(fn, "first-layer conv: cuTensorMapEncodeTiled is unavailable");
    int xw = a->W;
    while (xw * 8 > 256 || a->W % xw) --xw;
    cuuint64_t dims[4] = {(cuuint64_t)xw * 8, (cuuint64_t)(a->W / xw), (cuuint64_t)a->N, (cuuint64_t)a->H};
    cuuint64_t str[3] = {(cuuint64_t)xw * 16, (cuuint64_t)a->H * a->W * 16, (cuuint64_t)a->W * 16};
    cuuint32_t box[4] = {(cuuint32_t)xw * 8, (cuuint32_t)(a->W / xw), (cuuint32_t)p.NI, (cuuint32_t)p.D};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = fn(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->src), dims, str, box, est,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ICF_REQUIRE(r == CUDA_SUCCESS, "first-layer conv: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  const bool fast = (a->K & 31) == 0 && a->out_pitch == a->K && (reinterpret_cast<uintptr_t>(a->dst) & 15) == 0 &&
                    ((a->act == ICF_ACT_LRELU && a->slope >= 0.f && a->slope <= 1.f) || a->act == ICF_ACT_NONE);
  if (fast && a->act == ICF_ACT_NONE) p.slope = 1.f;            // max(x, 1*x) = x
  if (a->stride == 2 && !fast) return -1;
  const bool plain = fast && a->stride == 1 && a->act == ICF_ACT_NONE && !a->bias && !a->out_mask && !a->stats;
  static icf::SmemGuard guard_f, guard_g, guard_p, guard_2;
  if (int r = a->stride == 2 ? guard_2.ensure(reinterpret_cast<const void*>(conv_cm_kernel<true, false, 2>), smem, "first-layer conv")
              : plain ? guard_p.ensure(reinterpret_cast<const void*>(conv_cm_kernel<true, true, 1>), smem, "first-layer conv")
              : fast ? guard_f.ensure(reinterpret_cast<const void*>(conv_cm_kernel<true, false, 1>), smem, "first-layer conv")
                     : guard_g.ensure(reinterpret_cast<const void*>(conv_cm_kernel<false, false, 1>), smem, "first-layer conv"))
    return r;
  const int sms = icf::sm_count();
  const int64_t items = (int64_t)p.img_groups * p.row_blocks;
  const int grid = items < sms ? (int)items : sms;
  if (a->stride == 2) conv_cm_kernel<true, false, 2><<<grid, CM_THREADS, smem, st>>>(mx, p);
  else if (plain) conv_cm_kernel<true, true, 1><<<grid, CM_THREADS, smem, st>>>(mx, p);
  else if (fast) conv_cm_kernel<true, false, 1><<<grid, CM_THREADS, smem, st>>>(mx, p);
  else conv_cm_kernel<false, false, 1><<<grid, CM_THREADS, smem, st>>>(mx, p);
  return icf::check_launch("conv_cm");
}
