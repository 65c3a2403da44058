// Scatter-form transposed convolution for FEW output channels (sm_100a: tcgen05 / TMEM / TMA).
//
//   out[n, y+r, x+s, k] += sum_c in[n, y, x, c] * w[k][r*S+s][c]        stride 1, no padding, K <= 8, W+S-1 <= 32
//
// (the Cout = 1 generator tail `G.layers.8`, and the data gradient of the 5-channel first discriminator layer).  A
// gather-form implicit GEMM needs R*S*ceil(C/16) MMAs of N = 16 per 128 output pixels and is bound by the shared-
// memory reads of its A operand.  Here the taps move into the N dimension instead:
//     T[pixel][(k, tap)] = sum_c in[pixel][c] * w[k][tap][c]            ONE MMA chain of N = 8*taps per source row
// and the col2im (out[y+r][x+s] += T[y][x][tap]) happens in the epilogue REGISTERS: a tile is 4 images x 32 x-lanes
// ordered [image][x], so an epilogue warp owns one image row, the shift by s is a __shfl_up, the R partially summed
// output rows slide through registers, and every finished row leaves as one coalesced store per warp.
// ceil(C/16) MMAs per source row of 4 images instead of R*S*ceil(C/16) per 128 output pixels.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM), warps 2..5 epilogue (warp = image).
#include "icf_epilogue.cuh"

#include <cstdlib>
#include <cstring>

namespace {

using namespace icf_tc;

constexpr int SC_THREADS = 192, SC_SLOTS = 8, SC_ACC_COLS = 256;

struct ScParams {
  int N, H, W, C, K, P, Q, out_pitch;
  int taps, TP;                 // real taps, taps padded so that 8*TP is a multiple of 16
  int kdepth;                   // 16-wide K steps
  int tiles;                    // ceil(N / 4)   (shifted-operand variant: ceil(N / 3))
  uint32_t w_bytes;             // weight tile: 8*TP rows of 128 B   (shifted-operand variant: S blocks of wblk bytes)
  int NR;                       // shifted-operand variant: MMA N = K*R rounded up to 16
  int xtiles;                   // shifted-operand variant: ceil(Q / 16) column tiles
  unsigned long long* dbg;      // cycle counters of CTA 0 (build with ICF_WS_INSTRUMENT, run with ICF_SC_DEBUG), else NULL
  uint32_t wblk, wblk_tx;       // bytes between / deposited into the per-s weight blocks
  int act;
  float slope;
  int out_f32, mask_pitch;
  const float* bias;
  const float* mask;
  void* dst;
};

template <int R, int S, int KT>
__global__ void __launch_bounds__(SC_THREADS, 1) conv_sc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ ScParams p) {
  constexpr uint32_t SLOT_BYTES = 128 * 128;                      // [4 images][32 x][64 ch]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wtile = smem;
  uint8_t* slots = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + SC_SLOTS * SLOT_BYTES);   // full[8] empty[8] acc_full[2] acc_empty[2] w
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_SLOTS + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SC_SLOTS + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * SC_SLOTS + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * SC_SLOTS + 2 + b); };
  const uint32_t w_bar = bar_base + 8u * (2 * SC_SLOTS + 4);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    for (int s = 0; s < SC_SLOTS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * SC_ACC_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();              // warp-uniform loop, the elected lane issues
      mbar_expect_tx_if(w_bar, p.w_bytes, leader);
      tma_load_3d_if(smem_u32(wtile), &map_w, w_bar, 0, 0, 0, leader);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
        for (int y = 0; y < p.H; ++y) {
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx_if(full_bar(s), SLOT_BYTES, leader);
          tma_load_4d_if(smem_u32(slots) + (uint32_t)s * SLOT_BYTES, &map_a, full_bar(s), 0, 0, y, tile * 4, leader);
          if (++s == SC_SLOTS) { s = 0; ph ^= 1; }
        }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_one();              // warp-uniform issue loop, one elected lane issues
      const uint32_t idesc = make_idesc(128, 8 * p.TP, 0, 0);
      const uint64_t d0 = make_desc(0, 16, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), desc_lo = (uint32_t)d0;
      mbar_wait(w_bar, 0);
      tc_fence_after();
      const uint32_t b_lo = ((smem_u32(wtile) >> 4) & 0x3FFFu) | desc_lo;
      int s = 0;
      uint32_t ph = 0;
      uint32_t g = 0;                                   // source rows issued so far
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
        for (int y = 0; y < p.H; ++y, ++g) {
          const uint32_t buf = g & 1u;
          mbar_wait(acc_empty(buf), ((g >> 1) & 1u) ^ 1u);
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_lo = (((smem_u32(slots) + (uint32_t)s * SLOT_BYTES) >> 4) & 0x3FFFu) | desc_lo;
          for (int k = 0; k < p.kdepth; ++k)
            umma_bf16_lo(tmem_base + buf * SC_ACC_COLS, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, k ? 1u : 0u, leader);
          umma_commit_if(empty_bar(s), leader);
          umma_commit_if(acc_full(buf), leader);
          if (++s == SC_SLOTS) { s = 0; ph ^= 1; }
        }
    }
  } else {
    // ===== epilogue: warp = image of the tile, lane = x =====
    const int q4 = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int esize = p.out_f32 ? 4 : 2;
    float bias[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) bias[k] = (p.bias && k < p.K) ? __ldg(p.bias + k) : 0.f;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int n = tile * 4 + q4;
      const bool store = n < p.N && lane < p.Q;
      float mk[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) mk[k] = (p.mask && n < p.N && k < p.K) ? __ldg(p.mask + (int64_t)n * p.mask_pitch + k) : 1.f;
      float win[R][KT];                                 // partially summed output rows y .. y+R-1 of this thread's column
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < KT; ++k) win[r][k] = 0.f;
      auto emit = [&](int y_out) {                      // finished output row: bias + activation + mask, one store per pixel
        if (!store) return;
        float f[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          float x = win[0][k] + bias[k];
          if (p.act == ICF_ACT_LRELU) x = x > 0.f ? x : x * p.slope;
          else if (p.act == ICF_ACT_TANH) { if (k < p.K) x = tanhf(x); }
          f[k] = k < p.K ? x * mk[k] : 0.f;
        }
        uint8_t* o = reinterpret_cast<uint8_t*>(p.dst) + (((int64_t)n * p.P + y_out) * p.Q + lane) * p.out_pitch * esize;
        if (p.out_f32) {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k < p.K) reinterpret_cast<float*>(o)[k] = f[k];
        } else if (KT == 8 && p.out_pitch >= 8 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          uint4 w4;                                     // pitch padding is written as zeros
          w4.x = pack_bf16(f[0], f[KT > 1 ? 1 : 0]); w4.y = pack_bf16(f[KT > 2 ? 2 : 0], f[KT > 3 ? 3 : 0]);
          w4.z = pack_bf16(f[KT > 4 ? 4 : 0], f[KT > 5 ? 5 : 0]); w4.w = pack_bf16(f[KT > 6 ? 6 : 0], f[KT > 7 ? 7 : 0]);
          *reinterpret_cast<uint4*>(o) = w4;
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k < p.K) reinterpret_cast<__nv_bfloat16*>(o)[k] = __float2bfloat16_rn(f[k]);
        }
      };
      auto slide = [&]() {
#pragma unroll
        for (int r = 0; r + 1 < R; ++r)
#pragma unroll
          for (int k = 0; k < KT; ++k) win[r][k] = win[r + 1][k];
#pragma unroll
        for (int k = 0; k < KT; ++k) win[R - 1][k] = 0.f;
      };
      for (int y = 0; y < p.H; ++y, ++g) {
        const uint32_t buf = g & 1u;
        mbar_wait(acc_full(buf), (g >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (k < p.K) {
            constexpr int NCH = (R * S + 15) / 16;      // 16-column chunks of this channel's taps: all loaded, one wait
            uint32_t v[NCH][16];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) tmem_ld16(lane_addr + buf * SC_ACC_COLS + (uint32_t)(k * p.TP + ch * 16), v[ch]);
            tmem_ld_wait();
#pragma unroll
            for (int tap = 0; tap < R * S; ++tap) {
              const int r = tap / S, s = tap % S;
              const float t = __shfl_up_sync(0xffffffffu, __uint_as_float(v[tap / 16][tap % 16]), s);
              if (lane >= s) win[r][k] += t;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
        emit(y);                                        // row y has received its last contribution (tap row 0)
        slide();
      }
#pragma unroll
      for (int r = 0; r + 1 < R; ++r) {                 // the R-1 rows below the last source row
        emit(p.H + r);
        slide();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * SC_ACC_COLS>(tmem_base);
}

template <int R, int S, int KT>
int launch_sc(const CUtensorMap& ma, const CUtensorMap& mw, const ScParams& p, int grid, size_t smem, cudaStream_t st) {
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(conv_sc_kernel<R, S, KT>), smem, "scatter-form conv")) return r;
  conv_sc_kernel<R, S, KT><<<grid, SC_THREADS, smem, st>>>(ma, mw, p);
  return icf::check_launch("conv_sc");
}


// ------------------------------------------------------------------------------------------------------------------
// Shifted-operand variant for 2..8 output channels (the data gradient of `Dx.dx.1`, 32 -> 5 channels, 5x5).
// The scatter kernel above leaves the whole col2im to the epilogue: R*S*K shuffles + adds per source row and lane
// (125 for 5x5x5) make it epilogue-bound.  Here the x shift moves into the A operand instead:
//     T_y[x][(k, r)] = sum_s sum_c in[y][x - s][c] * w[k][r*S+s][c]
// is S MMA chains into ONE accumulator, chain s reading the SAME shared-memory tile from a start address shifted by s
// x-slots.  The tile is batch-innermost like the row-streaming kernel's: rows [x slot][8 images], so one x-slot is one
// 1024-byte swizzle atom and the shifted start stays atom-aligned (a start shifted by single 128-byte rows is legal but
// measured ~2.5x slower per MMA).  A tile covers 16 output columns of 8 images (128 MMA rows) and holds 20 x-slots
// (x0-4 .. x0+15; TMA zero-fills the columns outside the image).  The epilogue only slides the R output rows: K*R adds
// per source row, no shuffles; every lane stores its own pixel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SX_XT = 16, SX_PAD = 4, SX_IMGS = 8, SX_SLOTS = 6;   // output columns per tile, halo slots, images, ring (even: slot parity = row parity)
constexpr int SX_THREADS = SC_THREADS + 32;                        // + a second MMA-issuing warp (warp 6)

template <int R, int S, int KT>
__global__ void __launch_bounds__(SX_THREADS, 1) conv_sx_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_w,
                                                                const __grid_constant__ ScParams p) {
  constexpr uint32_t SLOT_BYTES = (SX_XT + SX_PAD) * SX_IMGS * 128, SLOT_TX = SLOT_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wtile = smem;
  uint8_t* slots = smem + ((p.w_bytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + SX_SLOTS * SLOT_BYTES);   // full[8] empty[8] acc_full[2] acc_empty[2] w
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_SLOTS + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SC_SLOTS + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (2 * SC_SLOTS + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (2 * SC_SLOTS + 2 + b); };
  const uint32_t w_bar = bar_base + 8u * (2 * SC_SLOTS + 4);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    for (int s = 0; s < SC_SLOTS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * SC_ACC_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef ICF_WS_INSTRUMENT
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
#else
  constexpr bool dbg_on = false;
#endif
  long long w0 = 0, w1 = 0;
  const long long t_begin = dbg_on ? clock64() : 0;
#define SX_WAIT(cnt, bar, par) do { if (dbg_on) { const long long t_ = clock64(); mbar_wait(bar, par); cnt += clock64() - t_; } else mbar_wait(bar, par); } while (0)

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();              // warp-uniform loop, the elected lane issues
      mbar_expect_tx_if(w_bar, p.wblk_tx * S, leader);
      for (int s = 0; s < S; ++s) tma_load_3d_if(smem_u32(wtile) + (uint32_t)s * p.wblk, &map_w, w_bar, 0, s, 0, leader);
      int sl = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int xt = tile % p.xtiles, nt = tile / p.xtiles;
        for (int y = 0; y < p.H; ++y) {
          SX_WAIT(w0, empty_bar(sl), ph ^ 1);
          mbar_expect_tx_if(full_bar(sl), SLOT_TX, leader);
          tma_load_4d_if(smem_u32(slots) + (uint32_t)sl * SLOT_BYTES, &map_a, full_bar(sl), 0, nt * SX_IMGS, xt * SX_XT - SX_PAD, y, leader);
          if (++sl == SX_SLOTS) { sl = 0; ph ^= 1; }
        }
      }
      if (dbg_on && lane == 0) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = w0; }
    }
  } else if (warp == 1 || warp == 6) {
    {
      // the whole warp runs the issue loop and one elected lane issues: inside an `if (lane == 0)` region every
      // tcgen05.mma costs ~190 cycles of warp time (measured), in warp-uniform code ~90.  TWO issuing warps: the kernel is
      // bound by that issue cost (ncu: the epilogue warps wait on acc_full for half of their samples, tensor pipe 12 %; ten
      // MMAs of 128 x 48 x 16 per source row) — warp 1 issues the even source rows into accumulator 0, warp 6 the odd ones into
      // accumulator 1 (slot parity = row parity because the ring has an even number of slots).
      const uint32_t wi = warp == 1 ? 0u : 1u;
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc(128, p.NR, 0, 0);
      const uint64_t d0 = make_desc(0, 16, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), desc_lo = (uint32_t)d0;      // low word without an address
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int sl = 0;
      uint32_t ph = 0;
      uint32_t g = 0;                                   // source rows issued so far (by both warps)
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
        for (int y = 0; y < p.H; ++y, ++g) {
          const uint32_t buf = g & 1u;
          if (buf == wi) {
            SX_WAIT(w0, acc_empty(buf), ((g >> 1) & 1u) ^ 1u);
            SX_WAIT(w1, full_bar(sl), ph);
            tc_fence_after();
            const uint32_t slot = smem_u32(slots) + (uint32_t)sl * SLOT_BYTES;
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
              const uint32_t a_lo = (((slot + (uint32_t)(SX_PAD - s) * 1024u) >> 4) & 0x3FFFu) | desc_lo;
              const uint32_t b_lo = (((smem_u32(wtile) + (uint32_t)s * p.wblk) >> 4) & 0x3FFFu) | desc_lo;
              for (int k = 0; k < p.kdepth; ++k)
                umma_bf16_lo(tmem_base + buf * SC_ACC_COLS, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (s | k) ? 1u : 0u, leader);
            }
            umma_commit_if(empty_bar(sl), leader);
            umma_commit_if(acc_full(buf), leader);
          }
          if (++sl == SX_SLOTS) { sl = 0; ph ^= 1; }
        }
      if (dbg_on && lane == 0 && wi == 0) { p.dbg[2] = clock64() - t_begin; p.dbg[3] = w0; p.dbg[4] = w1; }
    }
  } else {
    // ===== epilogue: TMEM lane m = x_local*8 + image; every thread owns one output column of one image =====
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int img = m & (SX_IMGS - 1), xl = m >> 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int esize = p.out_f32 ? 4 : 2;
    const int ncols = p.K * R;
    float bias[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) bias[k] = (p.bias && k < p.K) ? __ldg(p.bias + k) : 0.f;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int n = (tile / p.xtiles) * SX_IMGS + img, x = (tile % p.xtiles) * SX_XT + xl;
      const bool store = n < p.N && x < p.Q;
      float mk[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) mk[k] = (p.mask && store && k < p.K) ? __ldg(p.mask + (int64_t)n * p.mask_pitch + k) : 1.f;
      float win[R][KT];                                 // partially summed output rows y .. y+R-1 of this thread's column
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < KT; ++k) win[r][k] = 0.f;
      auto emit = [&](int y_out) {                      // finished output row: bias + activation + mask, one store per pixel
        if (!store) return;
        float f[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          float v = win[0][k] + bias[k];
          if (p.act == ICF_ACT_LRELU) v = v > 0.f ? v : v * p.slope;
          else if (p.act == ICF_ACT_TANH) { if (k < p.K) v = tanhf(v); }
          f[k] = k < p.K ? v * mk[k] : 0.f;
        }
        uint8_t* o = reinterpret_cast<uint8_t*>(p.dst) + (((int64_t)n * p.P + y_out) * p.Q + x) * p.out_pitch * esize;
        if (p.out_f32) {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k < p.K) reinterpret_cast<float*>(o)[k] = f[k];
        } else if (KT == 8 && p.out_pitch >= 8 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          uint4 w4;                                     // pitch padding is written as zeros
          w4.x = pack_bf16(f[0], f[1]); w4.y = pack_bf16(f[2], f[3]);
          w4.z = pack_bf16(f[4], f[5]); w4.w = pack_bf16(f[6], f[7]);
          *reinterpret_cast<uint4*>(o) = w4;
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (k < p.K) reinterpret_cast<__nv_bfloat16*>(o)[k] = __float2bfloat16_rn(f[k]);
        }
      };
      auto slide = [&]() {
#pragma unroll
        for (int r = 0; r + 1 < R; ++r)
#pragma unroll
          for (int k = 0; k < KT; ++k) win[r][k] = win[r + 1][k];
#pragma unroll
        for (int k = 0; k < KT; ++k) win[R - 1][k] = 0.f;
      };
      for (int y = 0; y < p.H; ++y, ++g) {
        const uint32_t buf = g & 1u;
        SX_WAIT(w0, acc_full(buf), (g >> 1) & 1u);
        tc_fence_after();
        constexpr int NCH = (KT * R + 15) / 16;         // 16-column chunks; column = k*R + r
        uint32_t v[NCH][16];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          if (ch * 16 < ncols) tmem_ld16(lane_addr + buf * SC_ACC_COLS + (uint32_t)(ch * 16), v[ch]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
#pragma unroll
        for (int k = 0; k < KT; ++k)
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (k < p.K) win[r][k] += __uint_as_float(v[(k * R + r) / 16][(k * R + r) % 16]);
        emit(y);                                        // row y has received its last contribution (tap row 0)
        slide();
      }
#pragma unroll
      for (int r = 0; r + 1 < R; ++r) {                 // the R-1 rows below the last source row
        emit(p.H + r);
        slide();
      }
    }
  }
  if (dbg_on && threadIdx.x == 64) { p.dbg[5] = clock64() - t_begin; p.dbg[6] = w0; }
#undef SX_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * SC_ACC_COLS>(tmem_base);
}

template <int R, int S, int KT>
int launch_sx(const CUtensorMap& ma, const CUtensorMap& mw, const ScParams& p, int grid, size_t smem, cudaStream_t st) {
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(conv_sx_kernel<R, S, KT>), smem, "scatter-form conv")) return r;
  conv_sx_kernel<R, S, KT><<<grid, SX_THREADS, smem, st>>>(ma, mw, p);
  return icf::check_launch("conv_sx");
}

}  // namespace

// returns 0 = launched, -1 = not this kernel's case, >0 = error
int icf_sc_conv_forward(const icf_conv_args* a, cudaStream_t st) {
  if (a->dtype != ICF_BF16 || a->accumulate || a->form != ICF_FORM_TRANSPOSED || a->stride != 1 || a->pad != 0) return -1;
  if (a->K > 8 || a->C > 64 || a->R != a->S || (a->R != 4 && a->R != 5) || a->win > 1 || a->stats) return -1;
  if (a->W + a->S - 1 > 32 || a->Q != a->W + a->S - 1 || a->P != a->H + a->R - 1) return -1;
  if ((a->in_pitch & 7) || (a->w_pitch & 7)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->src) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15)) return -1;
  static const bool off = []() { const char* e = getenv("ICF_DISABLE_SC"); return e && e[0] && e[0] != '0'; }();
  if (off) return -1;
  ScParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.H = a->H; p.W = a->W; p.C = a->C; p.K = a->K; p.P = a->P; p.Q = a->Q; p.out_pitch = a->out_pitch;
  p.taps = a->R * a->S;
  p.TP = (p.taps + 1) & ~1;
  if (8 * p.TP > SC_ACC_COLS) return -1;
  p.kdepth = icf::cdiv(a->C, 16);
  p.tiles = icf::cdiv(a->N, 4);
  p.w_bytes = (uint32_t)(8 * p.TP) * 128u;
  p.act = a->act; p.slope = a->slope; p.out_f32 = a->out_f32; p.mask_pitch = a->mask_pitch;
  p.bias = a->bias; p.mask = a->out_mask; p.dst = a->dst;
  const int sms = icf::sm_count();
  static const bool sx_off = []() { const char* e = getenv("ICF_DISABLE_SX"); return e && e[0] && e[0] != '0'; }();
  if (a->K > 1 && !sx_off && a->S - 1 <= SX_PAD) {
    // ---- shifted-operand variant ----
    p.NR = (a->K * a->R + 15) & ~15;
    const int krows = icf::cdiv(p.NR, a->R);                       // weight rows (k) one block holds, >= K
    p.wblk_tx = (uint32_t)(krows * a->R) * 128u;
    p.wblk = (p.wblk_tx + 1023u) & ~1023u;
    p.w_bytes = p.wblk * (uint32_t)a->S;
    p.xtiles = icf::cdiv(a->Q, SX_XT);
    p.tiles = icf::cdiv(a->N, SX_IMGS) * p.xtiles;
    CUtensorMap ma, mw;
    {
      // [N][H][W][C] read in (C, N, W, H) order as boxes {64 ch, 8 images, 20 x from x0-4, 1 row}: rows [x slot][image]
      cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->N, (cuuint64_t)a->W, (cuuint64_t)a->H};
      cuuint64_t str[3] = {(cuuint64_t)a->H * a->W * a->in_pitch * 2, (cuuint64_t)a->in_pitch * 2, (cuuint64_t)a->W * a->in_pitch * 2};
      cuuint32_t box[4] = {64, SX_IMGS, SX_XT + SX_PAD, 1};
      cuuint32_t est[4] = {1, 1, 1, 1};
      if (int r = encode_map(&ma, a->src, 4, dims, str, box, est)) return r;
    }
    {
      // packed weights [K rows][taps][w_pitch]: one box per filter column s = {64 ch, taps s, s+S, .. (R of them), krows}:
      // block row = k*R + r
      cuuint64_t dims[3] = {(cuuint64_t)a->C, (cuuint64_t)p.taps, (cuuint64_t)a->w_rows};
      cuuint64_t str[2] = {(cuuint64_t)a->w_pitch * 2, (cuuint64_t)p.taps * a->w_pitch * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)(a->R * a->S), (cuuint32_t)krows};
      cuuint32_t est[3] = {1, (cuuint32_t)a->S, 1};
      if (int r = encode_map(&mw, a->w, 3, dims, str, box, est)) return r;
    }
    const size_t smem = ((p.w_bytes + 1023u) & ~1023u) + (size_t)SX_SLOTS * (SX_XT + SX_PAD) * SX_IMGS * 128 + 1024 + 512;
    const int grid = p.tiles < sms ? p.tiles : sms;
    static unsigned long long* dbg_buf = []() -> unsigned long long* {
      const char* e = getenv("ICF_SC_DEBUG");
      void* d = nullptr;
      if (e && e[0] && e[0] != '0' && cudaMalloc(&d, 64) == cudaSuccess) cudaMemset(d, 0, 64);
      return reinterpret_cast<unsigned long long*>(d);
    }();
    p.dbg = dbg_buf;
    const int rc = a->R == 4 ? launch_sx<4, 4, 8>(ma, mw, p, grid, smem, st) : launch_sx<5, 5, 8>(ma, mw, p, grid, smem, st);
    if (rc == 0 && p.dbg) {
      unsigned long long h[8];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[icf sx] K=%d C=%d R=%d tiles=%d grid=%d | CTA0 cycles: producer %llu (wait empty %llu) | mma %llu (wait acc_empty %llu, wait full %llu) | epilogue %llu (wait acc_full %llu)\n",
              a->K, a->C, a->R, p.tiles, grid, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return rc;
  }
  CUtensorMap ma, mw;
  {
    // [N][H][W][C] read as boxes {64 ch, 32 x, 1 row, 4 images}: tile rows ordered [image][x]
    cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->in_pitch * 2, (cuuint64_t)a->W * a->in_pitch * 2, (cuuint64_t)a->H * a->W * a->in_pitch * 2};
    cuuint32_t box[4] = {64, 32, 1, 4};
    cuuint32_t est[4] = {1, 1, 1, 1};
    if (int r = encode_map(&ma, a->src, 4, dims, str, box, est)) return r;
  }
  {
    // packed weights [K rows][taps][w_pitch] read as ONE box {64 ch, TP taps, 8 rows}: tile row = k*TP + tap
    cuuint64_t dims[3] = {(cuuint64_t)a->C, (cuuint64_t)p.taps, (cuuint64_t)a->w_rows};
    cuuint64_t str[2] = {(cuuint64_t)a->w_pitch * 2, (cuuint64_t)p.taps * a->w_pitch * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)p.TP, 8};
    cuuint32_t est[3] = {1, 1, 1};
    if (int r = encode_map(&mw, a->w, 3, dims, str, box, est)) return r;
  }
  const size_t smem = ((p.w_bytes + 1023u) & ~1023u) + (size_t)SC_SLOTS * 128 * 128 + 1024 + 512;
  const int grid = p.tiles < sms ? p.tiles : sms;
  if (a->R == 4) return a->K == 1 ? launch_sc<4, 4, 1>(ma, mw, p, grid, smem, st) : launch_sc<4, 4, 8>(ma, mw, p, grid, smem, st);
  return a->K == 1 ? launch_sc<5, 5, 1>(ma, mw, p, grid, smem, st) : launch_sc<5, 5, 8>(ma, mw, p, grid, smem, st);
}
