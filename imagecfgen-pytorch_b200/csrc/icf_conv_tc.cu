// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//   D[m][k] = epi( sum_{tap, c} A[pix(m, tap)][c] * Wp[k][tap][c] )        m = output pixel, k = output channel
//
// One CTA computes a 128 x TILE_N tile of D.  The 128 rows are a BOX (tn images x ti rows x tj columns) of the
// output index space of one "class"; for a fixed filter tap the source pixels of such a box are again a box of
// the NHWC activation tensor, so a single 4-D TMA tiled load (traversal stride = conv stride, out-of-bounds =
// zero = padding) brings the 128 x 64-channel A operand of that tap into shared memory in the 128B-swizzled
// K-major layout tcgen05.mma reads.  The transposed form (ConvTranspose2d forward / Conv2d dgrad) is split into
// stride^2 output-parity classes, each a unit-stride gather over the taps of its parity — no zero insertion.
// Taps whose source box lies entirely in the padding are skipped by producer and MMA issuer alike.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/activation/Dropout2d mask -> global), one D row per thread.
#include "icf_epilogue.cuh"

namespace {

using namespace icf_tc;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int MAX_CLASSES = 16, MAX_TAPS = 25;
constexpr int NUM_THREADS = 192;
inline int tc_sm_count() { return icf::sm_count(); }

struct TapTable {
  int16_t ntaps[MAX_CLASSES];
  int8_t tap[MAX_CLASSES][MAX_TAPS];   // index r*S+s into the packed weights
  int8_t dy[MAX_CLASSES][MAX_TAPS];    // source row    = i*sstep + dy
  int8_t dx[MAX_CLASSES][MAX_TAPS];    // source column = j*sstep + dx
};

struct TcParams {
  int N, H, W, C;
  int P, Q, K, out_pitch;
  int sstep, ostep;
  int n_classes;
  int ti, tj, tn;
  int tiles_i, tiles_j, tiles_n, tiles_k;
  int kchunks, w_pitch;
  uint32_t a_tx_bytes;          // bytes one A box deposits: tn*ti*tj rows of 128 B
  int act;
  float slope;
  int out_f32, mask_pitch;
  const float* bias;
  const float* mask;
  void* dst;
  int splits;                   // > 1: the (tap, channel chunk) iterations of a tile are dealt to `splits` work items whose raw fp32
  float* partial;               //      accumulators are ADDED to partial[pixel][partial_pitch] (bias / activation / mask: splitk_finish_kernel)
  int partial_pitch;
  TapTable tt;
};

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel — PERSISTENT: a CTA walks tiles t = blockIdx.x, += gridDim.x (the grid is one resident wave).
// The shared-memory ring and its phases run on across tiles; with ACC = 2 the TMEM accumulator is double-buffered so
// that the epilogue of tile t (TMEM -> bias / activation / Dropout2d mask -> global) overlaps the TMA + MMA main loop of
// tile t+1 — most layers of these networks have short K loops (1-4 taps of a stride-2 class x 2-4 channel chunks), where
// the one-tile-per-CTA version spent more time in barrier set-up, TMEM allocation, pipeline fill and the epilogue than in
// the MMAs.  ACC = 1 (256-wide tiles: 2 x 256 columns would take the whole TMEM and forbid a second CTA per SM) keeps a
// single accumulator; there the second resident CTA provides the overlap.
// ------------------------------------------------------------------------------------------------
struct TileCoord {
  int cls, py, px, Pi, Qj, i0, j0, n0, k0;
  bool inside;
};

__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int t, int tile_n) {
  TileCoord c;
  const int kt = t % p.tiles_k; t /= p.tiles_k;
  const int jt = t % p.tiles_j; t /= p.tiles_j;
  const int it = t % p.tiles_i; t /= p.tiles_i;
  const int nt = t % p.tiles_n;
  c.cls = t / p.tiles_n;
  c.py = c.cls / p.ostep;
  c.px = c.cls - c.py * p.ostep;
  c.Pi = (p.P - c.py + p.ostep - 1) / p.ostep;
  c.Qj = (p.Q - c.px + p.ostep - 1) / p.ostep;
  c.i0 = it * p.ti; c.j0 = jt * p.tj; c.n0 = nt * p.tn; c.k0 = kt * tile_n;
  c.inside = c.i0 < c.Pi && c.j0 < c.Qj;     // false: the tile lies outside this (smaller) parity class
  return c;
}

// split-K: iterations [lo, hi) of the tile's n_iters belong to split sp
__device__ __forceinline__ void split_range(int n_iters, int sp, int splits, int& lo, int& hi) {
  lo = (int)(((int64_t)n_iters * sp) / splits);
  hi = (int)(((int64_t)n_iters * (sp + 1)) / splits);
}

// a tap is live for a tile when its source box touches the un-padded input
__device__ __forceinline__ bool tap_live(const TcParams& p, const TileCoord& c, int ti_) {
  const int ylo = c.i0 * p.sstep + p.tt.dy[c.cls][ti_], yhi = ylo + (p.ti - 1) * p.sstep;
  const int xlo = c.j0 * p.sstep + p.tt.dx[c.cls][ti_], xhi = xlo + (p.tj - 1) * p.sstep;
  return yhi >= 0 && ylo < p.H && xhi >= 0 && xlo < p.W;
}

template <int TILE_N, int STAGES, int ACC>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b,
                                                              const __grid_constant__ TcParams p) {
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr uint32_t B_BYTES = TILE_N * BLOCK_K * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_COLS = TILE_N < 32 ? 32 : TILE_N;
  constexpr int TMEM_COLS = ACC * ACC_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // full[S] empty[S] tmem_full[2] tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* sbias = reinterpret_cast<float*>(bars + 16);           // TILE_N floats, 128 B past the barriers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.n_classes * p.tiles_n * p.tiles_i * p.tiles_j * p.tiles_k;

  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int b_) { return bar_base + 8u * (2 * STAGES + b_); };
  auto tmem_empty_bar = [&](int b_) { return bar_base + 8u * (2 * STAGES + 2 + b_); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b_ = 0; b_ < 2; ++b_) {
      mbar_init(tmem_full_bar(b_), 1);
      mbar_init(tmem_empty_bar(b_), 4);      // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loops, the elected lane issues) =====
    const uint32_t leader = elect_one();
    int iter = 0;                                  // ring position, runs on across tiles
    for (int w = blockIdx.x; w < total_tiles * p.splits; w += gridDim.x) {
      const int t = w / p.splits, sp = w - t * p.splits;
      const TileCoord c = tile_coord(p, t, TILE_N);
      if (!c.inside) continue;
      const int ntaps = p.tt.ntaps[c.cls];
      int lo = 0, hi = 0x7fffffff, ii = 0;
      if (p.splits > 1) {
        int live = 0;
        for (int i = 0; i < ntaps; ++i) live += tap_live(p, c, i) ? 1 : 0;
        split_range(live * p.kchunks, sp, p.splits, lo, hi);
      }
      for (int ti_ = 0; ti_ < ntaps; ++ti_) {
        if (!tap_live(p, c, ti_)) continue;
        const int y = c.i0 * p.sstep + p.tt.dy[c.cls][ti_], x = c.j0 * p.sstep + p.tt.dx[c.cls][ti_];
        const int wcol = p.tt.tap[c.cls][ti_] * p.w_pitch;
        for (int kc = 0; kc < p.kchunks; ++kc, ++ii) {
          if (ii < lo || ii >= hi) continue;
          const int s = iter % STAGES;
          mbar_wait(empty_bar(s), ((iter / STAGES) & 1) ^ 1);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
          mbar_expect_tx_if(full_bar(s), p.a_tx_bytes + B_BYTES, leader);
          tma_load_4d_if(sa, &map_a, full_bar(s), kc * BLOCK_K, x, y, c.n0, leader);
          tma_load_2d_if(sb, &map_b, full_bar(s), wcol + kc * BLOCK_K, c.k0, leader);
          ++iter;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (inside an `if (lane == 0)` region
    // every tcgen05.mma cost ~190 cycles of warp time, in warp-uniform code ~90: measured on the scatter kernel) =====
    const uint32_t leader = elect_one();
    constexpr uint32_t idesc = make_idesc(BLOCK_M, TILE_N, 0, 0);
    const uint64_t d0 = make_desc(0, 16, 1024);
    const uint32_t desc_hi = (uint32_t)(d0 >> 32), desc_lo = (uint32_t)d0;      // low word without an address
    int iter = 0, done = 0;                        // ring position; tiles this CTA has accumulated so far
    for (int w = blockIdx.x; w < total_tiles * p.splits; w += gridDim.x) {
      const int t = w / p.splits, sp = w - t * p.splits;
      const TileCoord c = tile_coord(p, t, TILE_N);
      if (!c.inside) continue;
      const int ntaps = p.tt.ntaps[c.cls];
      int live = 0;
      for (int i = 0; i < ntaps; ++i) live += tap_live(p, c, i) ? 1 : 0;
      int n_iters = live * p.kchunks;
      if (p.splits > 1) {
        int lo, hi;
        split_range(n_iters, sp, p.splits, lo, hi);
        n_iters = hi - lo;
      }
      const int b_ = ACC == 2 ? (done & 1) : 0;
      const int use = ACC == 2 ? (done >> 1) : done;
      mbar_wait(tmem_empty_bar(b_), (use & 1) ^ 1);          // the epilogue has drained this accumulator (first use: free)
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(b_ * ACC_COLS);
      for (int i = 0; i < n_iters; ++i, ++iter) {
        const int s = iter % STAGES;
        mbar_wait(full_bar(s), (iter / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
        const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | desc_lo, b_lo = ((sb >> 4) & 0x3FFFu) | desc_lo;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_bf16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (i | k) ? 1u : 0u, leader);
        umma_commit_if(empty_bar(s), leader);
      }
      if (n_iters > 0) umma_commit_if(tmem_full_bar(b_), leader);
      else if (lane == 0) mbar_arrive(tmem_full_bar(b_));
      ++done;
    }
  } else {
    // ===== epilogue: one D row (TMEM lane) per thread =====
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int tij = p.ti * p.tj;
    const int tn_i = m / tij, rem = m - tn_i * tij;
    const int ti_i = rem / p.tj, tj_i = rem - ti_i * p.tj;
    const int esize = p.out_f32 ? 4 : 2;
    int done = 0;
    for (int w = blockIdx.x; w < total_tiles * p.splits; w += gridDim.x) {
      const int t = w / p.splits, sp = w - t * p.splits;
      const TileCoord c = tile_coord(p, t, TILE_N);
      if (!c.inside) continue;
      const int k0 = c.k0;
      const int ntaps = p.tt.ntaps[c.cls];
      int live = 0;
      for (int i = 0; i < ntaps; ++i) live += tap_live(p, c, i) ? 1 : 0;
      int n_iters = live * p.kchunks;
      if (p.splits > 1) {
        int lo, hi;
        split_range(n_iters, sp, p.splits, lo, hi);
        n_iters = hi - lo;
      }
      // bias tile -> shared memory (zero where absent / beyond K), visible to the four epilogue warps; the first barrier
      // keeps a fast warp from overwriting the previous tile's bias while a slow one still reads it
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int j = (int)threadIdx.x - 64; j < TILE_N; j += 128) sbias[j] = (p.bias && k0 + j < p.K) ? __ldg(p.bias + k0 + j) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int n = c.n0 + tn_i, ii = c.i0 + ti_i, jj = c.j0 + tj_i;
      const bool valid = (tn_i < p.tn) && n < p.N && ii < c.Pi && jj < c.Qj;
      const int op = c.py + ii * p.ostep, oq = c.px + jj * p.ostep;
      const int64_t pix = valid ? ((int64_t)n * p.P + op) * p.Q + oq : 0;
      const float* mrow = (p.mask && valid) ? p.mask + (int64_t)n * p.mask_pitch + k0 : nullptr;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.dst) + (pix * p.out_pitch + k0) * esize;
      // Dropout2d mask of this thread's row: while the main loop runs (these warps would idle on tmem_full) its lines are
      // pulled into L2, and the 16 values of group g+1 are loaded (four 16-byte loads) while group g is in the math — a
      // 1x1-spatial layer reads one mask row PER OUTPUT ROW; fetched on demand, 16 scalar loads per group from DRAM held the
      // 64-CTA GEMMs at ~35 us (ncu: the epilogue's FMUL waits on them, profiles/r01_final_stall_summary.txt).
      const int kcols = (p.K - k0) < TILE_N ? (p.K - k0) : TILE_N;
      const bool mvec = mrow && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0) && (kcols & 15) == 0;
      if (mrow) {
        for (int j = 0; j < kcols; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow + j));
      }
      float mk[16], mk_next[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) mk_next[j] = 1.f;
      auto load_mask = [&](int c0) {
        if (!mrow) return;
        if (mvec) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(mrow + c0 + j));
            mk_next[j] = t4.x; mk_next[j + 1] = t4.y; mk_next[j + 2] = t4.z; mk_next[j + 3] = t4.w;
          }
        } else {
          const int nv = p.K - (k0 + c0);
#pragma unroll
          for (int j = 0; j < 16; ++j) mk_next[j] = (j < nv) ? __ldg(mrow + c0 + j) : 0.f;
        }
      };
      load_mask(0);
      const int b_ = ACC == 2 ? (done & 1) : 0;
      const int use = ACC == 2 ? (done >> 1) : done;
      mbar_wait(tmem_full_bar(b_), use & 1);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b_ * ACC_COLS);
      const int ngroups = (kcols + 15) >> 4;
#pragma unroll 1
      for (int g = 0; g < ngroups; ++g) {
        const int c0 = g * 16;
        const int nv = p.K - (k0 + c0);
        uint32_t v[16];
        if (n_iters > 0) {
          tmem_ld16(t_acc + (uint32_t)c0, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) mk[j] = mk_next[j];
        if (g + 1 < ngroups) load_mask(c0 + 16);                  // next group's mask, in flight during this group's math
        if (n_iters > 0) tmem_ld_wait();
        if (g + 1 == ngroups) {
          // the accumulator has been read completely: hand it back before the math and the stores of the last group
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(b_));
        }
        if (!valid) continue;
        if (p.splits > 1) {                                       // raw partial sums; bias / activation / mask in splitk_finish_kernel
          if (n_iters > 0) {
            float* o = p.partial + pix * p.partial_pitch + k0 + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (j < nv)                                         // partial_pitch is padded to the tile, whole groups of 4 exist
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
          }
          continue;
        }
        const int nvl = nv < 16 ? nv : 16;
        int npad = (nvl + 7) & ~7;
        if (k0 + c0 + npad > p.out_pitch) npad = nvl;
        epi16(v, sbias + c0, mk, p.act, p.slope, nvl, npad, p.out_f32, orow + c0 * esize);
      }
      ++done;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// ------------------------------------------------------------------------------------------------
// weight-gradient kernel:  dw[a][tap][b] += sum_{pixels m} small[m][a] * big[pix(m, tap)][b]
// Both operands are [pixels][channels] tensors, i.e. MN-major for the MMA (the reduction runs over pixels).
// A K-block is a BOX (bn images x bp rows x bq columns, bn*bp*bq in {16,32,48,64}) of the small tensor's pixel
// grid, over-covering it where needed: out-of-range pixels of `small` are zero-filled by TMA, which also
// cancels whatever the matching `big` box holds there.  One CTA owns one (a-tile, b-tile, group of `tp` taps) and a
// strided subset of the K-blocks: the `small` box of a K-block is loaded ONCE and multiplied with the `big` boxes of
// all taps of the group (one TMEM accumulator per tap), so `small` is re-read taps/tp times instead of once per tap
// (the per-tap version was bound by L2->SM traffic).  Partial tiles are added into dw with fp32 vector reductions.
// ------------------------------------------------------------------------------------------------
struct WgParams {
  int N, P, Q, A;          // small: [N][P][Q][a_pitch]
  int H, W, B;             // big:   [N][H][W][b_pitch]
  int S, stride, pad, taps;
  int bq, bp, bn;          // K-block box
  int blocks_q, blocks_p, blocks_n;
  int a_tiles, b_tiles;
  uint32_t rows;           // bq*bp*bn
  int tp, tap_groups;      // taps per CTA, ceil(taps / tp)
  int stages;
  uint32_t stage_bytes, tmem_cols;
  float* dw;
};

template <int TILE_N>
__global__ void __launch_bounds__(NUM_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_s,
                                                               const __grid_constant__ CUtensorMap map_g,
                                                               const __grid_constant__ WgParams p) {
  constexpr uint32_t CHUNK_BYTES = 64 * 64 * 2;                  // one 64-channel x 64-pixel chunk (max rows)
  constexpr int NB = TILE_N / 64;                                // B chunks per tap
  const uint32_t STAGE_BYTES = p.stage_bytes;                    // [small chunk 0][small chunk 1][tp x NB big chunks]
  const int STAGES = p.stages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int tg = t % p.tap_groups; t /= p.tap_groups;
  const int bt = t % p.b_tiles;
  const int at = t / p.b_tiles;
  const int tap0 = tg * p.tp;
  const int ntap = p.taps - tap0 < p.tp ? p.taps - tap0 : p.tp;
  const int a0 = at * BLOCK_M, b0 = bt * TILE_N;
  const int n_blocks = p.blocks_q * p.blocks_p * p.blocks_n;
  const int split = blockIdx.y, splits = gridDim.y;
  const int my_blocks = split < n_blocks ? (n_blocks - split + splits - 1) / splits : 0;

  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int st) { return bar_base + 8u * st; };
  auto empty_bar = [&](int st) { return bar_base + 8u * (STAGES + st); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_s);
    prefetch_tmap(&map_g);
    for (int st = 0; st < STAGES; ++st) {
      mbar_init(full_bar(st), 1);
      mbar_init(empty_bar(st), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_rt(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t chunk_bytes = p.rows * 128u;                    // bytes one box deposits

  if (warp == 0) {
    {
      const uint32_t leader = elect_one();             // warp-uniform producer loop, the elected lane issues
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_blocks; ++i) {
        int blk = split + i * splits;
        const int qb = blk % p.blocks_q; blk /= p.blocks_q;
        const int pb = blk % p.blocks_p;
        const int nb = blk / p.blocks_p;
        const int q0 = qb * p.bq, p0 = pb * p.bp, n0 = nb * p.bn;
        mbar_wait(empty_bar(st), ph ^ 1);
        const uint32_t base = smem_u32(smem) + (uint32_t)st * STAGE_BYTES;
        mbar_expect_tx_if(full_bar(st), chunk_bytes * (uint32_t)(2 + ntap * NB), leader);
        tma_load_4d_if(base, &map_s, full_bar(st), a0, q0, p0, n0, leader);
        tma_load_4d_if(base + CHUNK_BYTES, &map_s, full_bar(st), a0 + 64, q0, p0, n0, leader);
        for (int tt = 0; tt < ntap; ++tt) {
          const int tap = tap0 + tt, r = tap / p.S, s = tap - r * p.S;
          const int x = q0 * p.stride - p.pad + s, y = p0 * p.stride - p.pad + r;
#pragma unroll
          for (int j = 0; j < NB; ++j)
            tma_load_4d_if(base + (uint32_t)(2 + tt * NB + j) * CHUNK_BYTES, &map_g, full_bar(st), b0 + 64 * j, x, y, n0, leader);
        }
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_one();             // warp-uniform issue loop, one elected lane issues
      constexpr uint32_t idesc = make_idesc(BLOCK_M, TILE_N, 1, 1);
      // MN-major, 128B swizzle: 64-channel chunks CHUNK_BYTES apart (LBO), 8-pixel groups 1024 B apart (SBO)
      const uint64_t d0 = make_desc(0, CHUNK_BYTES, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), desc_lo = (uint32_t)d0;      // low word without an address
      const int ksteps = (int)p.rows / UMMA_K;
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_blocks; ++i) {
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t base = smem_u32(smem) + (uint32_t)st * STAGE_BYTES;
        const uint32_t a_lo = ((base >> 4) & 0x3FFFu) | desc_lo;
        for (int tt = 0; tt < ntap; ++tt) {
          const uint32_t b_lo = (((base + (uint32_t)(2 + tt * NB) * CHUNK_BYTES) >> 4) & 0x3FFFu) | desc_lo;
          for (int k = 0; k < ksteps; ++k)   // 16 pixels = 2048 B = 128 descriptor units per step
            umma_bf16_lo(tmem_base + (uint32_t)(tt * TILE_N), a_lo + 128 * k, b_lo + 128 * k, desc_hi, idesc, (i | k) ? 1u : 0u,
                         leader);
        }
        umma_commit_if(empty_bar(st), leader);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
      if (my_blocks > 0) umma_commit_if(tmem_full_bar, leader);
      else if (lane == 0) mbar_arrive(tmem_full_bar);
    }
  } else {
    const int q4 = warp & 3;
    const int a = a0 + q4 * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (my_blocks > 0) {
      // All splits of one tile finish together and add into the SAME addresses: in accumulator order every address received a
      // burst of `splits` reductions at once (the L2 atomic unit serialises per address; the first-layer kernel spent as long
      // in this flush as in its main loop, ncu).  Each split starts at its own (tap, column group).
      constexpr int GROUPS = TILE_N / 16;
      const int units = ntap * GROUPS;
#pragma unroll 1
      for (int uu = 0; uu < units; ++uu) {
        {
          const int u = (uu + split) % units;
          const int tt = u / GROUPS, c0 = (u - tt * GROUPS) * 16;
          if (b0 + c0 >= p.B) continue;
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(tt * TILE_N + c0), v);
          tmem_ld_wait();
          if (a < p.A) {
            float* o = p.dw + ((int64_t)a * p.taps + tap0 + tt) * p.B + b0 + c0;
            if (b0 + c0 + 16 <= p.B && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
              // four 16-byte vector reductions instead of sixteen scalar atomics (the split-K epilogue is L2-atomic bound)
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (b0 + c0 + j < p.B) atomicAdd(o + j, __uint_as_float(v[j]));
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

// split-K finish: dst = act(partial + bias) * mask (bf16 or fp32), one thread per (pixel, 4 channels)
__global__ void splitk_finish_kernel(const float* __restrict__ partial, int partial_pitch, int64_t pixels, int pixels_per_sample, int K,
                                     const float* __restrict__ bias, int act, float slope, const float* __restrict__ mask, int mask_pitch,
                                     void* dst, int out_pitch, int out_f32) {
  const int kq = (K + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels * kq; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / kq;
    const int k0 = (int)(i - pix * kq) * 4;
    const float4 v4 = *reinterpret_cast<const float4*>(partial + pix * partial_pitch + k0);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    const float* mrow = mask ? mask + (pix / pixels_per_sample) * mask_pitch : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      if (k >= K) break;
      float f = icf::apply_act(v[j] + (bias ? bias[k] : 0.f), act, slope);
      if (mrow) f *= mrow[k];
      if (out_f32) reinterpret_cast<float*>(dst)[pix * out_pitch + k] = f;
      else reinterpret_cast<__nv_bfloat16*>(dst)[pix * out_pitch + k] = __float2bfloat16_rn(f);
    }
  }
}

template <int TILE_N, int STAGES, int ACC>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int64_t tiles, int ctas_per_sm, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BLOCK_K * 2 + TILE_N * BLOCK_K * 2) + 1024 + 256 + 1024;
  // the attribute is per device and idempotent: set it on every device this thread launches on, race-free across threads
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(conv_tc_kernel<TILE_N, STAGES, ACC>), smem, "tensor-core conv")) return r;
  int64_t grid = (int64_t)tc_sm_count() * ctas_per_sm;       // one resident wave of persistent CTAs
  if (grid > tiles) grid = tiles;
  conv_tc_kernel<TILE_N, STAGES, ACC><<<(unsigned)grid, NUM_THREADS, smem, st>>>(ma, mb, p);
  return icf::check_launch("conv_tc");
}

}  // namespace

// partial != NULL: split-K allowed — a zeroed fp32 scratch of partial_elems floats ([N*P*Q][K rounded up to the tile]); used when the
// tile grid alone would leave most SMs idle.  Returns -1 as icf_tc_conv_forward does.
static int tc_forward_impl(const icf_conv_args* a, cudaStream_t st, float* partial, int64_t partial_elems) {
  // shapes the tensor-core path takes; everything else runs on the direct / SIMT kernels
  if (a->dtype != ICF_BF16 || a->accumulate) return -1;
  // small-channel first / last layers ride the same kernel: TMA zero-fills the missing channels of the 64-wide
  // K chunk, so they cost tensor time but only their real bytes of HBM traffic (they are HBM-bound layers)
  if ((a->in_pitch & 7) || (a->w_pitch & 7)) return -1;
  if (a->win > 1 && a->pad != 0) return -1;   // the folded form reads a pre-padded tensor
  if ((reinterpret_cast<uintptr_t>(a->src) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15)) return -1;
  const int taps = a->R * a->S;
  if (taps > MAX_TAPS || a->stride > 4 || a->pad > 100) return -1;
  const bool gather = a->form == ICF_FORM_GATHER;

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.H = a->H; p.W = a->W; p.C = a->C;
  p.P = a->P; p.Q = a->Q; p.K = a->K; p.out_pitch = a->out_pitch;
  p.sstep = gather ? a->stride : 1;
  p.ostep = gather ? 1 : a->stride;
  p.n_classes = p.ostep * p.ostep;
  if (p.n_classes > MAX_CLASSES) return -1;
  // tap tables
  for (int cls = 0; cls < p.n_classes; ++cls) {
    const int py = cls / p.ostep, px = cls % p.ostep;
    int n = 0;
    for (int r = 0; r < a->R; ++r)
      for (int s = 0; s < a->S; ++s) {
        int dy, dx;
        if (gather) {
          dy = r - a->pad;
          dx = s - a->pad;
        } else {
          const int uy = py + a->pad - r, ux = px + a->pad - s;
          if (((uy % a->stride) + a->stride) % a->stride != 0 || ((ux % a->stride) + a->stride) % a->stride != 0) continue;
          dy = uy / a->stride;     // exact (uy is a multiple of stride, possibly negative)
          dx = ux / a->stride;
        }
        if (dy < -128 || dy > 127 || dx < -128 || dx > 127) return -1;
        p.tt.tap[cls][n] = (int8_t)(r * a->S + s);
        p.tt.dy[cls][n] = (int8_t)dy;
        p.tt.dx[cls][n] = (int8_t)dx;
        ++n;
      }
    p.tt.ntaps[cls] = (int16_t)n;
  }
  // tile box over the (largest) class index space
  const int Pi = (a->P + p.ostep - 1) / p.ostep, Qj = (a->Q + p.ostep - 1) / p.ostep;
  if ((int64_t)Pi * Qj <= BLOCK_M) {
    p.ti = Pi; p.tj = Qj;
    p.tn = BLOCK_M / (Pi * Qj);
    if (p.tn > a->N) p.tn = a->N;
  } else if (Qj <= BLOCK_M) {
    p.tj = Qj; p.ti = BLOCK_M / Qj; p.tn = 1;
  } else {
    p.tj = BLOCK_M; p.ti = 1; p.tn = 1;
  }
  if (p.tj * p.sstep > 256 || p.ti * p.sstep > 256) return -1;
  p.tiles_i = icf::cdiv(Pi, p.ti);
  p.tiles_j = icf::cdiv(Qj, p.tj);
  p.tiles_n = icf::cdiv(a->N, p.tn);
  p.kchunks = icf::cdiv(a->C, BLOCK_K);
  p.w_pitch = a->w_pitch;
  p.a_tx_bytes = (uint32_t)(p.tn * p.ti * p.tj) * (BLOCK_K * 2);
  p.act = a->act; p.slope = a->slope; p.out_f32 = a->out_f32; p.mask_pitch = a->mask_pitch;
  p.bias = a->bias; p.mask = a->out_mask; p.dst = a->dst;

  int tile_n = a->K > 128 ? 256 : (a->K > 64 ? 128 : (a->K > 32 ? 64 : 32));
  {
    // 256-wide tiles on a grid that fills less than ~3/4 of the SMs (the 1x1-spatial 512 -> 512 layers at batch 4096: 64 CTAs
    // on 148 SMs) leave the machine idle: twice as many 128-wide tiles (2 CTAs per SM: 96 KB ring, 128 TMEM columns) run the
    // same work on twice the SMs.  ICF_TC_TILE=128|256 forces either (tuning aid).
    static const int tile_env = []() { const char* e = getenv("ICF_TC_TILE"); return e ? atoi(e) : 0; }();
    const int64_t g256 = (int64_t)p.n_classes * p.tiles_n * p.tiles_i * p.tiles_j * icf::cdiv(a->K, 256);
    if (tile_n == 256 && (tile_env == 128 || (tile_env != 256 && g256 * 4 < (int64_t)tc_sm_count() * 3))) tile_n = 128;
  }
  p.tiles_k = icf::cdiv(a->K, tile_n);
  int64_t grid = (int64_t)p.n_classes * p.tiles_n * p.tiles_i * p.tiles_j * p.tiles_k;
  if (grid > 0x7FFFFFFF) return -1;
  p.splits = 1;
  if (partial && !a->stats) {
    // split-K: a tile grid much smaller than the machine (the 1024-channel layers of the spectrogram families at batch 32-128:
    // 2-8 tiles whose K loop walks 9-25 taps x 16 channel chunks of a 10-50 MB weight) is bound by what ONE SM can pull from L2
    const int64_t max_iters = (int64_t)taps * p.kchunks, cap = (int64_t)tc_sm_count() * 2;
    const int pitch = p.tiles_k * tile_n;
    int64_t sp = cap / grid;
    if (sp > max_iters / 4) sp = max_iters / 4;                    // at least ~4 ring stages per work item
    if (sp >= 2 && grid * 4 <= (int64_t)tc_sm_count() && (int64_t)a->N * a->P * a->Q * pitch <= partial_elems && (reinterpret_cast<uintptr_t>(partial) & 15) == 0) {
      p.splits = (int)sp;
      p.partial = partial;
      p.partial_pitch = pitch;
      grid *= sp;
    }
  }
  if (partial && p.splits == 1) return -1;                          // nothing to split: the caller goes through the normal dispatch

  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->in_pitch * 2, (cuuint64_t)a->W * a->in_pitch * 2,
                         (cuuint64_t)a->H * a->W * a->in_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(p.tj * p.sstep), (cuuint32_t)(p.ti * p.sstep), (cuuint32_t)p.tn};
    cuuint32_t est[4] = {1, (cuuint32_t)p.sstep, (cuuint32_t)p.sstep, 1};
    if (int r = encode_map(&ma, a->src, 4, dims, str, box, est)) return r;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * a->w_pitch, (cuuint64_t)a->w_rows};
    cuuint64_t str[1] = {(cuuint64_t)taps * a->w_pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)tile_n};
    cuuint32_t est[2] = {1, 1};
    if (int r = encode_map(&mb, a->w, 2, dims, str, box, est)) return r;
  }
  int r;
  switch (tile_n) {
    case 256: {
      // 2 stages = 96 KB and one 256-column accumulator: two CTAs per SM, one CTA's epilogue overlaps the other's main loop.
      // Measured (round 1, one tile per CTA): 8-18 % faster when the grid is at least ~1.5 waves (G.layers.2 fprop, the 256-wide
      // dgrads, the 771->4608 GEMM), 20-40 % slower on small grids, which keep the 4-stage ring with one CTA per SM
      // (ICF_TC_256_STAGES=2|4 forces either)
      static const int st_env = []() { const char* e = getenv("ICF_TC_256_STAGES"); return e ? atoi(e) : 0; }();
      const bool two = st_env == 2 || (st_env != 4 && grid >= 220);
      r = two ? launch_tc<256, 2, 1>(ma, mb, p, grid, 2, st) : launch_tc<256, 4, 2>(ma, mb, p, grid, 1, st);
      break;
    }
    case 128: r = launch_tc<128, 3, 2>(ma, mb, p, grid, 2, st); break;
    case 64: r = launch_tc<64, 4, 2>(ma, mb, p, grid, 2, st); break;
    default: r = launch_tc<32, 4, 2>(ma, mb, p, grid, 2, st); break;
  }
  if (r) return r;
  if (p.splits > 1) {
    const int64_t pixels = (int64_t)a->N * a->P * a->Q, work = pixels * ((a->K + 3) / 4);
    int64_t blocks = (work + 255) / 256;
    if (blocks > (int64_t)tc_sm_count() * 8) blocks = (int64_t)tc_sm_count() * 8;
    splitk_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(p.partial, p.partial_pitch, pixels, a->P * a->Q, a->K, a->bias, a->act, a->slope,
                                                          a->out_mask, a->mask_pitch, a->dst, a->out_pitch, a->out_f32);
    return icf::check_launch("splitk_finish");
  }
  if (a->stats) {
    if (a->out_f32) { icf::set_error("tensor-core conv: BatchNorm statistics need a bf16 destination"); return 1; }
    return icf_launch_col_stats(a->dst, ICF_BF16, a->out_pitch, (int64_t)a->N * a->P * a->Q, a->K, a->stats, st);
  }
  return 0;
}

int icf_tc_conv_forward(const icf_conv_args* a, cudaStream_t st) { return tc_forward_impl(a, st, nullptr, 0); }
int icf_tc_conv_forward_splitk(const icf_conv_args* a, cudaStream_t st, float* partial, int64_t partial_elems) {
  return tc_forward_impl(a, st, partial, partial_elems);
}


namespace {

template <int TILE_N>
int launch_wg(const CUtensorMap& ms, const CUtensorMap& mg, const WgParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(wgrad_tc_kernel<TILE_N>), smem, "tensor-core wgrad")) return r;
  wgrad_tc_kernel<TILE_N><<<grid, NUM_THREADS, smem, st>>>(ms, mg, p);
  return icf::check_launch("wgrad_tc");
}

// K-block box (bq, bp, bn) over the small tensor's pixel grid: rows = bq*bp*bn in {16,32,48,64}, chosen to waste
// as few zero-filled rows as possible
void pick_kblock(int N, int P, int Q, int* bq_, int* bp_, int* bn_) {
  double best = -1.0;
  int best_rows = 0;
  for (int bq = 1; bq <= 64; ++bq) {
    if (bq > Q + 15) break;
    for (int bp = 1; bp * bq <= 64; ++bp) {
      if (bp > P + 15) break;
      for (int bn = 1; bn * bp * bq <= 64; ++bn) {
        if (bn > N && bn > 1) break;
        const int rows = bq * bp * bn;
        if (rows % 16) continue;
        const double covered = (double)icf::cdiv(Q, bq) * bq * icf::cdiv(P, bp) * bp * icf::cdiv(N, bn) * bn;
        const double eff = (double)N * P * Q / covered;
        if (eff > best + 1e-9 || (eff > best - 1e-9 && rows > best_rows)) {
          best = eff; best_rows = rows;
          *bq_ = bq; *bp_ = bp; *bn_ = bn;
        }
      }
    }
  }
}

}  // namespace

namespace {
// returns 0 = launched (or, with `plan`, described), -1 = not this kernel's case, > 0 = error
int wgrad_impl(const icf_wgrad_args* a, cudaStream_t st, int32_t* plan, int32_t plan_words) {
  if (a->dtype != ICF_BF16) return -1;
  if ((a->a_pitch & 7) || (a->b_pitch & 7)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->small_t) & 15) || (reinterpret_cast<uintptr_t>(a->big_t) & 15)) return -1;
  if (a->stride > 4) return -1;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.N = a->N; p.P = a->P; p.Q = a->Q; p.A = a->A;
  p.H = a->H; p.W = a->W; p.B = a->B;
  p.S = a->S; p.stride = a->stride; p.pad = a->pad; p.taps = a->R * a->S;
  p.bq = p.bp = p.bn = 0;
  pick_kblock(a->N, a->P, a->Q, &p.bq, &p.bp, &p.bn);
  if (p.bq == 0) return -1;
  if (p.bq * a->stride > 256 || p.bp * a->stride > 256) return -1;
  p.rows = (uint32_t)(p.bq * p.bp * p.bn);
  p.blocks_q = icf::cdiv(a->Q, p.bq);
  p.blocks_p = icf::cdiv(a->P, p.bp);
  p.blocks_n = icf::cdiv(a->N, p.bn);
  const int tile_n = a->B > 128 ? 256 : (a->B > 64 ? 128 : 64);
  p.a_tiles = icf::cdiv(a->A, BLOCK_M);
  p.b_tiles = icf::cdiv(a->B, tile_n);
  p.dw = a->dw;
  const int64_t n_blocks = (int64_t)p.blocks_q * p.blocks_p * p.blocks_n;
  if (n_blocks > 0x7FFFFFFF) return -1;
  // taps per CTA: as many as keep >= 3 pipeline stages in ~200 KB and the accumulators in 512 TMEM columns, preferring a
  // divisor of the tap count (ICF_WG_TP overrides, tuning aid)
  {
    const int nb = tile_n / 64, max_tp = 512 / tile_n;
    int best = 1;
    for (int tp = 1; tp <= max_tp && tp <= p.taps && tp <= 4; ++tp) {
      const int stage = (2 + tp * nb) * 8192;
      if (200 * 1024 / stage < 3) break;
      if (p.taps % tp == 0 || p.taps % best != 0) best = tp;
    }
    static const int tp_env = []() { const char* e = getenv("ICF_WG_TP"); return e ? atoi(e) : 0; }();
    if (tp_env > 0) {
      best = tp_env > max_tp ? max_tp : tp_env;
      if (best > p.taps) best = p.taps;
      while (best > 1 && 200 * 1024 / ((2 + best * nb) * 8192) < 2) --best;
    }
    p.tp = best;
    p.tap_groups = icf::cdiv(p.taps, p.tp);
    p.stage_bytes = (uint32_t)((2 + p.tp * nb) * 8192);
    p.stages = 200 * 1024 / (int)p.stage_bytes;
    if (p.stages > 4) p.stages = 4;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.tp * tile_n)) cols <<= 1;
    p.tmem_cols = cols;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  const int ctas_per_sm = (smem <= 110 * 1024 && p.tmem_cols <= 256) ? 2 : 1;
  const int64_t tiles = (int64_t)p.a_tiles * p.b_tiles * p.tap_groups;
  int64_t splits = ((int64_t)icf::sm_count() * ctas_per_sm) / tiles;          // one resident wave: never a ragged second one
  if (splits > n_blocks) splits = n_blocks;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  if (plan) {     // icf_wgrad_plan: describe instead of launching (layout in include/icf.h)
    if (plan_words < 16) { icf::set_error("icf_wgrad_plan: out needs 16 words"); return 1; }
    const int32_t v[16] = {tile_n, p.tp, p.tap_groups, p.stages, (int32_t)p.stage_bytes, (int32_t)p.tmem_cols, (int32_t)smem,
                           ctas_per_sm, (int32_t)tiles, (int32_t)splits, p.bq, p.bp, p.bn, (int32_t)p.rows, (int32_t)n_blocks,
                           p.taps};
    for (int i = 0; i < 16; ++i) plan[i] = v[i];
    return 0;
  }
  CUtensorMap ms, mg;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->A, (cuuint64_t)a->Q, (cuuint64_t)a->P, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->a_pitch * 2, (cuuint64_t)a->Q * a->a_pitch * 2,
                         (cuuint64_t)a->P * a->Q * a->a_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.bq, (cuuint32_t)p.bp, (cuuint32_t)p.bn};
    cuuint32_t est[4] = {1, 1, 1, 1};
    if (int r = encode_map(&ms, a->small_t, 4, dims, str, box, est)) return r;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->B, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->N};
    cuuint64_t str[3] = {(cuuint64_t)a->b_pitch * 2, (cuuint64_t)a->W * a->b_pitch * 2,
                         (cuuint64_t)a->H * a->W * a->b_pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.bq * a->stride), (cuuint32_t)(p.bp * a->stride), (cuuint32_t)p.bn};
    cuuint32_t est[4] = {1, (cuuint32_t)a->stride, (cuuint32_t)a->stride, 1};
    if (int r = encode_map(&mg, a->big_t, 4, dims, str, box, est)) return r;
  }
  dim3 grid((unsigned)tiles, (unsigned)splits);
  switch (tile_n) {
    case 256: return launch_wg<256>(ms, mg, p, grid, smem, st);
    case 128: return launch_wg<128>(ms, mg, p, grid, smem, st);
    default: return launch_wg<64>(ms, mg, p, grid, smem, st);
  }
}
}  // namespace

int icf_tc_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st) { return wgrad_impl(a, st, nullptr, 0); }

int icf_wgrad_plan(const icf_wgrad_args* a, int32_t* out, int32_t out_words) {
  ICF_REQUIRE(a && out && out_words > 0, "icf_wgrad_plan: bad arguments");
  return wgrad_impl(a, nullptr, out, out_words);
}

