// Weight-stationary, row-streaming implicit-GEMM convolution for sm_100a (tcgen05.mma, TMEM accumulators, TMA).
//
// Serves the layers whose packed weights fit in shared memory (the small-channel first / last layers and the
// per-parity-class slabs of the stride-2 transposed convolutions) — the layers that are HBM-bound on the roofline
// and that an im2col-style kernel turns into L2-bound ones by fetching every activation once per filter tap.
//
//   * A persistent CTA owns one (parity class, 64-or-fewer output-channel tile): it loads that weight slab ONCE.
//   * Its work items are "columns": XG adjacent output x positions of NG images (XG*NG = 128 = the MMA M).
//     Operand rows are stored batch-innermost, [x][image][64 channels] in 128B-swizzled K-major form, so the
//     operand of filter tap (dy, dx) is the same shared-memory tile shifted by dx*NG rows — a multiple of the
//     1024-byte swizzle atom, i.e. a plain descriptor start-address offset.  Every source row is therefore
//     fetched from L2 once (plus the x halo), not once per tap.
//   * Source rows stream through a ring of slots in increasing y.  Row y feeds every output row i with
//     i*sstep + dy == y, each accumulating in its own TMEM accumulator (a ring of n_acc); an output row retires
//     to the epilogue warps as soon as its last source row has been issued, so TMEM -> bias/activation/Dropout2d
//     -> global stores overlap the MMAs of the following rows.
//   * Stride-2 gathers read each source row as two x-parity sub-rows (TMA element stride 2); transposed forms are
//     unit-stride gathers per output-parity class (no zero insertion).
//
// Warp roles: warp 0 = TMA producer, warps 1..WS_ISSUERS = MMA issuers (small-N MMAs are issue-bound, so the
// output rows are dealt round-robin to several issuing warps; warp 1 also owns the TMEM allocation), last four
// warps = epilogue.
#include "icf_epilogue.cuh"

#include <cstdlib>
#include <cstring>

namespace {

using namespace icf_tc;

constexpr int WS_ISSUERS = 3;                       // MMA-issuing warps: output row i belongs to issuer i % WS_ISSUERS
// (12 warps = 3 per SM sub-partition, which leaves 168 registers per thread; a 13th warp would cap it at 128)
constexpr int WS_EPI_GROUPS = 2;                    // epilogue warpgroups (4 warps each): output row g belongs to group g % 2
constexpr int WS_EPI0 = 32 * (1 + WS_ISSUERS);      // first epilogue thread
constexpr int WS_THREADS = WS_EPI0 + 128 * WS_EPI_GROUPS;
constexpr int WS_MAX_CLASSES = 4, WS_MAX_TAPS = 25, WS_MAX_SUBS = 2, WS_MAX_ACC = 16, WS_MAX_SLOTS = 8;
constexpr int WS_SMEM_BUDGET = 220 * 1024;
constexpr int WS_MAX_GROUPS = 8;
constexpr int WS_PROG_WORDS = 4096;                 // issuer schedules of all classes (4-byte actions), in the kernel parameters
                                                    // (one to two words per source row and issuer: 512-row images need ~2400)

struct WsTap {
  int16_t dy;        // source row = i*sstep + dy
  int16_t widx;      // tap index r*S+s into the packed weights
  uint32_t a_off16;  // offset (16-byte units) of this tap's operand inside a slot (sub-row + x shift)
};

struct WsGroup {     // taps that share dy (one filter row): they feed the same output row from one source row
  int16_t dy;
  int16_t dprev;     // dy minus the next smaller dy that feeds the same output rows (32767: none) — the group opens
                     // its output row (accumulate = 0) iff the previous contributor's source row lies above ylo
  int8_t first, count;
};

struct WsClass {
  int Pi, Qj, py, px;
  int ntaps, ylo, yhi, dymax;
  int x0[WS_MAX_SUBS];         // TMA x start = j0*sstep + x0[sub]
  int tiles_x, cta_begin, cta_count;
  int ngroups;
  int prog_off[WS_ISSUERS + 1];  // this class's issuer schedules inside WsParams::prog (contiguous, issuer after issuer)
  WsGroup grp[WS_MAX_GROUPS];
  WsTap taps[WS_MAX_TAPS];
};

struct WsParams {
  int N, P, Q, K, out_pitch;
  int sstep, ostep, n_classes;
  int XG, NG, ng_shift;
  int kchunks, kdepth_last, w_pitch;
  int rowb;                    // bytes per operand row in shared memory: 128 (64-channel K chunk, 128B swizzle) or
                               // 64 (C <= 32: 32-channel chunk, 64B swizzle — half the staging traffic and footprint)
  int nsub, n_slots, n_acc, acc_shift;
  int tiles_n, tiles_k;
  uint32_t slot_bytes, sub_bytes, kc_bytes, slab_bytes, tmem_cols;
  int act;
  float slope;
  int out_f32, mask_pitch;
  const float* bias;
  const float* mask;
  void* dst;
  float* stats;                // [2][K] BatchNorm statistics of the stored values (+=), fused into the TMA-store epilogue
  int tma_store;               // 1: epilogue stages bf16 rows in shared memory and leaves through a TMA tensor store
  // Issuer schedules, built on the host: which output rows a source row feeds, which of them an issuer owns, which
  // open or complete an accumulator is the same for every column.  One 4-byte action per MMA chain:
  //   bits 0-7 output row i | 8-12 first tap | 13-17 tap count | 18-19 kind (0 chain, 1 accumulator complete, 2 nothing)
  //   | 20 chain opens the accumulator | 21 last action of this source row
  uint32_t prog[WS_PROG_WORDS];
  unsigned long long* dbg;     // optional per-role cycle counters of CTA 0 (env ICF_WS_DEBUG), else NULL
  WsClass cls[WS_MAX_CLASSES];
};

// KD > 0: C spans KD 16-wide MMA steps (<= 8, i.e. up to two 64-channel chunks), issue loop fully unrolled;
// KD == 0: generic
// Per-role cycle counters (CTA 0) are a build-time option: `ICF_WS_INSTRUMENT=1 python build.py --force`, then run with
// ICF_WS_DEBUG=1.  The production issue loop must stay small enough for the L0 instruction cache.
#ifdef ICF_WS_INSTRUMENT
#define WS_TIMED_WAIT(counter, bar, par)                   \
  do {                                                     \
    if (dbg_on) {                                          \
      const long long t_ = clock64();                      \
      mbar_wait(bar, par);                                 \
      counter += clock64() - t_;                           \
    } else {                                               \
      mbar_wait(bar, par);                                 \
    }                                                      \
  } while (0)
#else
#define WS_TIMED_WAIT(counter, bar, par) mbar_wait(bar, par)
#endif

__device__ __forceinline__ void epi_bar(int grp) {   // named barrier of one epilogue warpgroup (ids 1, 2)
  if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}

template <int TILE_N, int KD>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_ws_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b,
                                                                const __grid_constant__ CUtensorMap map_o,
                                                                const __grid_constant__ WsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wslab = smem;
  uint8_t* slots = smem + p.slab_bytes;
  uint8_t* stage = slots + (size_t)p.n_slots * p.slot_bytes;      // 2 x [128 rows][TILE_N] bf16 (TMA-store epilogue)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * 128 * TILE_N * 2);
  // barrier layout: slot_full[8] slot_empty[8] acc_full[16] acc_empty[16] w_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WS_MAX_SLOTS + 2 * WS_MAX_ACC + 1);
  float* sbias = reinterpret_cast<float*>(bars + 64);          // TILE_N floats, 512 B past the barriers
  uint8_t* prog = reinterpret_cast<uint8_t*>(bars) + 1024;     // [tap table 256 B][this class's issuer schedules]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cls = 0;
  while (cls + 1 < p.n_classes && (int)blockIdx.x >= p.cls[cls].cta_begin + p.cls[cls].cta_count) ++cls;
  const WsClass& cl = p.cls[cls];
  const int local = (int)blockIdx.x - cl.cta_begin;
  const int kt = local % p.tiles_k, r0 = local / p.tiles_k, rstep = cl.cta_count / p.tiles_k;
  const int k0 = kt * TILE_N;
  const int ncols = p.tiles_n * cl.tiles_x;

  const uint32_t bar_base = smem_u32(bars);
  auto slot_full = [&](int s) { return bar_base + 8u * s; };
  auto slot_empty = [&](int s) { return bar_base + 8u * (WS_MAX_SLOTS + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * WS_MAX_SLOTS + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * WS_MAX_SLOTS + WS_MAX_ACC + a); };
  const uint32_t w_full = bar_base + 8u * (2 * WS_MAX_SLOTS + 2 * WS_MAX_ACC);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (p.tma_store) prefetch_tmap(&map_o);
    for (int s = 0; s < WS_MAX_SLOTS; ++s) {
      mbar_init(slot_full(s), 1);
      mbar_init(slot_empty(s), WS_ISSUERS);
    }
    for (int a = 0; a < WS_MAX_ACC; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);      // one arrival per epilogue warp
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  // this class's tap table and issuer schedules -> shared memory, one word per thread (parameter space is read with
  // independent loads here; the issue loop then pays one shared-memory load per action instead of indexed
  // parameter loads and skipped groups, which cost ~1000 cycles per source row)
  for (int t = (int)threadIdx.x; t < cl.ntaps; t += WS_THREADS)
    sts64(smem_u32(prog) + 8u * t, cl.taps[t].a_off16, (uint32_t)(t * p.kchunks) * (((uint32_t)TILE_N * (uint32_t)p.rowb) >> 4));
  for (int w = cl.prog_off[0] + (int)threadIdx.x; w < cl.prog_off[WS_ISSUERS]; w += WS_THREADS)
    reinterpret_cast<uint32_t*>(prog + 256)[w - cl.prog_off[0]] = p.prog[w];
  if (warp == 1) tmem_alloc_rt(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t slab_addr = smem_u32(wslab), slots_addr = smem_u32(slots);
#ifdef ICF_WS_INSTRUMENT
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
#else
  constexpr bool dbg_on = false;
#endif
  long long w0 = 0, w1 = 0;                 // cycles spent in this role's two kinds of barrier waits
  long long w2 = 0, w3 = 0, n_grp = 0;      // issuers: cycles in the MMA chains / in the commits, chains issued
  const long long t_begin = dbg_on ? clock64() : 0;
  const uint32_t W_BLOCK = (uint32_t)TILE_N * (uint32_t)p.rowb;      // one (tap, K chunk) block of the weight slab

  if (warp == 0) {
    // ===== TMA producer: weight slab once, then the source rows of every column =====
    {
      const uint32_t leader = elect_one();             // warp-uniform loop, the elected lane issues
      mbar_expect_tx_if(w_full, (uint32_t)(cl.ntaps * p.kchunks) * W_BLOCK, leader);
      for (int t = 0; t < cl.ntaps; ++t)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d_if(slab_addr + (uint32_t)(t * p.kchunks + kc) * W_BLOCK, &map_b, w_full,
                         cl.taps[t].widx * p.w_pitch + kc * (p.rowb >> 1), k0, leader);
      int s = 0;
      uint32_t ph = 0;
      for (int col = r0; col < ncols; col += rstep) {
        const int jt = col % cl.tiles_x, nt = col / cl.tiles_x;
        const int n0 = nt * p.NG, xs = jt * p.XG * p.sstep;
        for (int y = cl.ylo; y <= cl.yhi; ++y) {
          WS_TIMED_WAIT(w0, slot_empty(s), ph ^ 1);
          mbar_expect_tx_if(slot_full(s), p.slot_bytes, leader);
          const uint32_t base = slots_addr + (uint32_t)s * p.slot_bytes;
          for (int sub = 0; sub < p.nsub; ++sub)
            for (int kc = 0; kc < p.kchunks; ++kc)
              tma_load_4d_if(base + sub * p.sub_bytes + kc * p.kc_bytes, &map_a, slot_full(s), kc * (p.rowb >> 1), n0,
                             xs + cl.x0[sub], y, leader);
          if (++s == p.n_slots) { s = 0; ph ^= 1; }
        }
      }
      if (dbg_on && lane == 0) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = w0; }
    }
  } else if (warp <= WS_ISSUERS) {
    // ===== MMA issuers: whole warp runs the control flow, one elected lane issues.  Issuer wi owns the output rows
    // i with i % WS_ISSUERS == wi, so all MMAs into one accumulator come from one thread (ordered). =====
    constexpr uint32_t idesc = make_idesc(128, TILE_N, 0, 0);
    const uint32_t WB16 = W_BLOCK >> 4;
    const int wi = warp - 1;
    const uint32_t leader = elect_one();
    const int Pi = cl.Pi, ylo = cl.ylo, yhi = cl.yhi;
    const uint32_t tap_tab = smem_u32(prog), ent_tab = tap_tab + 256u + 4u * (uint32_t)(cl.prog_off[wi] - cl.prog_off[0]);
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint64_t d0 = make_desc_sw(0, 16, 8u * (uint32_t)p.rowb, p.rowb == 64 ? 4u : 2u);
    const uint32_t desc_hi = (uint32_t)(d0 >> 32), lbo_lo = (uint32_t)d0;     // low word without an address
    const uint32_t b_lo0 = ((slab_addr >> 4) & 0x3FFFu) | lbo_lo;
    const uint32_t kc16 = p.kc_bytes >> 4;
    const int amask = p.n_acc - 1;
    int s = 0;
    uint32_t ph = 0;
    int g_base = 0;                       // sequence number of this column's output row 0
    for (int col = r0; col < ncols; col += rstep) {
      uint32_t pc = ent_tab;
      uint32_t nxt = lds32(pc);
      for (int y = ylo; y <= yhi; ++y) {
        WS_TIMED_WAIT(w0, slot_full(s), ph);
        tc_fence_after();
        const uint32_t a_lo0 = (((slots_addr + (uint32_t)s * p.slot_bytes) >> 4) & 0x3FFFu) | lbo_lo;
        uint32_t flags;
#pragma unroll 1                                   // keep the issue loop small: it has to live in the L0 instruction cache
        do {
          const uint32_t e = nxt;
          pc += 4;
          nxt = lds32(pc);                         // every schedule ends with a spare word
          flags = e >> 18;
          const int g = g_base + (int)(e & 0xFFu), acc = g & amask;
          if ((flags & 3u) == 0u) {
            const long long t_mma = dbg_on ? clock64() : 0;
            uint32_t accum = 1u;
            if (flags & 4u) {
              WS_TIMED_WAIT(w1, acc_empty(acc), ((uint32_t)(g >> p.acc_shift) & 1u) ^ 1u);
              tc_fence_after();
              accum = 0u;
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * TILE_N;
            uint32_t tp = tap_tab + ((e >> 5) & 0xF8u);
            const uint32_t tp_end = tp + ((e >> 10) & 0xF8u);
#pragma unroll 1
            for (; tp < tp_end; tp += 8) {
              const uint2 tt = lds64(tp);
              const uint32_t a_lo = a_lo0 + tt.x, b_lo = b_lo0 + tt.y;
              if (KD > 0) {              // KD 16-wide K steps, chunk boundary every 4 (fully unrolled)
#pragma unroll
                for (int k = 0; k < KD; ++k) {
                  umma_bf16_lo(d_tmem, a_lo + (k >> 2) * kc16 + 2 * (k & 3), b_lo + (k >> 2) * WB16 + 2 * (k & 3), desc_hi,
                               idesc, accum, leader);
                  accum = 1u;
                }
              } else {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                  const int nk = (kc == p.kchunks - 1) ? p.kdepth_last : 4;
                  for (int k = 0; k < nk; ++k) {
                    umma_bf16_lo(d_tmem, a_lo + kc * kc16 + 2 * k, b_lo + kc * WB16 + 2 * k, desc_hi, idesc, accum, leader);
                    accum = 1u;
                  }
                }
              }
            }
            if (dbg_on) { w2 += clock64() - t_mma; ++n_grp; }
          } else if ((flags & 3u) == 1u) {
            umma_commit_if(acc_full(acc), leader);      // every MMA into this accumulator has been issued (by this thread)
          }
        } while (!(flags & 8u));
        umma_commit_if(slot_empty(s), leader);       // arrives once this issuer's MMAs on the slot have retired
        if (++s == p.n_slots) { s = 0; ph ^= 1; }
      }
      g_base += Pi;
    }
    if (dbg_on && lane == 0 && wi == 0) { p.dbg[2] = clock64() - t_begin; p.dbg[3] = w0; p.dbg[4] = w1; p.dbg[8] = w2; p.dbg[9] = w3; p.dbg[10] = n_grp; }
  } else {
    // ===== epilogue: TMEM lane m = x_local*NG + n_local =====
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int x_local = m >> p.ng_shift, n_local = m & (p.NG - 1);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // bias tile -> shared memory (zero where absent / beyond K), visible to the four epilogue warps
    // Two epilogue warpgroups alternate output rows: a single warp per SM sub-partition pays every instruction and
    // barrier at full latency (~1200 cycles per 128x32 row measured), two rows in flight hide half of it.  Each
    // group owns one staging buffer and one named barrier.
    const int grp = ((int)threadIdx.x - WS_EPI0) >> 7;
    for (int j = (int)threadIdx.x - WS_EPI0; j < TILE_N; j += 128 * WS_EPI_GROUPS) sbias[j] = (p.bias && k0 + j < p.K) ? __ldg(p.bias + k0 + j) : 0.f;
    asm volatile("bar.sync 3, %0;" ::"n"(128 * WS_EPI_GROUPS) : "memory");
    const int esize = p.out_f32 ? 4 : 2;
    // fused BatchNorm statistics: thread -> one 8-channel chunk (et % CH) of the staged tile and every (128/CH)-th row
    constexpr int CH = TILE_N / 8;
    const int et = ((int)threadIdx.x - WS_EPI0) & 127;
    const int st_ch = et % CH, st_r0 = et / CH;
    float st_sum[8], st_sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) st_sum[j] = st_sq[j] = 0.f;
    int g = 0;
    for (int col = r0; col < ncols; col += rstep) {
      const int jt = col % cl.tiles_x, nt = col / cl.tiles_x;
      const int n = nt * p.NG + n_local, jj = jt * p.XG + x_local;
      const bool valid = n < p.N && jj < cl.Qj;
      float mk[TILE_N];                      // Dropout2d mask row of this thread's image (1 where no mask)
      if (p.mask && valid) {
        const float* mrow = p.mask + (int64_t)n * p.mask_pitch + k0;
#pragma unroll
        for (int j = 0; j < TILE_N; ++j) mk[j] = (k0 + j < p.K) ? __ldg(mrow + j) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < TILE_N; ++j) mk[j] = 1.f;
      }
      if (p.tma_store) {
        // ---- staged epilogue: registers -> swizzled shared-memory rows -> one TMA tensor store per output row ----
        constexpr uint32_t ROW_B = TILE_N * 2, BUF_B = 128 * ROW_B;
        const uint32_t sw = TILE_N == 64 ? (uint32_t)(m & 7) : (TILE_N == 32 ? (uint32_t)((m >> 1) & 3) : (uint32_t)((m >> 2) & 1));
        const uint32_t srow = smem_u32(stage) + (uint32_t)m * ROW_B;
        const bool issuer = et == 0;
        const int ox = cl.px + jt * p.XG * p.ostep, n0 = nt * p.NG;
        const uint32_t buf = (uint32_t)grp * BUF_B;
        for (int i = 0; i < cl.Pi; ++i, ++g) {
          if ((g & (WS_EPI_GROUPS - 1)) != grp) continue;
          const int acc = g & (p.n_acc - 1);
          if (issuer) bulk_wait_read<0>();          // this group's previous store has drained the staging buffer
          WS_TIMED_WAIT(w0, acc_full(acc), (uint32_t)(g >> p.acc_shift) & 1u);
          tc_fence_after();
          epi_bar(grp);                             // buffer free (and the statistics pass over it finished)
          const long long t_ld = dbg_on ? clock64() : 0;
#pragma unroll
          for (int c0 = 0; c0 < TILE_N; c0 += 32) {
            uint32_t va[16], vb[16];
            tmem_ld16(lane_addr + (uint32_t)(acc * TILE_N + c0), va);
            if (TILE_N > 16) tmem_ld16(lane_addr + (uint32_t)(acc * TILE_N + c0 + 16), vb);
            tmem_ld_wait();
            if (dbg_on && c0 + 32 >= TILE_N) w1 += clock64() - t_ld;
            if (c0 + 32 >= TILE_N) {        // accumulator fully read: hand it back before the math
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty(acc));
            }
            const int nv = p.K - (k0 + c0);
            uint4 o0, o1, o2, o3;
            epi16_pack(va, sbias + c0, *reinterpret_cast<const float(*)[16]>(&mk[c0]), p.act, p.slope,
                       nv < 0 ? 0 : (nv < 16 ? nv : 16), o0, o1);
            sts128(srow + buf + (((uint32_t)(c0 / 8) ^ sw) << 4), o0);
            sts128(srow + buf + (((uint32_t)(c0 / 8 + 1) ^ sw) << 4), o1);
            if (TILE_N > 16) {
              epi16_pack(vb, sbias + c0 + 16, *reinterpret_cast<const float(*)[16]>(&mk[TILE_N > 16 ? c0 + 16 : 0]), p.act,
                         p.slope, nv < 16 ? 0 : (nv < 32 ? nv - 16 : 16), o2, o3);
              sts128(srow + buf + (((uint32_t)(c0 / 8 + 2) ^ sw) << 4), o2);
              sts128(srow + buf + (((uint32_t)(c0 / 8 + 3) ^ sw) << 4), o3);
            }
          }
          fence_proxy_async();
          epi_bar(grp);
          if (issuer) {
            tma_store_4d(&map_o, smem_u32(stage) + buf, k0, n0, ox, cl.py + i * p.ostep);
            bulk_commit();
          }
          if (p.stats) {
            // column sums of the staged (bf16-rounded, as stored) tile, 16 bytes per load; rows beyond the batch /
            // row end are skipped
            const bool full = n0 + p.NG <= p.N && jt * p.XG + p.XG <= cl.Qj;
            const uint8_t* tile = stage + buf;
#pragma unroll
            for (int rr = 0; rr < CH; ++rr) {
              const int row = st_r0 + rr * (128 / CH);
              if (!full && !(n0 + (row & (p.NG - 1)) < p.N && jt * p.XG + (row >> p.ng_shift) < cl.Qj)) continue;
              const uint32_t rsw = TILE_N == 64 ? (uint32_t)(row & 7) : (TILE_N == 32 ? (uint32_t)((row >> 1) & 3) : (uint32_t)((row >> 2) & 1));
              const uint4 w4 = *reinterpret_cast<const uint4*>(tile + (uint32_t)row * ROW_B + (((uint32_t)st_ch ^ rsw) << 4));
              const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xFFFF0000u);
                st_sum[2 * j] += lo;
                st_sq[2 * j] = fmaf(lo, lo, st_sq[2 * j]);
                st_sum[2 * j + 1] += hi;
                st_sq[2 * j + 1] = fmaf(hi, hi, st_sq[2 * j + 1]);
              }
            }
          }
        }
        continue;
      }
      // destination of output row 0 of this thread's pixel column, channel k0
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.dst) +
                      ((((int64_t)(valid ? n : 0) * p.P + cl.py) * p.Q + cl.px + (int64_t)(valid ? jj : 0) * p.ostep) * p.out_pitch + k0) * esize;
      const int64_t row_step = (int64_t)p.ostep * p.Q * p.out_pitch * esize;
      for (int i = 0; i < cl.Pi; ++i, ++g, orow += row_step) {
        if ((g & (WS_EPI_GROUPS - 1)) != grp) continue;
        const int acc = g & (p.n_acc - 1);
        WS_TIMED_WAIT(w0, acc_full(acc), (uint32_t)(g >> p.acc_shift) & 1u);
        tc_fence_after();
        const long long t_ld = dbg_on ? clock64() : 0;
#pragma unroll
        for (int c0 = 0; c0 < TILE_N; c0 += 32) {
          uint32_t va[16], vb[16];
          tmem_ld16(lane_addr + (uint32_t)(acc * TILE_N + c0), va);
          if (TILE_N > 16) tmem_ld16(lane_addr + (uint32_t)(acc * TILE_N + c0 + 16), vb);
          tmem_ld_wait();
          if (dbg_on && c0 + 32 >= TILE_N) w1 += clock64() - t_ld;
          if (c0 + 32 >= TILE_N) {          // accumulator fully read: hand it back before the math and stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(acc));
          }
          if (valid) {
            const int nv = p.K - (k0 + c0);
            if (nv > 0) {
              const int nvl = nv < 16 ? nv : 16;
              int npad = (nvl + 7) & ~7;
              if (k0 + c0 + npad > p.out_pitch) npad = nvl;
              epi16(va, sbias + c0, *reinterpret_cast<const float(*)[16]>(&mk[c0]), p.act, p.slope, nvl, npad, p.out_f32,
                    orow + c0 * esize);
            }
            if (TILE_N > 16 && nv > 16) {
              const int nvl = nv < 32 ? nv - 16 : 16;
              int npad = (nvl + 7) & ~7;
              if (k0 + c0 + 16 + npad > p.out_pitch) npad = nvl;
              epi16(vb, sbias + c0 + 16, *reinterpret_cast<const float(*)[16]>(&mk[TILE_N > 16 ? c0 + 16 : 0]), p.act,
                    p.slope, nvl, npad, p.out_f32, orow + (c0 + 16) * esize);
            }
          }
        }
      }
    }
    if (p.tma_store && et == 0) bulk_wait_all();
    if (p.tma_store && p.stats) {
      // lanes that share a chunk (same et % CH) first combine through shuffles, then one atomic per channel and warp
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = CH; o < 32; o <<= 1) {
          st_sum[j] += __shfl_xor_sync(0xffffffffu, st_sum[j], o);
          st_sq[j] += __shfl_xor_sync(0xffffffffu, st_sq[j], o);
        }
      }
      if (lane < CH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + st_ch * 8 + j;
          if (k < p.K) {
            atomicAdd(p.stats + k, st_sum[j]);
            atomicAdd(p.stats + p.K + k, st_sq[j]);
          }
        }
      }
    }
    if (dbg_on && threadIdx.x == WS_EPI0) { p.dbg[5] = clock64() - t_begin; p.dbg[6] = w0; p.dbg[7] = w1; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_rt(tmem_base, p.tmem_cols);
}

inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

using icf::sm_count;

template <int TILE_N, int KD>
int launch_ws(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const WsParams& p, int grid, size_t smem,
              cudaStream_t st) {
  static icf::SmemGuard guard;
  if (int r = guard.ensure(reinterpret_cast<const void*>(conv_ws_kernel<TILE_N, KD>), smem, "row-streaming conv")) return r;
  conv_ws_kernel<TILE_N, KD><<<(unsigned)grid, WS_THREADS, smem, st>>>(ma, mb, mo, p);
  return icf::check_launch("conv_ws");
}

}  // namespace

namespace {
// returns 0 = launched (or, with `plan`, described), -1 = geometry outside this kernel's envelope (caller tries the
// next kernel), >0 = error
int ws_forward_impl(const icf_conv_args* a, cudaStream_t st, int32_t* plan, int32_t plan_words) {
  if (a->dtype != ICF_BF16 || a->accumulate) return -1;
  if ((a->in_pitch & 7) || (a->w_pitch & 7)) return -1;
  if (a->win > 1 && a->pad != 0) return -1;
  if ((reinterpret_cast<uintptr_t>(a->src) & 15) || (reinterpret_cast<uintptr_t>(a->w) & 15)) return -1;
  const bool gather = a->form == ICF_FORM_GATHER;
  const int taps_all = a->R * a->S;
  if (taps_all > WS_MAX_TAPS || a->stride > 2) return -1;

  WsParams q;
  memset(&q, 0, sizeof(q));
  q.N = a->N; q.P = a->P; q.Q = a->Q; q.K = a->K; q.out_pitch = a->out_pitch;
  q.sstep = gather ? a->stride : 1;
  q.ostep = gather ? 1 : a->stride;
  q.n_classes = q.ostep * q.ostep;
  q.kchunks = icf::cdiv(a->C, 64);
  q.kdepth_last = icf::cdiv(a->C - 64 * (q.kchunks - 1), 16);
  q.rowb = a->C <= 32 ? 64 : 128;
  q.w_pitch = a->w_pitch;
  const int tile_n = a->K <= 16 ? 16 : (a->K <= 32 ? 32 : 64);
  q.tiles_k = icf::cdiv(a->K, tile_n);
  q.n_acc = 512 / tile_n > WS_MAX_ACC ? WS_MAX_ACC : 512 / tile_n;
  q.acc_shift = ilog2(q.n_acc);
  q.tmem_cols = (uint32_t)(q.n_acc * tile_n);

  // ---- taps per class; x-parity sub-rows ----
  struct RawTap { int dy, dx, widx; };
  RawTap raw[WS_MAX_CLASSES][WS_MAX_TAPS];
  int bmin[2] = {1 << 30, 1 << 30}, bmax[2] = {-(1 << 30), -(1 << 30)};
  bool sub_used[2] = {false, false};
  int max_ntaps = 0, max_q = 0;
  for (int c = 0; c < q.n_classes; ++c) {
    WsClass& cl = q.cls[c];
    cl.py = c / q.ostep; cl.px = c % q.ostep;
    cl.Pi = (a->P - cl.py + q.ostep - 1) / q.ostep;
    cl.Qj = (a->Q - cl.px + q.ostep - 1) / q.ostep;
    if (cl.Pi <= 0 || cl.Qj <= 0) return -1;
    int n = 0;
    for (int r = 0; r < a->R; ++r)
      for (int s = 0; s < a->S; ++s) {
        int dy, dx;
        if (gather) {
          dy = r - a->pad; dx = s - a->pad;
        } else {
          const int uy = cl.py + a->pad - r, ux = cl.px + a->pad - s;
          if (((uy % a->stride) + a->stride) % a->stride != 0 || ((ux % a->stride) + a->stride) % a->stride != 0) continue;
          dy = uy / a->stride; dx = ux / a->stride;
        }
        raw[c][n++] = {dy, dx, r * a->S + s};
        const int par = ((dx % q.sstep) + q.sstep) % q.sstep, b = floor_div(dx, q.sstep);
        sub_used[par] = true;
        bmin[par] = b < bmin[par] ? b : bmin[par];
        bmax[par] = b > bmax[par] ? b : bmax[par];
      }
    if (n == 0) return -1;                    // class without taps (bias-only outputs): other kernels handle it
    cl.ntaps = n;
    max_ntaps = n > max_ntaps ? n : max_ntaps;
    max_q = cl.Qj > max_q ? cl.Qj : max_q;
    int dymin = 1 << 30, dymax = -(1 << 30);
    for (int t = 0; t < n; ++t) { dymin = raw[c][t].dy < dymin ? raw[c][t].dy : dymin; dymax = raw[c][t].dy > dymax ? raw[c][t].dy : dymax; }
    cl.dymax = dymax;
    if ((dymax - dymin) / q.sstep + 1 > q.n_acc - 1) return -1;     // rows in flight must leave one accumulator to drain
    cl.ylo = dymin > 0 ? dymin : 0;
    const int yh = (cl.Pi - 1) * q.sstep + dymax;
    cl.yhi = yh < a->H - 1 ? yh : a->H - 1;
    if (cl.ylo > cl.yhi) return -1;
    for (int i = 0; i < cl.Pi; ++i) {         // every output row needs at least one in-bounds source row
      bool any = false;
      for (int t = 0; t < n && !any; ++t) { const int y = i * q.sstep + raw[c][t].dy; any = y >= 0 && y < a->H; }
      if (!any) return -1;
    }
  }
  int sub_index[2] = {-1, -1};
  q.nsub = 0;
  int brange = 0;
  for (int par = 0; par < q.sstep; ++par)
    if (sub_used[par]) {
      sub_index[par] = q.nsub++;
      brange = bmax[par] - bmin[par] > brange ? bmax[par] - bmin[par] : brange;
    }
  // ---- column shape: XG output x positions times NG images ----
  double best = 1e30;
  for (int xg = 1; xg <= 16; xg *= 2) {
    // useful MMA rows of a column: x positions inside the row times images inside the batch (a column is XG x positions of
    // NG = 128/XG images — with 32 images per step a 1 x 128 column would be three quarters padding)
    const int ng = 128 / xg;
    const double eff = (double)max_q / (icf::cdiv(max_q, xg) * xg) * ((double)a->N / ((double)icf::cdiv(a->N, ng) * ng));
    // wasted MMA rows (1/eff) against x-halo re-fetch; stride-2 gathers load two parity sub-rows per source row and
    // measured better with wide columns (G.layers.6 dgrad: XG 16 beats XG 2 by 25 %)
    const double cost = (1.0 / eff) * (1.0 + (q.sstep == 2 ? 0.5 : 0.25) * brange / xg);
    if (cost < best - 1e-9 || (cost < best + 1e-9 && xg > q.XG)) { best = cost; q.XG = xg; }
  }
  {
    static const int xg_env = []() { const char* e = getenv("ICF_WS_XG"); return e ? atoi(e) : 0; }();   // tuning aid
    if (xg_env == 1 || xg_env == 2 || xg_env == 4 || xg_env == 8 || xg_env == 16) q.XG = xg_env;
  }
  q.NG = 128 / q.XG;
  q.ng_shift = ilog2(q.NG);
  const int hx = q.XG + brange;
  if (hx * q.sstep > 256) return -1;
  q.kc_bytes = (uint32_t)(hx * q.NG) * (uint32_t)q.rowb;
  q.sub_bytes = q.kc_bytes * (uint32_t)q.kchunks;
  q.slot_bytes = q.sub_bytes * (uint32_t)q.nsub;
  q.slab_bytes = ((uint32_t)(max_ntaps * q.kchunks * tile_n) * (uint32_t)q.rowb + 1023u) & ~1023u;
  const int64_t stage_bytes = 2 * 128 * tile_n * 2;
  if ((int64_t)q.slab_bytes + stage_bytes + 2 * (int64_t)q.slot_bytes > WS_SMEM_BUDGET) return -1;
  q.tiles_n = icf::cdiv(a->N, q.NG);

  // ---- tap tables, CTA partition over (class, k-tile) ----
  double work[WS_MAX_CLASSES], total = 0.0;
  for (int c = 0; c < q.n_classes; ++c) {
    WsClass& cl = q.cls[c];
    cl.tiles_x = icf::cdiv(cl.Qj, q.XG);
    for (int t = 0; t < cl.ntaps; ++t) {
      const int dx = raw[c][t].dx;
      const int par = ((dx % q.sstep) + q.sstep) % q.sstep, b = floor_div(dx, q.sstep);
      cl.taps[t].dy = (int16_t)raw[c][t].dy;
      cl.taps[t].widx = (int16_t)raw[c][t].widx;
      cl.taps[t].a_off16 = ((uint32_t)sub_index[par] * q.sub_bytes + (uint32_t)((b - bmin[par]) * q.NG) * (uint32_t)q.rowb) >> 4;
      if (t == 0 || raw[c][t].dy != raw[c][t - 1].dy) {
        if (cl.ngroups == WS_MAX_GROUPS) return -1;
        cl.grp[cl.ngroups].dy = (int16_t)raw[c][t].dy;
        cl.grp[cl.ngroups].first = (int8_t)t;
        cl.grp[cl.ngroups].count = 0;
        ++cl.ngroups;
      }
      ++cl.grp[cl.ngroups - 1].count;
    }
    for (int g1 = 0; g1 < cl.ngroups; ++g1) {        // previous contributor to the same output row: next smaller dy
      int best = -(1 << 30);
      for (int g2 = 0; g2 < cl.ngroups; ++g2) {
        const int d2 = cl.grp[g2].dy, d1 = cl.grp[g1].dy;
        if (d2 < d1 && d2 > best) best = d2;
      }
      cl.grp[g1].dprev = best == -(1 << 30) ? (int16_t)32767 : (int16_t)(cl.grp[g1].dy - best);
    }
    for (int par = 0; par < q.sstep; ++par)
      if (sub_used[par]) cl.x0[sub_index[par]] = bmin[par] * q.sstep + par;
    work[c] = (double)cl.tiles_x * cl.Pi * cl.ntaps;
    total += work[c];
  }
  // ---- issuer schedules (see WsParams::prog) ----
  int64_t prog_bytes = 0;                            // shared memory: tap table + the longest class's issuer schedules
  {
    int used = 0;
    for (int c = 0; c < q.n_classes; ++c) {
      WsClass& cl = q.cls[c];
      if (cl.Pi > 256) return -1;
      for (int wi = 0; wi < WS_ISSUERS; ++wi) {
        cl.prog_off[wi] = used;
        for (int y = cl.ylo; y <= cl.yhi; ++y) {
          const int row_begin = used;
          for (int gi = 0; gi < cl.ngroups; ++gi) {
            const int num = y - cl.grp[gi].dy;
            const int i = q.sstep == 1 ? num : (num >> 1);
            if (num < 0 || i >= cl.Pi || i % WS_ISSUERS != wi || (q.sstep == 2 && (num & 1))) continue;
            const bool opens = y - cl.grp[gi].dprev < cl.ylo;     // first in-bounds contributor of output row i
            if (used >= WS_PROG_WORDS - 2) return -1;
            q.prog[used++] = (uint32_t)i | ((uint32_t)cl.grp[gi].first << 8) | ((uint32_t)cl.grp[gi].count << 13) | (opens ? 1u << 20 : 0u);
          }
          for (int i = wi; i < cl.Pi; i += WS_ISSUERS) {           // output rows whose last source row is y
            int yc = i * q.sstep + cl.dymax;
            yc = yc < cl.ylo ? cl.ylo : (yc > cl.yhi ? cl.yhi : yc);
            if (yc != y) continue;
            if (used >= WS_PROG_WORDS - 2) return -1;
            q.prog[used++] = (uint32_t)i | (1u << 18);
          }
          if (used == row_begin) {
            if (used >= WS_PROG_WORDS - 2) return -1;
            q.prog[used++] = 2u << 18;
          }
          q.prog[used - 1] |= 1u << 21;
        }
        q.prog[used++] = 2u << 18;                                // spare word: the issue loop reads one action ahead
      }
      cl.prog_off[WS_ISSUERS] = used;
      const int64_t need = 256 + 4 * (int64_t)(used - cl.prog_off[0]);
      prog_bytes = need > prog_bytes ? need : prog_bytes;
    }
    prog_bytes = (prog_bytes + 255) & ~(int64_t)255;
  }
  if ((int64_t)q.slab_bytes + stage_bytes + prog_bytes + 2 * (int64_t)q.slot_bytes > WS_SMEM_BUDGET) return -1;
  q.n_slots = (int)((WS_SMEM_BUDGET - (int64_t)q.slab_bytes - stage_bytes - prog_bytes) / q.slot_bytes);
  if (q.n_slots > WS_MAX_SLOTS) q.n_slots = WS_MAX_SLOTS;
  const int budget = sm_count();
  int grid = 0;
  for (int c = 0; c < q.n_classes; ++c) {
    WsClass& cl = q.cls[c];
    const int ncols = q.tiles_n * cl.tiles_x;
    int per_k = (int)((double)budget * work[c] / total / q.tiles_k);
    if (per_k < 1) per_k = 1;
    if (per_k > ncols) per_k = ncols;
    cl.cta_begin = grid;
    cl.cta_count = per_k * q.tiles_k;
    grid += cl.cta_count;
  }
  q.act = a->act; q.slope = a->slope; q.out_f32 = a->out_f32; q.mask_pitch = a->mask_pitch;
  q.bias = a->bias; q.mask = a->out_mask; q.dst = a->dst;
  static unsigned long long* dbg_buf = []() -> unsigned long long* {
    const char* e = getenv("ICF_WS_DEBUG");
    void* d = nullptr;
    if (e && e[0] && e[0] != '0' && cudaMalloc(&d, 128) == cudaSuccess) cudaMemset(d, 0, 128);
    return reinterpret_cast<unsigned long long*>(d);
  }();
  q.dbg = dbg_buf;

  if (plan) {
    // ---- icf_ws_plan: describe instead of launching (layout documented in include/icf.h) ----
    const int need = 16 + WS_MAX_CLASSES * 48 + WS_PROG_WORDS;
    if (plan_words < need) { icf::set_error("icf_ws_plan: out needs %d words", need); return 1; }
    memset(plan, 0, sizeof(int32_t) * (size_t)need);
    plan[0] = q.n_classes; plan[1] = q.XG; plan[2] = q.NG; plan[3] = q.n_slots; plan[4] = q.n_acc; plan[5] = tile_n;
    plan[6] = q.sstep; plan[7] = q.ostep; plan[8] = WS_ISSUERS; plan[9] = q.cls[q.n_classes - 1].prog_off[WS_ISSUERS];
    plan[10] = grid; plan[11] = q.tiles_k; plan[12] = q.tiles_n; plan[13] = (int32_t)prog_bytes;
    for (int c = 0; c < q.n_classes; ++c) {
      const WsClass& cl = q.cls[c];
      int32_t* o = plan + 16 + 48 * c;
      o[0] = cl.Pi; o[1] = cl.Qj; o[2] = cl.py; o[3] = cl.px; o[4] = cl.ylo; o[5] = cl.yhi; o[6] = cl.dymax; o[7] = cl.ngroups;
      o[8] = cl.ntaps; o[9] = cl.tiles_x; o[10] = cl.cta_begin; o[11] = cl.cta_count;
      for (int wi = 0; wi <= WS_ISSUERS; ++wi) o[12 + wi] = cl.prog_off[wi];
      for (int g = 0; g < cl.ngroups; ++g) {
        o[16 + 4 * g] = cl.grp[g].dy; o[17 + 4 * g] = cl.grp[g].dprev; o[18 + 4 * g] = cl.grp[g].first; o[19 + 4 * g] = cl.grp[g].count;
      }
    }
    for (int w = 0; w < plan[9]; ++w) plan[16 + WS_MAX_CLASSES * 48 + w] = (int32_t)q.prog[w];
    return 0;
  }

  CUtensorMap ma, mb, mo;
  memset(&mo, 0, sizeof(mo));
  // TMA-store epilogue: bf16 destination whose pixels are 16-byte aligned rows (pitch % 8 == 0)
  q.tma_store = (!a->out_f32 && (a->out_pitch & 7) == 0 && (reinterpret_cast<uintptr_t>(a->dst) & 15) == 0) ? 1 : 0;
  {
    static const bool off = []() { const char* e = getenv("ICF_WS_NO_TMA_STORE"); return e && e[0] && e[0] != '0'; }();
    if (off) q.tma_store = 0;
  }
  q.stats = (q.tma_store && a->stats) ? a->stats : nullptr;
  if (q.tma_store) {
    int kp = (a->K + 7) & ~7;                       // ragged K: the pitch padding is written as zeros
    if (kp > a->out_pitch) kp = a->K;
    cuuint64_t dims[4] = {(cuuint64_t)kp, (cuuint64_t)a->N, (cuuint64_t)a->Q, (cuuint64_t)a->P};
    cuuint64_t str[3] = {(cuuint64_t)a->P * a->Q * a->out_pitch * 2, (cuuint64_t)a->out_pitch * 2,
                         (cuuint64_t)a->Q * a->out_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)tile_n, (cuuint32_t)q.NG, (cuuint32_t)(q.XG * q.ostep), 1};
    cuuint32_t est[4] = {1, 1, (cuuint32_t)q.ostep, 1};
    const CUtensorMapSwizzle swz = tile_n == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (tile_n == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (int r = encode_map_swz(&mo, a->dst, 4, dims, str, box, est, swz)) return r;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->N, (cuuint64_t)a->W, (cuuint64_t)a->H};
    cuuint64_t str[3] = {(cuuint64_t)a->H * a->W * a->in_pitch * 2, (cuuint64_t)a->in_pitch * 2,
                         (cuuint64_t)a->W * a->in_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)(q.rowb / 2), (cuuint32_t)q.NG, (cuuint32_t)(hx * q.sstep), 1};
    cuuint32_t est[4] = {1, 1, (cuuint32_t)q.sstep, 1};
    if (int r = encode_map_swz(&ma, a->src, 4, dims, str, box, est, q.rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps_all * a->w_pitch, (cuuint64_t)a->w_rows};
    cuuint64_t str[1] = {(cuuint64_t)taps_all * a->w_pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)(q.rowb / 2), (cuuint32_t)tile_n};
    cuuint32_t est[2] = {1, 1};
    if (int r = encode_map_swz(&mb, a->w, 2, dims, str, box, est, q.rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  }
  // shared memory: always more than half an SM's worth so that exactly one CTA (and one TMEM allocation) is resident
  size_t smem = (size_t)q.slab_bytes + (size_t)q.n_slots * q.slot_bytes + (size_t)stage_bytes + 1024 + 1024 + (size_t)prog_bytes;

  if (smem < 120 * 1024) smem = 120 * 1024;
  int r;
  const int kd = q.kchunks <= 2 ? 4 * (q.kchunks - 1) + q.kdepth_last : 0;
#define ICF_WS_CASE(TN)                                                        \
  switch (kd) {                                                                \
    case 1: r = launch_ws<TN, 1>(ma, mb, mo, q, grid, smem, st); break;        \
    case 2: r = launch_ws<TN, 2>(ma, mb, mo, q, grid, smem, st); break;        \
    case 3: r = launch_ws<TN, 3>(ma, mb, mo, q, grid, smem, st); break;        \
    case 4: r = launch_ws<TN, 4>(ma, mb, mo, q, grid, smem, st); break;        \
    case 5: r = launch_ws<TN, 5>(ma, mb, mo, q, grid, smem, st); break;        \
    case 6: r = launch_ws<TN, 6>(ma, mb, mo, q, grid, smem, st); break;        \
    case 7: r = launch_ws<TN, 7>(ma, mb, mo, q, grid, smem, st); break;        \
    case 8: r = launch_ws<TN, 8>(ma, mb, mo, q, grid, smem, st); break;        \
    default: r = launch_ws<TN, 0>(ma, mb, mo, q, grid, smem, st); break;       \
  }
  switch (tile_n) {
    case 16: ICF_WS_CASE(16) break;
    case 32: ICF_WS_CASE(32) break;
    default: ICF_WS_CASE(64) break;
  }
#undef ICF_WS_CASE
  if (r) return r;
  const bool stats_fused = q.stats != nullptr;
  if (q.dbg) {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, q.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr,
            "[icf ws] K=%d C=%d R=%d stride=%d form=%d tile_n=%d XG=%d NG=%d slots=%d slot=%uB slab=%uB grid=%d | CTA0 cycles: "
            "producer %llu (wait empty %llu) | mma %llu (wait full %llu, wait acc %llu, chains %llu in %llu, commits %llu) | epilogue %llu (wait acc_full %llu, tmem ld %llu)\n",
            a->K, a->C, a->R, a->stride, a->form, tile_n, q.XG, q.NG, q.n_slots, q.slot_bytes, q.slab_bytes, grid, h[0], h[1],
            h[2], h[3], h[4], h[10], h[8], h[9], h[5], h[6], h[7]);
  }
  if (a->stats && !stats_fused) {
    if (a->out_f32) { icf::set_error("row-streaming conv: BatchNorm statistics need a bf16 destination"); return 1; }
    return icf_launch_col_stats(a->dst, ICF_BF16, a->out_pitch, (int64_t)a->N * a->P * a->Q, a->K, a->stats, st);
  }
  return 0;
}
}  // namespace

int icf_ws_conv_forward(const icf_conv_args* a, cudaStream_t st) { return ws_forward_impl(a, st, nullptr, 0); }

int icf_ws_plan(const icf_conv_args* a, int32_t* out, int32_t out_words) {
  ICF_REQUIRE(a && out && out_words > 0, "icf_ws_plan: bad arguments");
  return ws_forward_impl(a, nullptr, out, out_words);
}

