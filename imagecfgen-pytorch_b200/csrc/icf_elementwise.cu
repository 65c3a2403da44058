// HBM-bound kernels of the BiGAN hot path: attribute/latent feature assembly, BatchNorm pieces, the fused
// activation/dropout/BN backward, BCE-with-logits, Adam, weight (un)packing.
#include "icf_common.cuh"

namespace {

// sample index of a pixel: 32-bit unsigned division whenever the pixel index fits (a 64-bit division costs ~10x
// more issue slots than the rest of a 16-byte-per-thread streaming iteration)
__device__ __forceinline__ int64_t sample_of(int64_t pix, int pps) {
  return pix <= 0xffffffffLL ? (int64_t)((uint32_t)pix / (uint32_t)pps) : pix / pps;
}

constexpr int EW_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// argmax over rows (first maximum wins, torch.argmax semantics; all-zero row -> 0)
// ------------------------------------------------------------------------------------------------
__global__ void argmax_rows_kernel(const void* x, int dt, int n, int k, int32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int best = 0;
  if (dt == 2) {
    const int32_t* p = reinterpret_cast<const int32_t*>(x) + (int64_t)i * k;
    int32_t bv = p[0];
    for (int j = 1; j < k; ++j) if (p[j] > bv) { bv = p[j]; best = j; }
  } else if (dt == 3) {
    const int64_t* p = reinterpret_cast<const int64_t*>(x) + (int64_t)i * k;
    int64_t bv = p[0];
    for (int j = 1; j < k; ++j) if (p[j] > bv) { bv = p[j]; best = j; }
  } else {
    float bv = icf::ld_any(x, dt, (int64_t)i * k);
    for (int j = 1; j < k; ++j) {
      float v = icf::ld_any(x, dt, (int64_t)i * k + j);
      if (v > bv) { bv = v; best = j; }
    }
  }
  out[i] = best;
}

// ------------------------------------------------------------------------------------------------
// image feature stack: [x | tanh(upsample16(emb[idx])) ... | const planes ...] * mask, NHWC pitch 8
// A block works on one sample at a time: the sample's 16x16 attribute planes (tanh applied, Dropout2d mask folded in) and
// its constant planes are staged in shared memory once (256 tanh per plane instead of one per pixel), then the threads
// stream the pixels: one 4-byte read of the image and one 16-byte store of the (<= 8 channel) vector per pixel.
// grid = (samples, pixel slabs): big images (512x512) are split so that a small batch still fills the machine.
// ------------------------------------------------------------------------------------------------
constexpr int IMGFEAT_MAX_W = 1024;
__global__ void __launch_bounds__(EW_THREADS) imgfeat_fwd_kernel(const icf_imgfeat_args a) {
  __shared__ float plane[ICF_MAX_PLANES - 1][256];
  __shared__ float cst[8], mk0_s;
  __shared__ uint8_t cxs[IMGFEAT_MAX_W], cys[IMGFEAT_MAX_W];           // cell column / row of every padded x / y (255 = border)
  const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
  const int rows_per = (Hp + (int)gridDim.y - 1) / (int)gridDim.y;
  const int y_lo = (int)blockIdx.y * rows_per, y_hi = min(Hp, y_lo + rows_per);
  const bool vec = a.dtype == ICF_BF16 && a.feat_pitch == 8;           // one 16-byte store per pixel
  const bool xf32 = a.x_dtype == ICF_F32;
  const int n_emb = a.n_emb < ICF_MAX_PLANES - 1 ? a.n_emb : ICF_MAX_PLANES - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = (int)blockDim.x, NW = NT >> 5;                        // 64 threads for small images, 256 for large ones
  for (int xp = threadIdx.x; xp < Wp; xp += NT) {
    const int x = xp - a.pad;
    cxs[xp] = (uint8_t)((x >= 0 && x < a.W) ? min((x * 16) / a.W, 15) : 255);   // nearest: floor(dst*16/size); 255 = border
  }
  for (int yp = threadIdx.x; yp < Hp; yp += NT) {
    const int y = yp - a.pad;
    cys[yp] = (uint8_t)((y >= 0 && y < a.H) ? min((y * 16) / a.H, 15) : 255);
  }
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    const float* mk = a.mask ? a.mask + (int64_t)n * a.mask_pitch : nullptr;
    for (int t = threadIdx.x; t < n_emb * 256; t += NT) {
      const int e = t >> 8, cell = t & 255;
      const float v = tanhf(a.emb_table[e][(int64_t)a.emb_index[e][n] * 256 + cell]);
      plane[e][cell] = (mk && 1 + e < 8) ? v * mk[1 + e] : v;
    }
    if (threadIdx.x < 8) {                                             // cst[channel]: constant planes, 0 elsewhere
      const int e = (int)threadIdx.x - 1 - n_emb;
      float v = 0.f;
      if (e >= 0 && e < a.n_cont) { v = a.cont[e][n]; if (mk) v *= mk[threadIdx.x]; }
      cst[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) mk0_s = mk ? mk[0] : 1.f;
    __syncthreads();
    // One warp per (padded) image row, 32 pixels per pass.  The first version of this loop was INSTRUCTION bound (ncu: 52 % SM
    // busy at 2 % of DRAM, ~220 instructions per pixel): two integer divisions per pixel and pass, 64-bit index arithmetic and a
    // run-time dtype switch per access.  Now the nearest-neighbour cell of a row / column comes from the two tables above, the
    // row base pointers are computed once per row, and the loads of UN rows are issued before their math and stores.
    const float mk0 = mk0_s;
    constexpr int UN = 4;
    const int64_t x_img = (int64_t)n * a.H * a.W;
    const int64_t o_img = (int64_t)n * Hp * Wp;
    for (int r0 = y_lo + warp; r0 < y_hi; r0 += NW * UN) {
      for (int xp = lane; xp < Wp; xp += 32) {
        const int cx = cxs[xp];
        float xv[UN];
        int cellv[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int yp = r0 + u * NW;
          xv[u] = 0.f;
          cellv[u] = -1;
          if (yp < y_hi) {
            const int cy = cys[yp];
            if (cy != 255 && cx != 255) {
              cellv[u] = cy * 16 + cx;
              const int64_t xi = (x_img + (int64_t)(yp - a.pad) * a.W + (xp - a.pad)) * a.x_pitch;
              xv[u] = xf32 ? reinterpret_cast<const float*>(a.x)[xi]
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.x)[xi]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int yp = r0 + u * NW;
          if (yp >= y_hi) continue;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = 0.f;
          if (cellv[u] >= 0) {
            f[0] = xv[u] * mk0;
#pragma unroll
            for (int j = 1; j < 8; ++j) {                               // embedded planes first, then the constant planes
              const int e = j - 1;
              f[j] = e < n_emb ? plane[e][cellv[u]] : cst[j];
            }
          }
          const int64_t o = o_img + (int64_t)yp * Wp + xp;
          if (vec) {
            uint4 w;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
            w.x = *reinterpret_cast<uint32_t*>(&p0); w.y = *reinterpret_cast<uint32_t*>(&p1);
            w.z = *reinterpret_cast<uint32_t*>(&p2); w.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.feat) + o * 8) = w;
          } else {
            for (int ch = 0; ch < a.feat_pitch; ++ch) icf::st_any(a.feat, a.dtype, o * a.feat_pitch + ch, ch < 8 ? f[ch] : 0.f);
          }
        }
      }
    }
    __syncthreads();
  }
}

// gradient into the embedding tables: one warp per (sample, plane, cell): reduce the cell's pixel block
__global__ void imgfeat_bwd_kernel(const icf_imgfeat_args a) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)a.N * a.n_emb * 256;
  if (warp >= total) return;
  const int cell = (int)(warp & 255);
  const int e = (int)((warp >> 8) % a.n_emb);
  const int n = (int)((warp >> 8) / a.n_emb);
  const int cy = cell >> 4, cx = cell & 15;
  // destination rows y with floor(y*16/H) == cy  <=>  y in [ceil(cy*H/16), ceil((cy+1)*H/16))
  const int y0 = (cy * a.H + 15) / 16, y1 = ((cy + 1) * a.H + 15) / 16;
  const int x0 = (cx * a.W + 15) / 16, x1 = ((cx + 1) * a.W + 15) / 16;
  const int bw = x1 - x0, cnt = (y1 - y0) * bw;
  const int ch = 1 + e;
  float sum = 0.f;
  for (int t = lane; t < cnt; t += 32) {
    const int y = y0 + t / bw, x = x0 + t % bw;
    sum += icf::ld_any(a.dfeat, a.dtype, (((int64_t)n * a.H + y) * a.W + x) * a.feat_pitch + ch);
  }
  sum = icf::warp_sum(sum);
  if (lane == 0) {
    const int idx = a.emb_index[e][n];
    const float t = tanhf(a.emb_table[e][(int64_t)idx * 256 + cell]);
    float g = sum * (1.f - t * t);
    if (a.mask) g *= a.mask[(int64_t)n * a.mask_pitch + ch];
    if (g != 0.f) atomicAdd(a.demb_table[e] + (int64_t)idx * 256 + cell, g);
  }
}

// same gradient, one THREAD per (sample, plane, cell): for small images (28x28: 1-4 pixels per cell) a warp per
// cell would leave 28+ lanes idle
__global__ void imgfeat_bwd_small_kernel(const icf_imgfeat_args a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.N * a.n_emb * 256;
  if (t >= total) return;
  const int cell = (int)(t & 255);
  const int e = (int)((t >> 8) % a.n_emb);
  const int n = (int)((t >> 8) / a.n_emb);
  const int cy = cell >> 4, cx = cell & 15;
  const int y0 = (cy * a.H + 15) / 16, y1 = ((cy + 1) * a.H + 15) / 16;
  const int x0 = (cx * a.W + 15) / 16, x1 = ((cx + 1) * a.W + 15) / 16;
  const int ch = 1 + e;
  float sum = 0.f;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x)
      sum += icf::ld_any(a.dfeat, a.dtype, (((int64_t)n * a.H + y) * a.W + x) * a.feat_pitch + ch);
  const int idx = a.emb_index[e][n];
  const float tv = tanhf(a.emb_table[e][(int64_t)idx * 256 + cell]);
  float g = sum * (1.f - tv * tv);
  if (a.mask) g *= a.mask[(int64_t)n * a.mask_pitch + ch];
  if (g != 0.f) atomicAdd(a.demb_table[e] + (int64_t)idx * 256 + cell, g);
}

// ------------------------------------------------------------------------------------------------
// latent feature vector: [z | onehot_i @ table_i ... | cont ... | 0 pad]
// ------------------------------------------------------------------------------------------------
__global__ void latfeat_fwd_kernel(const icf_latfeat_args a) {
  const int64_t total = (int64_t)a.N * a.feat_pitch;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i / a.feat_pitch);
  const int j = (int)(i - (int64_t)n * a.feat_pitch);
  float v = 0.f;
  if (j < a.latent) {
    v = icf::ld_any(a.z, a.z_dtype, (int64_t)n * a.z_pitch + j);
  } else {
    const int jj = j - a.latent;
    const int e = jj >> 8;
    if (e < a.n_emb) {
      const int col = jj & 255;
      const int K = a.emb_k[e];
      const float* oh = a.onehot[e] + (int64_t)n * K;
      const float* tb = a.emb_table[e];
      for (int k = 0; k < K; ++k) v = fmaf(oh[k], tb[k * 256 + col], v);
    } else {
      const int ci = jj - a.n_emb * 256;
      if (ci < a.n_cont) v = a.cont[ci][n];
    }
  }
  icf::st_any(a.feat, a.dtype, i, v);
}

// dz, d_onehot, d_cont: one thread per element of the un-padded feature row
__global__ void latfeat_bwd_rows_kernel(const icf_latfeat_args a) {
  const int width = a.latent + a.n_cont;
  int tot_k = 0;
  for (int e = 0; e < a.n_emb; ++e) tot_k += a.emb_k[e];
  const int per_row = width + tot_k;
  const int64_t total = (int64_t)a.N * per_row;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i / per_row);
  int j = (int)(i - (int64_t)n * per_row);
  const int64_t row = (int64_t)n * a.feat_pitch;
  if (j < a.latent) {
    if (a.dz) a.dz[(int64_t)n * a.latent + j] = icf::ld_any(a.dfeat, a.dtype, row + j);
    return;
  }
  j -= a.latent;
  if (j < a.n_cont) {
    if (a.dcont[j]) a.dcont[j][n] = icf::ld_any(a.dfeat, a.dtype, row + a.latent + a.n_emb * 256 + j);
    return;
  }
  j -= a.n_cont;
  for (int e = 0; e < a.n_emb; ++e) {
    if (j < a.emb_k[e]) {
      if (a.donehot[e]) {
        const float* tb = a.emb_table[e] + (int64_t)j * 256;
        float s = 0.f;
        for (int col = 0; col < 256; ++col)
          s = fmaf(icf::ld_any(a.dfeat, a.dtype, row + a.latent + e * 256 + col), tb[col], s);
        a.donehot[e][(int64_t)n * a.emb_k[e] + j] = s;
      }
      return;
    }
    j -= a.emb_k[e];
  }
}

// d_table[e][k][col] += sum_n onehot[n][k] * dfeat[n][latent + e*256 + col]; block per (e,k), thread per col
// the sample range is split over blockIdx.y (a serial loop over the whole batch is latency-bound)
__global__ void latfeat_bwd_table_kernel(const icf_latfeat_args a, int e) {
  const int k = blockIdx.x;
  const int col = threadIdx.x;   // 256 threads
  const int K = a.emb_k[e];
  const int per = (a.N + gridDim.y - 1) / gridDim.y;
  const int n_lo = blockIdx.y * per, n_hi = min(a.N, n_lo + per);
  float s = 0.f;
  for (int n = n_lo; n < n_hi; ++n) {
    const float oh = a.onehot[e][(int64_t)n * K + k];
    if (oh != 0.f)
      s = fmaf(oh, icf::ld_any(a.dfeat, a.dtype, (int64_t)n * a.feat_pitch + a.latent + e * 256 + col), s);
  }
  if (s != 0.f) atomicAdd(a.demb_table[e] + (int64_t)k * 256 + col, s);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalize: stats -> scale/shift (+ running stats), then clear stats for the next forward
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(float* stats, int C, double count, const float* gamma, const float* beta,
                                   float eps, float momentum, float* rmean, float* rvar, int64_t* nbt,
                                   float* scale, float* shift, float* save_mean, float* save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = (double)stats[c] / count;
  double var = (double)stats[C + c] / count - mean * mean;   // biased variance normalises
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - (float)mean * g * invstd;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
  if (rvar) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
  }
  stats[c] = 0.f;
  stats[C + c] = 0.f;
}

// u = mask[n,c] * (scale[c]*y + shift[c])   (any of scale/shift/mask may be NULL), with dtype conversion
__global__ void scale_shift_mask_kernel(const void* y, int ydt, int ypitch, void* u, int udt, int upitch,
                                        int64_t pixels, int pps, int C, const float* scale,
                                        const float* shift, const float* mask, int mpitch) {
  const int64_t total = pixels * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C;
    const int c = (int)(i - pix * C);
    float v = icf::ld_any(y, ydt, pix * ypitch + c);
    if (scale) v = fmaf(v, scale[c], shift ? shift[c] : 0.f);
    if (mask) v *= mask[sample_of(pix, pps) * mpitch + c];
    icf::st_any(u, udt, pix * upitch + c, v);
  }
}

// per-channel sums of du and du*xhat over all pixels; grid (channel groups of 32, pixel slabs)
__global__ void bn_bwd_reduce_kernel(const void* dU, int ddt, int dpitch, const void* y, int ydt, int ypitch,
                                     int64_t pixels, int pps, int C, const float* mask, int mpitch,
                                     const float* mean, const float* invstd, float* sums) {
  __shared__ float red[2][8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;     // 32 channels x 8 pixel lanes
  const int c = blockIdx.x * 32 + cx;
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c];
    for (int64_t pix = (int64_t)blockIdx.y * 8 + py; pix < pixels; pix += (int64_t)gridDim.y * 8) {
      float g = icf::ld_any(dU, ddt, pix * dpitch + c);
      if (mask) g *= mask[sample_of(pix, pps) * mpitch + c];
      const float xh = (icf::ld_any(y, ydt, pix * ypitch + c) - mu) * is;
      s0 += g;
      s1 = fmaf(g, xh, s1);
    }
  }
  red[0][py][cx] = s0;
  red[1][py][cx] = s1;
  __syncthreads();
  if (py == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t0 += red[0][k][cx]; t1 += red[1][k][cx]; }
    atomicAdd(sums + c, t0);
    atomicAdd(sums + C + c, t1);
  }
}

// few-channel variant (C <= 4, no BatchNorm): one thread per pixel, block-reduced bias gradient — the 32-channel-lane
// kernel below would run one lane in 32 for the single-channel image gradient of the last generator layer
__global__ void act_backward_fewc_kernel(const icf_actbwd_args a) {
  __shared__ float red[32];
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < a.pixels; pix += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = sample_of(pix, a.pixels_per_sample);
    float gv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < a.C; ++c) {
      float g = icf::ld_any(a.dOut, a.d_dtype, pix * a.d_pitch + c);
      const float yv = icf::ld_any(a.y, a.y_dtype, pix * a.y_pitch + c);
      if (a.out_mask) g *= a.out_mask[n * a.mask_pitch + c];
      g *= icf::act_grad_from_output(yv, a.act, a.slope);
      gv[c] = g;
      sb[c] += g;
    }
    if (a.p_dtype == ICF_BF16 && a.p_pitch == 8) {
      // the whole 16-byte pixel in one store, pitch padding as ZEROS: consumers that reduce over the padded pixel (the
      // channel-major conv kernel reads the raw 8-channel rows) must never meet uninitialised memory there
      __nv_bfloat162 p0 = __floats2bfloat162_rn(gv[0], gv[1]), p1 = __floats2bfloat162_rn(gv[2], gv[3]);
      uint4 w = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), 0u, 0u);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dPre) + pix * 8) = w;
    } else {
      for (int c = 0; c < a.C; ++c) icf::st_any(a.dPre, a.p_dtype, pix * a.p_pitch + c, gv[c]);
    }
  }
  if (a.dbias) {
    for (int c = 0; c < a.C; ++c) {
      const float t = icf::block_sum(sb[c], red);
      if (threadIdx.x == 0) atomicAdd(a.dbias + (a.bias_mod > 0 ? c % a.bias_mod : c), t);
    }
  }
}

// fused backward of [bias -> act -> dropout] (+ BatchNorm(+dropout) that consumed the output)
__global__ void act_backward_kernel(const icf_actbwd_args a) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float sb = 0.f;
  if (c < a.C) {
    float g_scale = 1.f, m0 = 0.f, m1 = 0.f, mu = 0.f, is = 0.f;
    const bool bn = a.bn_sums != nullptr;
    if (bn) {
      const float iw = a.bn_inv_world > 0.f ? a.bn_inv_world : 1.f;
      const float invM = iw / (float)a.pixels;
      mu = a.bn_mean[c];
      is = a.bn_invstd[c];
      g_scale = a.bn_gamma[c] * is;
      m0 = a.bn_sums[c] * invM;
      m1 = a.bn_sums[a.C + c] * invM;
      if (blockIdx.y == 0 && py == 0) {
        if (a.bn_dgamma) a.bn_dgamma[c] += a.bn_sums[a.C + c] * iw;
        if (a.bn_dbeta) a.bn_dbeta[c] += a.bn_sums[c] * iw;
      }
    }
    for (int64_t pix = (int64_t)blockIdx.y * 8 + py; pix < a.pixels; pix += (int64_t)gridDim.y * 8) {
      const int64_t n = sample_of(pix, a.pixels_per_sample);
      float g = icf::ld_any(a.dOut, a.d_dtype, pix * a.d_pitch + c);
      const float yv = icf::ld_any(a.y, a.y_dtype, pix * a.y_pitch + c);
      if (bn) {
        if (a.bn_mask) g *= a.bn_mask[n * a.bn_mask_pitch + c];
        const float xh = (yv - mu) * is;
        g = g_scale * (g - m0 - xh * m1);
      }
      if (a.out_mask) g *= a.out_mask[n * a.mask_pitch + c];
      g *= icf::act_grad_from_output(yv, a.act, a.slope);
      icf::st_any(a.dPre, a.p_dtype, pix * a.p_pitch + c, g);
      sb += g;
    }
  }
  if (a.dbias) {
    red[py][cx] = sb;
    __syncthreads();
    if (py == 0 && c < a.C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][cx];
      atomicAdd(a.dbias + (a.bias_mod > 0 ? c % a.bias_mod : c), t);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// 8-channel vector variants (bf16: one 16-byte access, fp32: two) of the HBM-bound passes.  A thread keeps
// the same channel group for its whole pixel loop, so per-channel parameters and partial sums live in
// registers; block partials meet in shared memory, then one global atomic per channel per block.
// ------------------------------------------------------------------------------------------------
struct V8 { float v[8]; };

__device__ __forceinline__ V8 ld8(const void* base, int dtype, int64_t idx) {
  V8 r;
  if (dtype == ICF_F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  return r;
}
__device__ __forceinline__ V8 cvt8(const uint4& u) {          // 8 packed bf16 -> fp32
  V8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return r;
}
__device__ __forceinline__ uint4 ldraw8(const void* base, int64_t idx) {
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
}
// VU (template parameter of the streaming kernels): pixels in flight per thread in the bf16 loops (memory-level
// parallelism); tuning aids ICF_EW_VU (2, 4, 8) and ICF_EW_CAP (blocks per SM of the grid)
inline int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

__device__ __forceinline__ void st8(void* base, int dtype, int64_t idx, const V8& r) {
  if (dtype == ICF_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
  } else {
    uint4 u;
    uint32_t* w = &u.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 t = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  }
}
__device__ __forceinline__ V8 ldf8(const float* p) {   // 8 consecutive floats, 16-byte aligned
  V8 r;
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

constexpr int VT = 256;          // threads per block of the vector kernels
constexpr int GPB_MAX = 32;      // channel groups (of 8) a block spans: <= 32 keeps every reduction on the shuffle + row path of
                                 // flush_partials (wider blocks fell into its shared-memory float-atomic path: the 512- and
                                 // 1024-channel 1x1 layers paid 16-18 us for 4-8 MB tensors)
constexpr int VSM = 2048;        // channels a block can hold partial sums for

// flush per-thread channel partials, then one global atomic per channel per block.  Up to 256 channels per block
// (the layers of the step): lanes that share a channel group combine through shuffles, each warp stores its row, the
// rows are summed — no shared-memory atomics (float atomics on shared memory are compare-and-swap loops; 64 threads
// on one address cost tens of microseconds per block).  Wider blocks keep the atomic path (<= 4 threads per address).
template <int NS>
__device__ __forceinline__ void flush_partials(float (*acc)[8], int cbase_block, int c0, int C, int span,
                                               float* const* dst, float* sm) {
  const int gpb = span >> 3;
  if (gpb <= 32 && (gpb & (gpb - 1)) == 0 && blockDim.x == VT) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = gpb; o < 32; o <<= 1) {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[s][j] += __shfl_xor_sync(0xffffffffu, acc[s][j], o);
    }
    if (lane < gpb) {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[(warp * NS + s) * span + lane * 8 + j] = acc[s][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS * span; i += VT) {
      const int s = i / span, cc = i - s * span, c = cbase_block + cc;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < VT / 32; ++w) t += sm[(w * NS + s) * span + cc];
      if (c < C && dst[s]) atomicAdd(dst[s] + c, t);
    }
    return;
  }
  for (int i = threadIdx.x; i < NS * span; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  if (c0 < C) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sm[s * span + (c0 - cbase_block) + j], acc[s][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * span; i += blockDim.x) {
    const int s = i / span, c = cbase_block + (i - s * span);
    if (c < C && dst[s]) atomicAdd(dst[s] + c, sm[i]);
  }
}

// thread -> (channel group, first pixel, pixel stride).  cg = C/8 groups; a block covers min(cg, VT) groups.
// Each block owns a CONTIGUOUS range of pixel rows [.., pix_end) and walks it `rows` pixels at a time, so a thread stays
// inside one image for many iterations: the sample index (a division) and the per-(sample, channel) Dropout2d mask
// values are refreshed only when the image changes (MaskCache).
struct VMap { int c0, cbase_block, span; int64_t pix0, pstride, pix_end; };
__device__ __forceinline__ VMap vmap(int C, int64_t pixels) {
  const int cg = C >> 3;
  const int gpb = cg < GPB_MAX ? cg : GPB_MAX;        // channel groups per block
  const int cblocks = (cg + gpb - 1) / gpb;           // blocks along channels
  const int cb = blockIdx.x % cblocks;
  const int64_t pb = blockIdx.x / cblocks, npb = gridDim.x / cblocks;
  const int rows = VT / gpb;                          // pixel rows per block pass
  VMap m;
  m.cbase_block = cb * gpb * 8;
  m.span = gpb * 8;
  m.c0 = m.cbase_block + (threadIdx.x % gpb) * 8;
  int64_t chunk = (pixels + npb - 1) / npb;
  chunk = (chunk + rows - 1) / rows * rows;
  m.pix0 = pb * chunk + threadIdx.x / gpb;
  m.pstride = rows;
  m.pix_end = (pb + 1) * chunk < pixels ? (pb + 1) * chunk : pixels;
  if ((int)threadIdx.x / gpb >= rows) m.c0 = 1 << 30;   // leftover threads when gpb does not divide the block
  return m;
}
struct MaskCache { int64_t lo = 1, hi = 0; V8 v; };
__device__ __forceinline__ const V8& cached_mask(MaskCache& mc, const float* mask, int64_t pitch, int64_t pix, int pps, int c0) {
  if (pix < mc.lo || pix >= mc.hi) {
    const int64_t n = sample_of(pix, pps);
    mc.lo = n * pps;
    mc.hi = mc.lo + pps;
    mc.v = ldf8(mask + n * pitch + c0);
  }
  return mc.v;
}

inline int vgrid(int C, int64_t pixels, int per_sm = 8) {
  static const int cap_env = env_int("ICF_EW_CAP");
  if (cap_env > 0) per_sm = cap_env;
  const int cg = C >> 3;
  const int gpb = cg < GPB_MAX ? cg : GPB_MAX;
  const int cblocks = (cg + gpb - 1) / gpb;
  const int rows = VT / gpb;
  int64_t pblocks = (pixels + rows - 1) / rows;
  const int64_t cap = ((int64_t)icf::sm_count() * per_sm + cblocks - 1) / cblocks;
  if (pblocks > cap) pblocks = cap;
  if (pblocks < 1) pblocks = 1;
  return (int)(pblocks * cblocks);
}

template <int VU>
__global__ void __launch_bounds__(VT) scale_shift_mask_v8(const void* y, int ydt, int ypitch, void* u, int udt, int upitch,
                                                          int64_t pixels, int pps, int C, const float* scale,
                                                          const float* shift, const float* mask, int mpitch) {
  const VMap m = vmap(C, pixels);
  if (m.c0 >= C) return;
  V8 sc, sh;
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc.v[j] = scale ? scale[m.c0 + j] : 1.f; sh.v[j] = (scale && shift) ? shift[m.c0 + j] : 0.f; }
  MaskCache mc;
  auto process = [&](int64_t pix, V8 v) {
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] = fmaf(v.v[j], sc.v[j], sh.v[j]);
    }
    if (mask) {
      const V8& mk = cached_mask(mc, mask, mpitch, pix, pps, m.c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] *= mk.v[j];
    }
    st8(u, udt, pix * upitch + m.c0, v);
  };
  if (ydt == ICF_BF16) {
    for (int64_t pix = m.pix0; pix < m.pix_end; pix += VU * m.pstride) {
      uint4 yr[VU];
#pragma unroll
      for (int k = 0; k < VU; ++k) {
        const int64_t pu = pix + k * m.pstride;
        if (pu < m.pix_end) yr[k] = ldraw8(y, pu * ypitch + m.c0);
      }
#pragma unroll
      for (int k = 0; k < VU; ++k) {
        const int64_t pu = pix + k * m.pstride;
        if (pu < m.pix_end) process(pu, cvt8(yr[k]));
      }
    }
  } else {
    for (int64_t pix = m.pix0; pix < m.pix_end; pix += m.pstride) process(pix, ld8(y, ydt, pix * ypitch + m.c0));
  }
}

template <int VU>
__global__ void __launch_bounds__(VT, VU > 4 ? 2 : 3) bn_bwd_reduce_v8(const void* dU, int ddt, int dpitch, const void* y, int ydt, int ypitch,
                                                       int64_t pixels, int pps, int C, const float* mask, int mpitch,
                                                       const float* mean, const float* invstd, float* sums) {
  __shared__ float sm[2 * VSM];
  const VMap m = vmap(C, pixels);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (m.c0 < C) {
    const V8 mu = ldf8(mean + m.c0), is = ldf8(invstd + m.c0);
    MaskCache mc;
    auto process = [&](int64_t pix, V8 g, const V8& yv) {
      if (mask) {
        const V8& mk = cached_mask(mc, mask, mpitch, pix, pps, m.c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] *= mk.v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {        // raw moments; sum g*xhat = invstd*(sum g*y - mean*sum g) is applied once below
        acc[0][j] += g.v[j];
        acc[1][j] = fmaf(g.v[j], yv.v[j], acc[1][j]);
      }
    };
    if (ddt == ICF_BF16 && ydt == ICF_BF16) {
      for (int64_t pix = m.pix0; pix < m.pix_end; pix += VU * m.pstride) {
        uint4 gr[VU], yr[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          const int64_t pu = pix + u * m.pstride;
          if (pu < m.pix_end) {
            gr[u] = ldraw8(dU, pu * dpitch + m.c0);
            yr[u] = ldraw8(y, pu * ypitch + m.c0);
          }
        }
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          const int64_t pu = pix + u * m.pstride;
          if (pu < m.pix_end) process(pu, cvt8(gr[u]), cvt8(yr[u]));
        }
      }
    } else {
      for (int64_t pix = m.pix0; pix < m.pix_end; pix += m.pstride)
        process(pix, ld8(dU, ddt, pix * dpitch + m.c0), ld8(y, ydt, pix * ypitch + m.c0));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[1][j] = is.v[j] * (acc[1][j] - mu.v[j] * acc[0][j]);
  }
  float* dst[2] = {sums, sums + C};
  flush_partials<2>(acc, m.cbase_block, m.c0, C, m.span, dst, sm);
}

__global__ void __launch_bounds__(VT) col_stats_v8(const void* y, int ydt, int ypitch, int64_t pixels, int C, float* stats) {
  __shared__ float sm[2 * VSM];
  const VMap m = vmap(C, pixels);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (m.c0 < C) {
    for (int64_t pix = m.pix0; pix < m.pix_end; pix += m.pstride) {
      const V8 v = ld8(y, ydt, pix * ypitch + m.c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += v.v[j]; acc[1][j] = fmaf(v.v[j], v.v[j], acc[1][j]); }
    }
  }
  float* dst[2] = {stats, stats + C};
  flush_partials<2>(acc, m.cbase_block, m.c0, C, m.span, dst, sm);
}

template <int VU>
__global__ void __launch_bounds__(VT, 2) act_backward_v8(const icf_actbwd_args a) {
  __shared__ float sm[VSM];
  const VMap m = vmap(a.C, a.pixels);
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  if (m.c0 < a.C) {
    const bool bn = a.bn_sums != nullptr;
    // BatchNorm backward folded into one affine form per channel:
    //   gamma*invstd*(g - m0 - (y - mu)*invstd*m1)  =  cA*g + cB*y + cC
    V8 cA, cB, cC;
    if (bn) {
      const float iw = a.bn_inv_world > 0.f ? a.bn_inv_world : 1.f;
      const float invM = iw / (float)a.pixels;
      const V8 mu = ldf8(a.bn_mean + m.c0), is = ldf8(a.bn_invstd + m.c0);
      const V8 ga = ldf8(a.bn_gamma + m.c0), s0 = ldf8(a.bn_sums + m.c0), s1 = ldf8(a.bn_sums + a.C + m.c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gs = ga.v[j] * is.v[j], m0 = s0.v[j] * invM, m1 = s1.v[j] * invM;
        cA.v[j] = gs;
        cB.v[j] = -gs * is.v[j] * m1;
        cC.v[j] = gs * (mu.v[j] * is.v[j] * m1 - m0);
      }
      if (m.pix0 == 0) {            // exactly one thread per channel group has pix0 == 0
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (a.bn_dgamma) a.bn_dgamma[m.c0 + j] += s1.v[j] * iw;
          if (a.bn_dbeta) a.bn_dbeta[m.c0 + j] += s0.v[j] * iw;
        }
      }
    }
    MaskCache mc_bn, mc_out;
    auto process = [&](int64_t pix, V8 g, const V8& yv) {
      if (bn) {
        if (a.bn_mask) {
          const V8& mk = cached_mask(mc_bn, a.bn_mask, a.bn_mask_pitch, pix, a.pixels_per_sample, m.c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) g.v[j] *= mk.v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] = fmaf(cA.v[j], g.v[j], fmaf(cB.v[j], yv.v[j], cC.v[j]));
      }
      if (a.out_mask) {
        const V8& mk = cached_mask(mc_out, a.out_mask, a.mask_pitch, pix, a.pixels_per_sample, m.c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] *= mk.v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g.v[j] *= icf::act_grad_from_output(yv.v[j], a.act, a.slope);
        acc[0][j] += g.v[j];
      }
      st8(a.dPre, a.p_dtype, pix * a.p_pitch + m.c0, g);
    };
    if (a.d_dtype == ICF_BF16 && a.y_dtype == ICF_BF16) {
      // raw 16-byte loads of VU pixels first, math afterwards: enough bytes in flight at 2-3 blocks per SM
      for (int64_t pix = m.pix0; pix < m.pix_end; pix += VU * m.pstride) {
        uint4 gr[VU], yr[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          const int64_t pu = pix + u * m.pstride;
          if (pu < m.pix_end) {
            gr[u] = ldraw8(a.dOut, pu * a.d_pitch + m.c0);
            yr[u] = ldraw8(a.y, pu * a.y_pitch + m.c0);
          }
        }
#pragma unroll
        for (int u = 0; u < VU; ++u) {
          const int64_t pu = pix + u * m.pstride;
          if (pu < m.pix_end) process(pu, cvt8(gr[u]), cvt8(yr[u]));
        }
      }
    } else {
      for (int64_t pix = m.pix0; pix < m.pix_end; pix += m.pstride)
        process(pix, ld8(a.dOut, a.d_dtype, pix * a.d_pitch + m.c0), ld8(a.y, a.y_dtype, pix * a.y_pitch + m.c0));
    }
  }
  if (a.dbias) {
    float* dst[1] = {a.dbias};
    flush_partials<1>(acc, m.cbase_block, m.c0, a.C, m.span, dst, sm);
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int vu_choice() {
  static const int v = env_int("ICF_EW_VU");
  return v;
}

}  // namespace

// BatchNorm statistics of a stored tensor (used after a tensor-core conv, whose epilogue does not reduce columns)
int icf_launch_col_stats(const void* y, int ydt, int ypitch, int64_t pixels, int C, float* stats, cudaStream_t st) {
  if ((C & 7) || (ypitch & 7) || !al16(y)) { icf::set_error("col_stats: needs 8-channel aligned tensors"); return 1; }
  col_stats_v8<<<vgrid(C, pixels), VT, 0, st>>>(y, ydt, ypitch, pixels, C, stats);
  return icf::check_launch("col_stats_v8");
}

namespace {
inline bool al32f(const void* p, int dt) { return (reinterpret_cast<uintptr_t>(p) & (dt == ICF_F32 ? 15 : 15)) == 0; }

// ------------------------------------------------------------------------------------------------
// BCE with logits (mean) forward+backward; sigmoid mean (phase-D scores). Single block.
// ------------------------------------------------------------------------------------------------
__global__ void bce_logits_kernel(const void* logits, int ldt, int lpitch, int n, float target, float weight,
                                  float* loss_out, void* dl, int ddt, int dpitch) {
  __shared__ float red[32];
  float s = 0.f;
  const float invn = 1.f / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float l = icf::ld_any(logits, ldt, (int64_t)i * lpitch);
    s += fmaxf(l, 0.f) - l * target + log1pf(expf(-fabsf(l)));
    if (dl) {
      const float sig = 1.f / (1.f + expf(-l));
      icf::st_any(dl, ddt, (int64_t)i * dpitch, weight * (sig - target) * invn);
    }
  }
  s = icf::block_sum(s, red);
  if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, weight * s * invn);
}

__global__ void sigmoid_mean_kernel(const void* logits, int ldt, int lpitch, int n, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    s += 1.f / (1.f + expf(-icf::ld_any(logits, ldt, (int64_t)i * lpitch)));
  s = icf::block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s / (float)n);
}

// ------------------------------------------------------------------------------------------------
// Adam over a flat buffer. state = {step, lr, beta1, beta2, eps, grad_scale}; the tick kernel advances
// the step so that a captured graph replays correctly.
// ------------------------------------------------------------------------------------------------
__global__ void adam_tick_kernel(float* state) { state[0] += 1.f; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* __restrict__ state) {
  const float step = state[0], lr = state[1], b1 = state[2], b2 = state[3], eps = state[4], gs = state[5];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2s = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vvp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = gg[k] * gs;
        mm[k] = b1 * mm[k] + (1.f - b1) * gr;
        vvp[k] = b2 * vvp[k] + (1.f - b2) * gr * gr;
        pp[k] -= step_size * mm[k] / (sqrtf(vvp[k]) / bc2s + eps);
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t k = i; k < n; ++k) {
        const float gr = g[k] * gs;
        m[k] = b1 * m[k] + (1.f - b1) * gr;
        v[k] = b2 * v[k] + (1.f - b2) * gr * gr;
        p[k] -= step_size * m[k] / (sqrtf(v[k]) / bc2s + eps);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight pack / gradient unpack (3-index permutation, see icf.h)
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ src, void* dst, int ddt, const icf_perm p) {
  const int64_t total = p.d0_pad * p.d1 * p.d2_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i2 = i % p.d2_pad;
    const int64_t t = i / p.d2_pad;
    const int64_t i1 = t % p.d1, i0 = t / p.d1;
    float v = 0.f;
    if (i2 < p.d2 && i0 < p.d0) v = src[i0 * p.s0 + i1 * p.s1 + i2 * p.s2];
    icf::st_any(dst, ddt, i, v);
  }
}

// kind 0 body with 32-bit index arithmetic, for the jobs whose middle index (the filter tap) is contiguous in the source
// (s1 == 1: every Conv2d / ConvTranspose2d weight).  A (row i0, 64-wide i2 chunk) tile goes through shared memory: it is READ
// in source order (runs of d1 contiguous floats) and WRITTEN in packed order (128-byte runs of bf16), so both sides move whole
// sectors.  Element-wise gathering of the packed order read one 4-byte element per 32-byte sector: 399 MB of L2->SM traffic
// for 20 MB of weights (ncu, profiles/r02_ncu_small_kernels.txt), 70 us per network.
constexpr int PK_CH = 64, PK_T = 32;
__device__ __forceinline__ void pack_job_tiled(const icf_pack_job& jb, float* tile) {
  const icf_perm& p = jb.p;
  const uint32_t d0 = (uint32_t)p.d0, d1 = (uint32_t)p.d1, d2 = (uint32_t)p.d2, d2p = (uint32_t)p.d2_pad;
  const uint32_t s0 = (uint32_t)p.s0, s2 = (uint32_t)p.s2;
  const uint32_t chunks = (d2p + PK_CH - 1) / PK_CH, ntiles = (uint32_t)p.d0_pad * chunks;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t i0 = t / chunks, c0 = (t - i0 * chunks) * PK_CH;
    const uint32_t cw = min((uint32_t)PK_CH, d2p - c0);                  // chunk width (multiple of 8)
    // load: e = i2l * d1 + i1 (source order)
    for (uint32_t e = threadIdx.x; e < cw * d1; e += blockDim.x) {
      const uint32_t i2l = e / d1, i1 = e - i2l * d1, i2 = c0 + i2l;
      float v = 0.f;
      if (i0 < d0 && i2 < d2) v = jb.src[i0 * s0 + i2 * s2 + i1];
      tile[i1 * (PK_CH + 1) + i2l] = v;
    }
    __syncthreads();
    // store: pairs of consecutive i2 (packed order)
    const uint32_t hw = cw >> 1;
    for (uint32_t e = threadIdx.x; e < hw * d1; e += blockDim.x) {
      const uint32_t i1 = e / hw, pp = e - i1 * hw;
      const float v0 = tile[i1 * (PK_CH + 1) + 2 * pp], v1 = tile[i1 * (PK_CH + 1) + 2 * pp + 1];
      const uint32_t q = (((i0 * d1 + i1) * d2p) + c0) / 2 + pp;
      if (jb.dst_dtype == ICF_F32) reinterpret_cast<float2*>(jb.dst)[q] = make_float2(v0, v1);
      else reinterpret_cast<__nv_bfloat162*>(jb.dst)[q] = __floats2bfloat162_rn(v0, v1);
    }
    __syncthreads();
  }
}

// 2-D transpose (d1 == 1, source contiguous along i0: the dgrad copy of a 1x1 conv / Linear weight): 32 x 64 tiles
__device__ __forceinline__ void pack_job_transpose(const icf_pack_job& jb, float* tile) {
  const icf_perm& p = jb.p;
  const uint32_t d0 = (uint32_t)p.d0, d2 = (uint32_t)p.d2, d2p = (uint32_t)p.d2_pad, s2 = (uint32_t)p.s2, d0p = (uint32_t)p.d0_pad;
  const uint32_t t0 = (d0p + 31) / 32, t2 = (d2p + PK_CH - 1) / PK_CH;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 256 threads: 8 rows of 32
  for (uint32_t t = blockIdx.x; t < t0 * t2; t += gridDim.x) {
    const uint32_t b0 = (t / t2) * 32, b2 = (t % t2) * PK_CH;
    for (uint32_t r = ty; r < PK_CH; r += 8) {                           // tile[i2l][i0l], coalesced along i0
      const uint32_t i0 = b0 + tx, i2 = b2 + r;
      tile[r * 33 + tx] = (i0 < d0 && i2 < d2) ? jb.src[i0 + i2 * s2] : 0.f;
    }
    __syncthreads();
    for (uint32_t r = ty; r < 32; r += 8) {                              // one packed row (i0) per warp pass, 32 pairs along i2
      const uint32_t i0 = b0 + r, i2 = b2 + 2 * tx;
      if (i0 < d0p && i2 < d2p) {
        const float v0 = tile[(2 * tx) * 33 + r], v1 = tile[(2 * tx + 1) * 33 + r];
        const uint32_t q = (i0 * d2p + i2) >> 1;
        if (jb.dst_dtype == ICF_F32) reinterpret_cast<float2*>(jb.dst)[q] = make_float2(v0, v1);
        else reinterpret_cast<__nv_bfloat162*>(jb.dst)[q] = __floats2bfloat162_rn(v0, v1);
      }
    }
    __syncthreads();
  }
}

// same arithmetic without the tile (any strides): two packed elements per thread
__device__ __forceinline__ void pack_job32(const icf_pack_job& jb) {
  const icf_perm& p = jb.p;
  const uint32_t d2p = (uint32_t)p.d2_pad, d1 = (uint32_t)p.d1, d2 = (uint32_t)p.d2, d0 = (uint32_t)p.d0;
  const uint32_t s0 = (uint32_t)p.s0, s1 = (uint32_t)p.s1, s2 = (uint32_t)p.s2;
  const uint32_t pairs = (uint32_t)((p.d0_pad * p.d1 * p.d2_pad) >> 1);          // d2_pad is even on this path
  const uint32_t half = d2p >> 1;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += gridDim.x * blockDim.x) {
    const uint32_t t = q / half, i2 = (q - t * half) * 2;
    const uint32_t i0 = t / d1, i1 = t - i0 * d1;
    float v0 = 0.f, v1 = 0.f;
    if (i0 < d0) {
      const uint32_t base = i0 * s0 + i1 * s1 + i2 * s2;
      if (i2 < d2) v0 = jb.src[base];
      if (i2 + 1 < d2) v1 = jb.src[base + s2];
    }
    if (jb.dst_dtype == ICF_F32) {
      reinterpret_cast<float2*>(jb.dst)[q] = make_float2(v0, v1);
    } else {
      reinterpret_cast<__nv_bfloat162*>(jb.dst)[q] = __floats2bfloat162_rn(v0, v1);
    }
  }
}

__global__ void pack_multi_kernel(const icf_pack_job* __restrict__ jobs) {
  __shared__ float pk_tile[PK_CH * 33 + 64];   // >= PK_T * (PK_CH + 1) (tap tiles) and PK_CH * 33 (transpose tiles)
  const icf_pack_job jb = jobs[blockIdx.y];
  if (jb.kind == 0) {
    const icf_perm p = jb.p;
    const int64_t total = p.d0_pad * p.d1 * p.d2_pad;
    const int64_t src_span = (p.d0 - 1) * p.s0 + (p.d1 - 1) * p.s1 + (p.d2 - 1) * p.s2;
    if ((p.d2_pad & 1) == 0 && total < 0x7fffffffLL && src_span < 0x7fffffffLL && p.d0 > 0 && p.d1 > 0 && p.d2 > 0 &&
        (reinterpret_cast<uintptr_t>(jb.dst) & 7) == 0) {
      if (p.s1 == 1 && p.d1 <= PK_T && (p.d2_pad & 7) == 0 && p.d1 > 1) pack_job_tiled(jb, pk_tile);
      else if (p.d1 == 1 && p.s0 == 1 && p.s2 > 1 && blockDim.x == 256) pack_job_transpose(jb, pk_tile);
      else pack_job32(jb);
      return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i2 = i % p.d2_pad;
      const int64_t t = i / p.d2_pad;
      const int64_t i1 = t % p.d1, i0 = t / p.d1;
      float v = 0.f;
      if (i2 < p.d2 && i0 < p.d0) v = jb.src[i0 * p.s0 + i1 * p.s1 + i2 * p.s2];
      icf::st_any(jb.dst, jb.dst_dtype, i, v);
    }
  } else {
    const icf_perm4 p = jb.p4;
    const int64_t total = p.d0 * p.d1 * p.row_pitch;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t col = i % p.row_pitch;
      const int64_t t = i / p.row_pitch;
      const int64_t i1 = t % p.d1, i0 = t / p.d1;
      const int64_t i2 = col / p.d3_pad, i3 = col % p.d3_pad;
      float v = 0.f;
      if (i2 < p.d2 && i3 < p.d3) v = jb.src[i0 * p.s0 + i1 * p.s1 + i2 * p.s2 + i3 * p.s3];
      icf::st_any(jb.dst, jb.dst_dtype, i, v);
    }
  }
}

__global__ void unpack_multi_kernel(const icf_pack_job* __restrict__ jobs) {
  const icf_pack_job jb = jobs[blockIdx.y];
  float* dst = reinterpret_cast<float*>(jb.dst);
  if (jb.kind == 0) {
    const icf_perm p = jb.p;
    const int64_t total = p.d0 * p.d1 * p.d2;
    const int64_t span = (p.d0 - 1) * p.s0 + (p.d1 - 1) * p.s1 + (p.d2 - 1) * p.s2;
    if (total < 0x7fffffffLL && span < 0x7fffffffLL && p.d0 * p.d1 * p.d2_pad < 0x7fffffffLL) {
      // 32-bit index arithmetic: the 64-bit divisions of the general loop kept this copy instruction-bound (ncu: 51 % SM busy
      // at 6 % of DRAM)
      const uint32_t d1 = (uint32_t)p.d1, d2 = (uint32_t)p.d2, d2p = (uint32_t)p.d2_pad;
      const uint32_t s0 = (uint32_t)p.s0, s1 = (uint32_t)p.s1, s2 = (uint32_t)p.s2, tot = (uint32_t)total;
      for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
        const uint32_t t = i / d2, i2 = i - t * d2;
        const uint32_t i0 = t / d1, i1 = t - i0 * d1;
        dst[i0 * s0 + i1 * s1 + i2 * s2] = jb.src[t * d2p + i2];
      }
      return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i2 = i % p.d2;
      const int64_t t = i / p.d2;
      const int64_t i1 = t % p.d1, i0 = t / p.d1;
      dst[i0 * p.s0 + i1 * p.s1 + i2 * p.s2] = jb.src[(i0 * p.d1 + i1) * p.d2_pad + i2];
    }
  } else {
    const icf_perm4 p = jb.p4;
    const int64_t total = p.d0 * p.d1 * p.d2 * p.d3;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      int64_t t = i;
      const int64_t i3 = t % p.d3; t /= p.d3;
      const int64_t i2 = t % p.d2; t /= p.d2;
      const int64_t i1 = t % p.d1, i0 = t / p.d1;
      dst[i0 * p.s0 + i1 * p.s1 + i2 * p.s2 + i3 * p.s3] = jb.src[(i0 * p.d1 + i1) * p.row_pitch + i2 * p.d3_pad + i3];
    }
  }
}

__global__ void unpack_kernel(const float* __restrict__ src, float* dst, const icf_perm p, int atomic_add) {
  const int64_t total = p.d0 * p.d1 * p.d2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i2 = i % p.d2;
    const int64_t t = i / p.d2;
    const int64_t i1 = t % p.d1, i0 = t / p.d1;
    const float v = src[(i0 * p.d1 + i1) * p.d2_pad + i2];
    float* o = dst + i0 * p.s0 + i1 * p.s1 + i2 * p.s2;
    if (atomic_add) atomicAdd(o, v);
    else *o = v;
  }
}

__global__ void pack4_kernel(const float* __restrict__ src, void* dst, int ddt, const icf_perm4 p) {
  const int64_t total = p.d0 * p.d1 * p.row_pitch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t col = i % p.row_pitch, row = i / p.row_pitch;
    const int64_t i1 = row % p.d1, i0 = row / p.d1;
    const int64_t i2 = col / p.d3_pad, i3 = col - i2 * p.d3_pad;
    float v = 0.f;
    if (i2 < p.d2 && i3 < p.d3) v = src[i0 * p.s0 + i1 * p.s1 + i2 * p.s2 + i3 * p.s3];
    icf::st_any(dst, ddt, i, v);
  }
}

__global__ void unpack4_kernel(const float* __restrict__ src, float* dst, const icf_perm4 p) {
  const int64_t total = p.d0 * p.d1 * p.d2 * p.d3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int64_t i3 = t % p.d3; t /= p.d3;
    const int64_t i2 = t % p.d2; t /= p.d2;
    const int64_t i1 = t % p.d1, i0 = t / p.d1;
    dst[i0 * p.s0 + i1 * p.s1 + i2 * p.s2 + i3 * p.s3] = src[(i0 * p.d1 + i1) * p.row_pitch + i2 * p.d3_pad + i3];
  }
}

__global__ void cast_kernel(const void* src, int sdt, void* dst, int ddt, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    icf::st_any(dst, ddt, i, icf::ld_any(src, sdt, i));
}

__global__ void fill_kernel(float* dst, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = v;
}

inline int ew_grid(int64_t total, int per_thread = 1) {
  int64_t blocks = (total + (int64_t)EW_THREADS * per_thread - 1) / ((int64_t)EW_THREADS * per_thread);
  const int64_t cap = (int64_t)icf::sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline int pixel_slabs(int64_t pixels, int chan_groups) {
  int64_t want = ((int64_t)icf::sm_count() * 8) / (chan_groups > 0 ? chan_groups : 1);
  int64_t maxs = (pixels + 7) / 8;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace

extern "C" {

int icf_argmax_rows(const void* x, int32_t x_dtype, int32_t n, int32_t k, int32_t* out, void* stream) {
  ICF_REQUIRE(x && out && n >= 0 && k >= 1 && x_dtype >= 0 && x_dtype <= 3, "icf_argmax_rows: bad arguments");
  if (n == 0) return 0;
  argmax_rows_kernel<<<icf::cdiv(n, 128), 128, 0, icf::as_stream(stream)>>>(x, x_dtype, n, k, out);
  return icf::check_launch("argmax_rows");
}

int icf_image_features_fwd(const icf_imgfeat_args* a, void* stream) {
  ICF_REQUIRE(a && a->x && a->feat, "icf_image_features_fwd: null pointer");
  ICF_REQUIRE(a->n_emb >= 0 && a->n_cont >= 0 && a->n_emb <= ICF_MAX_PLANES && a->n_cont <= ICF_MAX_PLANES &&
                  1 + a->n_emb + a->n_cont <= a->feat_pitch && 1 + a->n_emb + a->n_cont <= 8,
              "icf_image_features_fwd: %d emb + %d cont planes do not fit pitch %d (at most 8 feature channels)", a->n_emb,
              a->n_cont, a->feat_pitch);
  const int64_t total = (int64_t)a->N * (a->H + 2 * a->pad) * (a->W + 2 * a->pad);
  if (total == 0) return 0;
  const int sms = icf::sm_count();
  const int Hp = a->H + 2 * a->pad, Wp = a->W + 2 * a->pad;
  ICF_REQUIRE(Wp <= IMGFEAT_MAX_W && Hp <= IMGFEAT_MAX_W, "icf_image_features_fwd: images larger than 1024 (padded) are not supported");
  // The per-sample chain (index load -> table load -> tanh -> barrier -> image loads -> stores) is latency, not bandwidth: with
  // 256-thread blocks three resident blocks per SM each walked ~9 samples of 28x28 pixels one after the other (51 us for 67 MB).
  // Small images get 64-thread blocks, 12+ of them per SM, so that many samples are in flight on every SM.
  const int threads = (int64_t)Hp * Wp <= 4096 ? 64 : EW_THREADS;
  const int per_sm = threads == 64 ? 16 : 8;
  int gx = a->N < sms * per_sm ? a->N : sms * per_sm, gy = 1;
  while (gx * gy < 2 * sms && Hp / (gy * 2) >= 8 && gy < 64) gy *= 2;            // few, large images: split the rows
  imgfeat_fwd_kernel<<<dim3((unsigned)gx, (unsigned)gy), threads, 0, icf::as_stream(stream)>>>(*a);
  return icf::check_launch("imgfeat_fwd");
}

int icf_image_features_bwd(const icf_imgfeat_args* a, void* stream) {
  ICF_REQUIRE(a && a->dfeat, "icf_image_features_bwd: null pointer");
  if (a->n_emb == 0 || a->N == 0) return 0;
  const int64_t warps = (int64_t)a->N * a->n_emb * 256;
  if (icf::cdiv(a->H, 16) * icf::cdiv(a->W, 16) <= 8) {      // a handful of pixels per embedding cell
    imgfeat_bwd_small_kernel<<<icf::cdiv(warps, EW_THREADS), EW_THREADS, 0, icf::as_stream(stream)>>>(*a);
    return icf::check_launch("imgfeat_bwd_small");
  }
  imgfeat_bwd_kernel<<<icf::cdiv(warps * 32, EW_THREADS), EW_THREADS, 0, icf::as_stream(stream)>>>(*a);
  return icf::check_launch("imgfeat_bwd");
}

int icf_latent_features_fwd(const icf_latfeat_args* a, void* stream) {
  ICF_REQUIRE(a && a->z && a->feat, "icf_latent_features_fwd: null pointer");
  ICF_REQUIRE(a->latent + 256 * a->n_emb + a->n_cont <= a->feat_pitch,
              "icf_latent_features_fwd: feature row %d > pitch %d", a->latent + 256 * a->n_emb + a->n_cont,
              a->feat_pitch);
  const int64_t total = (int64_t)a->N * a->feat_pitch;
  if (total == 0) return 0;
  latfeat_fwd_kernel<<<icf::cdiv(total, EW_THREADS), EW_THREADS, 0, icf::as_stream(stream)>>>(*a);
  return icf::check_launch("latfeat_fwd");
}

int icf_latent_features_bwd(const icf_latfeat_args* a, void* stream) {
  ICF_REQUIRE(a && a->dfeat, "icf_latent_features_bwd: null pointer");
  if (a->N == 0) return 0;
  cudaStream_t st = icf::as_stream(stream);
  int tot_k = 0;
  bool any_row = a->dz != nullptr;
  for (int e = 0; e < a->n_emb; ++e) { tot_k += a->emb_k[e]; any_row = any_row || a->donehot[e]; }
  for (int e = 0; e < a->n_cont; ++e) any_row = any_row || a->dcont[e];
  if (any_row) {
    const int64_t total = (int64_t)a->N * (a->latent + a->n_cont + tot_k);
    latfeat_bwd_rows_kernel<<<icf::cdiv(total, EW_THREADS), EW_THREADS, 0, st>>>(*a);
    if (int r = icf::check_launch("latfeat_bwd_rows")) return r;
  }
  for (int e = 0; e < a->n_emb; ++e) {
    if (!a->demb_table[e]) continue;
    latfeat_bwd_table_kernel<<<dim3(a->emb_k[e], icf::cdiv(a->N, 64)), 256, 0, st>>>(*a, e);
    if (int r = icf::check_launch("latfeat_bwd_table")) return r;
  }
  return 0;
}

int icf_bn_finalize(float* stats, int32_t C, double count, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  ICF_REQUIRE(stats && scale && shift && C > 0 && count > 0, "icf_bn_finalize: bad arguments");
  bn_finalize_kernel<<<icf::cdiv(C, 128), 128, 0, icf::as_stream(stream)>>>(
      stats, C, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, scale,
      shift, save_mean, save_invstd);
  return icf::check_launch("bn_finalize");
}

int icf_scale_shift_mask(const void* y, int32_t y_dtype, int32_t y_pitch, void* u, int32_t u_dtype,
                         int32_t u_pitch, int64_t pixels, int32_t pixels_per_sample, int32_t C,
                         const float* scale, const float* shift, const float* mask, int32_t mask_pitch,
                         void* stream) {
  ICF_REQUIRE(y && u && C > 0 && pixels_per_sample > 0, "icf_scale_shift_mask: bad arguments");
  if (pixels == 0) return 0;
  if ((C & 7) == 0 && (y_pitch & 7) == 0 && (u_pitch & 7) == 0 && al16(y) && al16(u) && (!mask || ((mask_pitch & 3) == 0 && al16(mask))) &&
      (!scale || al16(scale)) && (!shift || al16(shift))) {
#define ICF_SSM(V) scale_shift_mask_v8<V><<<vgrid(C, pixels, 3), VT, 0, icf::as_stream(stream)>>>(y, y_dtype, y_pitch, u, u_dtype, u_pitch, \
                                                          pixels, pixels_per_sample, C, scale, shift, mask, mask_pitch)
    const int vu = vu_choice();
    if (vu == 2) ICF_SSM(2); else if (vu == 8) ICF_SSM(8); else ICF_SSM(4);
#undef ICF_SSM
    return icf::check_launch("scale_shift_mask_v8");
  }
  scale_shift_mask_kernel<<<ew_grid(pixels * C, 4), EW_THREADS, 0, icf::as_stream(stream)>>>(
      y, y_dtype, y_pitch, u, u_dtype, u_pitch, pixels, pixels_per_sample, C, scale, shift, mask, mask_pitch);
  return icf::check_launch("scale_shift_mask");
}

int icf_bn_bwd_reduce(const void* dU, int32_t d_dtype, int32_t d_pitch, const void* y, int32_t y_dtype,
                      int32_t y_pitch, int64_t pixels, int32_t pixels_per_sample, int32_t C, const float* mask,
                      int32_t mask_pitch, const float* save_mean, const float* save_invstd, float* sums,
                      void* stream) {
  ICF_REQUIRE(dU && y && save_mean && save_invstd && sums && C > 0, "icf_bn_bwd_reduce: bad arguments");
  if (pixels == 0) return 0;
  if ((C & 7) == 0 && (d_pitch & 7) == 0 && (y_pitch & 7) == 0 && al16(dU) && al16(y) && al16(save_mean) && al16(save_invstd) &&
      (!mask || ((mask_pitch & 3) == 0 && al16(mask)))) {
#define ICF_BBR(V, PER_SM) bn_bwd_reduce_v8<V><<<vgrid(C, pixels, PER_SM), VT, 0, icf::as_stream(stream)>>>(                 \
      dU, d_dtype, d_pitch, y, y_dtype, y_pitch, pixels, pixels_per_sample, C, mask, mask_pitch, save_mean, save_invstd, sums)
    const int vu = vu_choice();            // one resident wave: measured best (profiles/r01_ew_bench.log)
    if (vu == 2) ICF_BBR(2, 3); else if (vu == 8) ICF_BBR(8, 2); else ICF_BBR(4, 3);
#undef ICF_BBR
    return icf::check_launch("bn_bwd_reduce_v8");
  }
  const int groups = icf::cdiv(C, 32);
  dim3 grid(groups, pixel_slabs(pixels, groups));
  bn_bwd_reduce_kernel<<<grid, 256, 0, icf::as_stream(stream)>>>(dU, d_dtype, d_pitch, y, y_dtype, y_pitch,
                                                                   pixels, pixels_per_sample, C, mask,
                                                                   mask_pitch, save_mean, save_invstd, sums);
  return icf::check_launch("bn_bwd_reduce");
}

int icf_act_backward(const icf_actbwd_args* a, void* stream) {
  ICF_REQUIRE(a && a->dOut && a->y && a->dPre && a->C > 0 && a->pixels_per_sample > 0,
              "icf_act_backward: bad arguments");
  if (a->bn_sums)
    ICF_REQUIRE(a->bn_gamma && a->bn_mean && a->bn_invstd, "icf_act_backward: incomplete BatchNorm state");
  if (a->pixels == 0) return 0;
  {
    bool ok = (a->C & 7) == 0 && (a->d_pitch & 7) == 0 && (a->y_pitch & 7) == 0 && (a->p_pitch & 7) == 0 && al16(a->dOut) &&
              al16(a->y) && al16(a->dPre) && a->bias_mod == 0;
    if (a->out_mask) ok = ok && (a->mask_pitch & 3) == 0 && al16(a->out_mask);
    if (a->bn_sums) ok = ok && al16(a->bn_sums) && al16(a->bn_gamma) && al16(a->bn_mean) && al16(a->bn_invstd) && (a->C & 3) == 0 &&
                      (!a->bn_mask || ((a->bn_mask_pitch & 3) == 0 && al16(a->bn_mask)));
    if (ok) {
      const int vu = vu_choice();
      if (vu == 2) act_backward_v8<2><<<vgrid(a->C, a->pixels, 2), VT, 0, icf::as_stream(stream)>>>(*a);
      else if (vu == 8) act_backward_v8<8><<<vgrid(a->C, a->pixels, 2), VT, 0, icf::as_stream(stream)>>>(*a);
      else act_backward_v8<4><<<vgrid(a->C, a->pixels, 2), VT, 0, icf::as_stream(stream)>>>(*a);
      return icf::check_launch("act_backward_v8");
    }
  }
  if (a->C <= 4 && !a->bn_sums) {
    int64_t blocks = (a->pixels + 255) / 256;
    if (blocks > icf::sm_count() * 8) blocks = icf::sm_count() * 8;
    act_backward_fewc_kernel<<<(unsigned)blocks, 256, 0, icf::as_stream(stream)>>>(*a);
    return icf::check_launch("act_backward_fewc");
  }
  const int groups = icf::cdiv(a->C, 32);
  dim3 grid(groups, pixel_slabs(a->pixels, groups));
  act_backward_kernel<<<grid, 256, 0, icf::as_stream(stream)>>>(*a);
  return icf::check_launch("act_backward");
}

int icf_bce_logits(const void* logits, int32_t l_dtype, int32_t l_pitch, int32_t n, float target, float weight,
                   float* loss_out, void* dlogits, int32_t d_dtype, int32_t d_pitch, void* stream) {
  ICF_REQUIRE(logits && n > 0, "icf_bce_logits: bad arguments");
  bce_logits_kernel<<<1, 1024, 0, icf::as_stream(stream)>>>(logits, l_dtype, l_pitch, n, target, weight,
                                                            loss_out, dlogits, d_dtype, d_pitch);
  return icf::check_launch("bce_logits");
}

int icf_sigmoid_mean(const void* logits, int32_t l_dtype, int32_t l_pitch, int32_t n, float* score_out,
                     void* stream) {
  ICF_REQUIRE(logits && score_out && n > 0, "icf_sigmoid_mean: bad arguments");
  sigmoid_mean_kernel<<<1, 1024, 0, icf::as_stream(stream)>>>(logits, l_dtype, l_pitch, n, score_out);
  return icf::check_launch("sigmoid_mean");
}

int icf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state,
                  void* stream) {
  ICF_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n >= 0, "icf_adam_step: bad arguments");
  ICF_REQUIRE((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) % 16 == 0,
              "icf_adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = icf::as_stream(stream);
  adam_tick_kernel<<<1, 1, 0, st>>>(state);
  if (n > 0) adam_kernel<<<ew_grid(n, 4), EW_THREADS, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, state);
  return icf::check_launch("adam");
}

int icf_pack(const float* src, void* dst, int32_t dst_dtype, const icf_perm* p, void* stream) {
  ICF_REQUIRE(src && dst && p && p->d2_pad >= p->d2 && p->d0_pad >= p->d0, "icf_pack: bad arguments");
  const int64_t total = p->d0_pad * p->d1 * p->d2_pad;
  if (total == 0) return 0;
  pack_kernel<<<ew_grid(total), EW_THREADS, 0, icf::as_stream(stream)>>>(src, dst, dst_dtype, *p);
  return icf::check_launch("pack");
}

int icf_pack_multi(const icf_pack_job* jobs, int32_t n_jobs, int64_t max_elems, void* stream) {
  ICF_REQUIRE(jobs && n_jobs >= 0 && max_elems >= 0, "icf_pack_multi: bad arguments");
  if (n_jobs == 0 || max_elems == 0) return 0;
  int64_t bx = (max_elems + 4 * EW_THREADS - 1) / (4 * EW_THREADS);   // ~4 elements per thread of the largest job
  if (bx > 1024) bx = 1024;
  pack_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_jobs), EW_THREADS, 0, icf::as_stream(stream)>>>(jobs);
  return icf::check_launch("pack_multi");
}

int icf_unpack_multi(const icf_pack_job* jobs, int32_t n_jobs, int64_t max_elems, void* stream) {
  ICF_REQUIRE(jobs && n_jobs >= 0 && max_elems >= 0, "icf_unpack_multi: bad arguments");
  if (n_jobs == 0 || max_elems == 0) return 0;
  int64_t bx = (max_elems + 4 * EW_THREADS - 1) / (4 * EW_THREADS);
  if (bx > 1024) bx = 1024;
  unpack_multi_kernel<<<dim3((unsigned)bx, (unsigned)n_jobs), EW_THREADS, 0, icf::as_stream(stream)>>>(jobs);
  return icf::check_launch("unpack_multi");
}

int icf_unpack(const float* src_packed, float* dst, const icf_perm* p, int32_t atomic_add, void* stream) {
  ICF_REQUIRE(src_packed && dst && p, "icf_unpack: bad arguments");
  const int64_t total = p->d0 * p->d1 * p->d2;
  if (total == 0) return 0;
  unpack_kernel<<<ew_grid(total), EW_THREADS, 0, icf::as_stream(stream)>>>(src_packed, dst, *p, atomic_add);
  return icf::check_launch("unpack");
}

int icf_pack4(const float* src, void* dst, int32_t dst_dtype, const icf_perm4* p, void* stream) {
  ICF_REQUIRE(src && dst && p && p->d3_pad >= p->d3 && p->row_pitch >= p->d2 * p->d3_pad, "icf_pack4: bad arguments");
  const int64_t total = p->d0 * p->d1 * p->row_pitch;
  if (total == 0) return 0;
  pack4_kernel<<<ew_grid(total), EW_THREADS, 0, icf::as_stream(stream)>>>(src, dst, dst_dtype, *p);
  return icf::check_launch("pack4");
}

int icf_unpack4(const float* src_packed, float* dst, const icf_perm4* p, void* stream) {
  ICF_REQUIRE(src_packed && dst && p, "icf_unpack4: bad arguments");
  const int64_t total = p->d0 * p->d1 * p->d2 * p->d3;
  if (total == 0) return 0;
  unpack4_kernel<<<ew_grid(total), EW_THREADS, 0, icf::as_stream(stream)>>>(src_packed, dst, *p);
  return icf::check_launch("unpack4");
}

int icf_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
  ICF_REQUIRE(src && dst && n >= 0, "icf_cast: bad arguments");
  if (n == 0) return 0;
  cast_kernel<<<ew_grid(n, 4), EW_THREADS, 0, icf::as_stream(stream)>>>(src, src_dtype, dst, dst_dtype, n);
  return icf::check_launch("cast");
}

int icf_fill_f32(float* dst, float value, int64_t n, void* stream) {
  ICF_REQUIRE(dst && n >= 0, "icf_fill_f32: bad arguments");
  if (n == 0) return 0;
  fill_kernel<<<ew_grid(n, 4), EW_THREADS, 0, icf::as_stream(stream)>>>(dst, value, n);
  return icf::check_launch("fill");
}

}  // extern "C"
