// SIMT implicit-GEMM convolution kernels: the fp32 path (1e-3 parity mode) and the fallback for shapes the
// tcgen05 path does not take (odd channel counts such as the K=1 logits layer).
//
// forward / dgrad : dst[m][k] = epi( sum_{tap,c} src[pix(m,tap)][c] * w[k][tap][c] ),  m = (n,p,q)
// wgrad           : dw[a][tap][b] += sum_m small[m][a] * big[pix(m,tap)][b]
#include "icf_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 8, NT = 256;

template <typename T, int FORM, typename TO>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const icf_conv_args a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float Ss[2][BN];
  const int tid = threadIdx.x;
  const int PQ = a.P * a.Q;
  const int64_t M = (int64_t)a.N * PQ;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int taps = a.R * a.S;

  // loader role: one (row, channel-pair) per thread for both operand tiles
  const int lrow = tid >> 2, lc = (tid & 3) * 2;
  const int64_t lm = m0 + lrow;
  const bool lvalid = lm < M;
  int ln = 0, lp = 0, lq = 0;
  if (lvalid) {
    ln = (int)(lm / PQ);
    int rem = (int)(lm - (int64_t)ln * PQ);
    lp = rem / a.Q;
    lq = rem - lp * a.Q;
  }
  const int wrow = n0 + lrow;
  const bool bvalid = wrow < a.w_rows;
  const T* __restrict__ src = reinterpret_cast<const T*>(a.src);
  const T* __restrict__ w = reinterpret_cast<const T*>(a.w);

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps; ++tap) {
    const int r = tap / a.S, s = tap - r * a.S;
    int iy, ix;
    bool v = lvalid;
    if (FORM == ICF_FORM_GATHER) {
      iy = lp * a.stride - a.pad + r;
      ix = lq * a.stride - a.pad + s;
      v = v && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    } else {
      const int uy = lp + a.pad - r, ux = lq + a.pad - s;
      v = v && uy >= 0 && ux >= 0 && (uy % a.stride) == 0 && (ux % a.stride) == 0;
      iy = uy / a.stride;
      ix = ux / a.stride;
      v = v && iy < a.H && ix < a.W;
    }
    if (!__syncthreads_or(v)) continue;   // no pixel of this tile touches the tap (padding / parity)
    const T* ap = src + ((int64_t)(ln * a.H + iy) * a.W + ix) * a.in_pitch;
    const T* bp = w + ((int64_t)wrow * taps + tap) * a.w_pitch;
    for (int c0 = 0; c0 < a.C; c0 += BK) {
      const int c = c0 + lc;
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
      if (v) {
        if (c < a.C) a0 = icf::ldf(ap + c);
        if (c + 1 < a.C) a1 = icf::ldf(ap + c + 1);
      }
      if (bvalid) {
        if (c < a.C) b0 = icf::ldf(bp + c);
        if (c + 1 < a.C) b1 = icf::ldf(bp + c + 1);
      }
      As[lc][lrow] = a0;
      As[lc + 1][lrow] = a1;
      Bs[lc][lrow] = b0;
      Bs[lc + 1][lrow] = b1;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias + activation + Dropout2d mask (+ BatchNorm partial sums) -------------------
  TO* __restrict__ dst = reinterpret_cast<TO*>(a.dst);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int n = (int)(m / PQ);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = n0 + tx * 4 + j;
      if (k >= a.K) continue;
      float val = acc[i][j] + (a.bias ? a.bias[k] : 0.f);
      val = icf::apply_act(val, a.act, a.slope);
      if (a.out_mask) val *= a.out_mask[(int64_t)n * a.mask_pitch + k];
      TO* o = dst + m * a.out_pitch + k;
      if (a.accumulate) val += icf::ldf(o);
      icf::stf(o, val);
      const float stored = icf::ldf(o);   // statistics of what the consumer will read
      s1[j] += stored;
      s2[j] += stored * stored;
    }
  }
  if (a.stats) {
    if (tid < BN) { Ss[0][tid] = 0.f; Ss[1][tid] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&Ss[0][tx * 4 + j], s1[j]);
      atomicAdd(&Ss[1][tx * 4 + j], s2[j]);
    }
    __syncthreads();
    if (tid < BN && n0 + tid < a.K) {
      atomicAdd(a.stats + n0 + tid, Ss[0][tid]);
      atomicAdd(a.stats + a.K + n0 + tid, Ss[1][tid]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) wgrad_simt_kernel(const icf_wgrad_args a, int splits,
                                                        int64_t pix_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int taps = a.R * a.S;
  const int tap = blockIdx.z % taps;
  const int split = blockIdx.z / taps;
  const int r = tap / a.S, s = tap - r * a.S;
  const int PQ = a.P * a.Q;
  const int64_t M = (int64_t)a.N * PQ;
  const int64_t mb = (int64_t)split * pix_per_split;
  const int64_t me = min(M, mb + pix_per_split);
  const int a0 = blockIdx.x * BM, b0 = blockIdx.y * BN;
  const T* __restrict__ sm = reinterpret_cast<const T*>(a.small_t);
  const T* __restrict__ bg = reinterpret_cast<const T*>(a.big_t);

  const int lpix = tid >> 5;           // 0..7
  const int lch = (tid & 31) * 2;      // 0..62
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t mc = mb; mc < me; mc += BK) {
    const int64_t m = mc + lpix;
    float x0 = 0.f, x1 = 0.f, y0 = 0.f, y1 = 0.f;
    if (m < me) {
      const int n = (int)(m / PQ);
      const int rem = (int)(m - (int64_t)n * PQ);
      const int p = rem / a.Q, q = rem - p * a.Q;
      const T* sp = sm + m * a.a_pitch;
      if (a0 + lch < a.A) x0 = icf::ldf(sp + a0 + lch);
      if (a0 + lch + 1 < a.A) x1 = icf::ldf(sp + a0 + lch + 1);
      const int iy = p * a.stride - a.pad + r, ix = q * a.stride - a.pad + s;
      if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
        const T* bp = bg + ((int64_t)(n * a.H + iy) * a.W + ix) * a.b_pitch;
        if (b0 + lch < a.B) y0 = icf::ldf(bp + b0 + lch);
        if (b0 + lch + 1 < a.B) y1 = icf::ldf(bp + b0 + lch + 1);
      }
    }
    As[lpix][lch] = x0;
    As[lpix][lch + 1] = x1;
    Bs[lpix][lch] = y0;
    Bs[lpix][lch + 1] = y1;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ai = a0 + ty * 4 + i;
    if (ai >= a.A) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int bj = b0 + tx * 4 + j;
      if (bj >= a.B) continue;
      atomicAdd(a.dw + ((int64_t)ai * taps + tap) * a.B + bj, acc[i][j]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Weight gradient when `big` has ONE channel (the Cout = 1 last generator layer: small = X [N,P,Q,A],
// big = dY [N,H,W,1]):   dw[a][r*S+s] += sum_{n,p,q} X[n,p,q,a] * dY[n, p*stride-pad+r, q*stride-pad+s]
// A tensor-core tile would be 1/64 full; this is an HBM-bound streaming reduction instead: one block walks whole
// images, keeps the image's dY plane in shared memory (broadcast reads), lane = channel pair, warp = pixel.
// ------------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) wgrad_b1_kernel(const icf_wgrad_args a, int band_rows, int bands, int dy_rows_max) {
  extern __shared__ float sdy[];                   // [dy_rows_max * W] rows of the current band, then [8][KS*KS][64] scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = a.H * a.W, PQ = a.P * a.Q;
  const __nv_bfloat16* __restrict__ sm = reinterpret_cast<const __nv_bfloat16*>(a.small_t);
  const __nv_bfloat16* __restrict__ bg = reinterpret_cast<const __nv_bfloat16*>(a.big_t);
  const int items = a.N * bands;
  for (int a0 = 0; a0 < a.A; a0 += 64) {           // 64 channels per pass (one bf16 pair per lane)
    float acc[KS * KS][2];
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) acc[t][0] = acc[t][1] = 0.f;
    const int ch = a0 + 2 * lane;
    // work item = (image, band of rows of X): only the dY rows that band touches are staged (a 512 x 512 plane does not fit
    // shared memory; whole-image items also left a batch of 64 on 64 of the 148 SMs)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int n = item / bands, band = item - n * bands;
      const int p_lo = band * band_rows, p_hi = min(a.P, p_lo + band_rows);
      const int y_lo = max(0, p_lo * a.stride - a.pad), y_hi = min(a.H, (p_hi - 1) * a.stride - a.pad + KS);
      __syncthreads();
      for (int i = threadIdx.x; i < (y_hi - y_lo) * a.W; i += blockDim.x)
        sdy[i] = __bfloat162float(bg[((int64_t)n * HW + (int64_t)y_lo * a.W + i) * a.b_pitch]);
      __syncthreads();
      for (int pix = p_lo * a.Q + warp; pix < p_hi * a.Q; pix += 8) {
        const int p = pix / a.Q, q = pix - p * a.Q;
        float x0 = 0.f, x1 = 0.f;
        if (ch + 1 < a.A) {
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(sm + ((int64_t)n * PQ + pix) * a.a_pitch + ch);
          x0 = __bfloat162float(v.x);
          x1 = __bfloat162float(v.y);
        } else if (ch < a.A) {
          x0 = __bfloat162float(sm[((int64_t)n * PQ + pix) * a.a_pitch + ch]);
        }
        const int y0 = p * a.stride - a.pad, xx0 = q * a.stride - a.pad;
#pragma unroll
        for (int r = 0; r < KS; ++r) {
          const int y = y0 + r;
          const bool yin = y >= y_lo && y < y_hi;
#pragma unroll
          for (int c = 0; c < KS; ++c) {
            const int x = xx0 + c;
            const float v = (yin && x >= 0 && x < a.W) ? sdy[(y - y_lo) * a.W + x] : 0.f;
            acc[r * KS + c][0] = fmaf(x0, v, acc[r * KS + c][0]);
            acc[r * KS + c][1] = fmaf(x1, v, acc[r * KS + c][1]);
          }
        }
      }
    }
    // cross-warp reduction, then one atomic per (channel, tap) and block
    __syncthreads();
    float* red = sdy + (size_t)dy_rows_max * a.W;
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) {
      red[(warp * KS * KS + t) * 64 + 2 * lane] = acc[t][0];
      red[(warp * KS * KS + t) * 64 + 2 * lane + 1] = acc[t][1];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KS * KS * 64; i += blockDim.x) {
      const int t = i / 64, c = i - t * 64;
      if (a0 + c >= a.A) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[(w * KS * KS + t) * 64 + c];
      atomicAdd(a.dw + ((int64_t)(a0 + c) * KS * KS + t) * a.B, v);
    }
  }
}

template <int KS>
int launch_wgrad_b1(const icf_wgrad_args* a, cudaStream_t st) {
  // rows of X per work item: as many as keep the staged dY rows within ~48 KB, and enough items to fill the machine twice
  const int sms = icf::sm_count();
  int band_rows = a->P;
  auto dy_rows = [&](int br) { return (br - 1) * a->stride + KS; };
  while (band_rows > 1 && ((size_t)dy_rows(band_rows) * a->W * sizeof(float) > 48 * 1024 ||
                           (int64_t)a->N * icf::cdiv(a->P, band_rows) < 2 * sms))
    band_rows = (band_rows + 1) / 2;
  const int bands = icf::cdiv(a->P, band_rows);
  int dy_rows_max = dy_rows(band_rows);
  if (dy_rows_max > a->H) dy_rows_max = a->H;
  const size_t smem = ((size_t)dy_rows_max * a->W + 8 * KS * KS * 64) * sizeof(float);
  if (smem > 200 * 1024) return -1;
  static icf::SmemGuard guard;
  if (smem > 48 * 1024 && guard.ensure(reinterpret_cast<const void*>(wgrad_b1_kernel<KS>), smem, "single-channel wgrad")) return -1;
  const int64_t items = (int64_t)a->N * bands;
  const int cap = sms * 2;
  const int grid = (int)(items < cap ? items : cap);
  wgrad_b1_kernel<KS><<<grid, 256, smem, st>>>(*a, band_rows, bands, dy_rows_max);
  return icf::check_launch("wgrad_b1");
}

template <typename T, int FORM>
int launch_conv(const icf_conv_args* a, cudaStream_t st) {
  const int64_t M = (int64_t)a->N * a->P * a->Q;
  dim3 grid(icf::cdiv(M, BM), icf::cdiv(a->K, BN));
  const bool out_f32 = a->out_f32 || a->dtype == ICF_F32;
  if (out_f32) conv_simt_kernel<T, FORM, float><<<grid, NT, 0, st>>>(*a);
  else conv_simt_kernel<T, FORM, T><<<grid, NT, 0, st>>>(*a);
  return icf::check_launch("conv_simt");
}

}  // namespace

int icf_simt_conv_forward(const icf_conv_args* a, cudaStream_t st) {
  if (a->dtype == ICF_F32) {
    return a->form == ICF_FORM_GATHER ? launch_conv<float, ICF_FORM_GATHER>(a, st)
                                      : launch_conv<float, ICF_FORM_TRANSPOSED>(a, st);
  }
  return a->form == ICF_FORM_GATHER ? launch_conv<__nv_bfloat16, ICF_FORM_GATHER>(a, st)
                                    : launch_conv<__nv_bfloat16, ICF_FORM_TRANSPOSED>(a, st);
}

int icf_simt_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st) {
  const int taps = a->R * a->S;
  const int64_t M = (int64_t)a->N * a->P * a->Q;
  const int gx = icf::cdiv(a->A, BM), gy = icf::cdiv(a->B, BN);
  // enough splits over the pixel dimension to fill the SMs a few times
  int64_t want = ((int64_t)icf::sm_count() * 4 + (int64_t)gx * gy * taps - 1) / ((int64_t)gx * gy * taps);
  int64_t max_splits = (M + 255) / 256;
  int splits = (int)(want < 1 ? 1 : (want > max_splits ? max_splits : want));
  if (splits < 1) splits = 1;
  if ((int64_t)taps * splits > 65535) splits = 65535 / taps;
  int64_t pps = (M + splits - 1) / splits;
  pps = (pps + BK - 1) / BK * BK;
  dim3 grid(gx, gy, taps * splits);
  if (a->dtype == ICF_F32) wgrad_simt_kernel<float><<<grid, NT, 0, st>>>(*a, splits, pps);
  else wgrad_simt_kernel<__nv_bfloat16><<<grid, NT, 0, st>>>(*a, splits, pps);
  return icf::check_launch("wgrad_simt");
}

// single-channel `big` operand (bf16): streaming reduction kernel; -1 = not applicable
int icf_b1_conv_wgrad(const icf_wgrad_args* a, cudaStream_t st) {
  if (a->dtype != ICF_BF16 || a->B != 1 || a->win > 1 || a->R != a->S || (a->a_pitch & 1)) return -1;
  if ((reinterpret_cast<uintptr_t>(a->small_t) & 3) != 0) return -1;
  switch (a->R) {
    case 3: return launch_wgrad_b1<3>(a, st);
    case 4: return launch_wgrad_b1<4>(a, st);
    case 5: return launch_wgrad_b1<5>(a, st);
    default: return -1;
  }
}
