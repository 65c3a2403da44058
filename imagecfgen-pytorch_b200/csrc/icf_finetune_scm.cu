// Kernels of the two callers either side of the hot path, all HBM-streaming:
//   * fine-tune losses (finetune_mnist_bigan.py:68-86, finetune_whale_bigan.py:54-73): reconstruction MSE forward +
//     gradient in one pass over the reconstruction, latent penalty mean(z^2) folded into the latent gradient;
//   * attribute-SCM intervention (attribute_scms/graph.py:144-184): abduction + regeneration of a conditional
//     affine -> sigmoid -> affine mechanism (the thickness -> intensity mechanism of attribute_scms/mnist.py:28-33,48),
//     fused with the min-max rescale of mnist_gan_counterfactuals.py:57-68, and the index -> one-hot / masked swap of
//     mnist_bigan_score.py:83-91.
#include "icf_common.cuh"

namespace {

constexpr int LT = 256;

// loss_out[0] += weight * mean((x - xr)^2);  dxr = weight * 2 (xr - x) / count.   x: fp32 [N][P] (target_stride = P) or one
// image [P] broadcast over the batch (target_stride = 0); xr: [N*P][xr_pitch] (channel 0), dxr: [N*P][d_pitch] (channel 0)
__global__ void __launch_bounds__(LT) mse_loss_kernel(const float* __restrict__ x, int64_t target_stride, const void* xr, int xr_dtype,
                                                      int xr_pitch, int64_t n_img, int64_t P, float weight, const float* extra, float* loss_out,
                                                      void* dxr, int d_dtype, int d_pitch) {
  __shared__ float red[32];
  const int64_t total = n_img * P;
  const float inv = 1.f / (float)total;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < total; i += (int64_t)gridDim.x * LT) {
    const int64_t n = i / P, pp = i - n * P;
    const float t = x[n * target_stride + pp];
    const float r = icf::ld_any(xr, xr_dtype, i * xr_pitch);
    const float d = r - t;
    s = fmaf(d, d, s);
    if (dxr) icf::st_any(dxr, d_dtype, i * d_pitch, 2.f * weight * inv * d);
  }
  const float t = icf::block_sum(s, red);
  if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, weight * (t * inv + ((blockIdx.x == 0 && extra) ? extra[0] : 0.f)));
}

// column mean of an [N][P] fp32 matrix and the mean over columns of the (biased) column variance:
//   xbar[p] = mean_n x[n][p];   var_out[0] += mean_p( mean_n x[n][p]^2 - xbar[p]^2 )
__global__ void __launch_bounds__(LT) col_mean_kernel(const float* __restrict__ x, int64_t N, int64_t P, float* xbar, float* var_out) {
  __shared__ float red[32];
  float v = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * LT + threadIdx.x; p < P; p += (int64_t)gridDim.x * LT) {
    float s = 0.f, q = 0.f;
    for (int64_t n = 0; n < N; ++n) {
      const float t = x[n * P + p];
      s += t;
      q = fmaf(t, t, q);
    }
    const float m = s / (float)N;
    xbar[p] = m;
    v += q / (float)N - m * m;
  }
  const float t = icf::block_sum(v, red);
  if (threadIdx.x == 0 && var_out) atomicAdd(var_out, t / (float)P);
}

// loss_out[0] += weight * mean(z^2);  dz[n][j] (+)= weight * 2 z / count   (z: [N][z_pitch], dz: fp32 [N][latent])
__global__ void __launch_bounds__(LT) latent_l2_kernel(const void* z, int z_dtype, int z_pitch, int64_t N, int latent, float weight,
                                                       float* loss_out, float* dz, int accumulate) {
  __shared__ float red[32];
  const int64_t total = N * latent;
  const float inv = 1.f / (float)total;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < total; i += (int64_t)gridDim.x * LT) {
    const int64_t n = i / latent;
    const int j = (int)(i - n * latent);
    const float v = icf::ld_any(z, z_dtype, n * z_pitch + j);
    s = fmaf(v, v, s);
    if (dz) dz[i] = (accumulate ? dz[i] : 0.f) + 2.f * weight * inv * v;
  }
  const float t = icf::block_sum(s, red);
  if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, weight * t * inv);
}

// One conditional affine -> sigmoid -> affine mechanism, per sample:
//   abduction      u = (v - lo) / span;  s = logit(u);  eps = (s - loc(p)) / scale(p)
//   regeneration   s' = loc(p') + scale(p') * eps;      v' = lo + span * sigmoid(s')
// (loc, log scale) = hyper-network of the parent value: out = W2 relu(W1 p + b1) + b2, hidden width H (pyro's
// ConditionalAutoRegressiveNN for a 1-D variable with a 1-D context; H = 0: the closed form loc = a0 + a1 p,
// log scale = a2 of the ground-truth SCM, create_train_dataset.py:42-46), log scale clamped to [clip_lo, clip_hi].
// The counterfactual value and both parents are also written min-max scaled to [-1, 1] (mnist.py:205-209) when asked.
__global__ void __launch_bounds__(LT) scm_affine_cf_kernel(const icf_scm_affine_args a) {
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * LT) {
    const float v = a.value[i], p = a.parent[i];
    const float pc = a.parent_cf ? a.parent_cf[i] : p + a.parent_shift;
    float loc[2], ls[2];
    const float ctx[2] = {p, pc};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (a.hidden > 0) {
        float o0 = a.b2[0], o1 = a.b2[1];
        for (int h = 0; h < a.hidden; ++h) {
          const float t = fmaxf(fmaf(a.w1[h], ctx[k], a.b1[h]), 0.f);
          o0 = fmaf(a.w2[h], t, o0);
          o1 = fmaf(a.w2[a.hidden + h], t, o1);
        }
        loc[k] = o0;
        ls[k] = o1;
      } else {
        loc[k] = fmaf(a.closed[1], ctx[k], a.closed[0]);
        ls[k] = a.closed[2];
      }
      ls[k] = fminf(fmaxf(ls[k], a.clip_lo), a.clip_hi);
    }
    float u = (v - a.lo) / a.span;
    u = fminf(fmaxf(u, a.u_min), a.u_max);
    const float s = logf(u) - log1pf(-u);
    const float eps = (s - loc[0]) * expf(-ls[0]);
    const float s2 = fmaf(expf(ls[1]), eps, loc[1]);
    const float vcf = a.lo + a.span / (1.f + expf(-s2));
    if (a.noise_out) a.noise_out[i] = eps;
    if (a.value_cf) a.value_cf[i] = vcf;
    if (a.parent_cf_out) a.parent_cf_out[i] = pc;
    if (a.value_cf_scaled) a.value_cf_scaled[i] = 2.f * (vcf - a.v_min) / (a.v_max - a.v_min) - 1.f;
    if (a.parent_cf_scaled) a.parent_cf_scaled[i] = 2.f * (pc - a.p_min) / (a.p_max - a.p_min) - 1.f;
  }
}

// rows[n][:] = one_hot(idx_new[n]) where mask[n] != 0 (or mask == NULL), else the existing row stays (torch.eye(K)[idx] and
// the masked swap of mnist_bigan_score.py:83-91, audiomnist_cf_eval.py:82-83); index dtype int32 or int64
__global__ void __launch_bounds__(LT) onehot_swap_kernel(const void* idx, int idx64, const uint8_t* mask, int64_t n, int K, float* rows) {
  const int64_t total = n * K;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < total; i += (int64_t)gridDim.x * LT) {
    const int64_t r = i / K;
    const int k = (int)(i - r * K);
    if (mask && !mask[r]) continue;
    const int64_t j = idx64 ? reinterpret_cast<const int64_t*>(idx)[r] : (int64_t)reinterpret_cast<const int32_t*>(idx)[r];
    rows[i] = (j == k) ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// "Taps as channels": a stride-s transposed convolution with very few output channels (the generator's single-channel
// tail ConvTranspose2d(C, 1, 5, 2, 2, 1), the data gradient of the first encoder / discriminator conv towards the
// attribute-plane channels) is a plain GEMM over the INPUT pixels,  T[pixel][(k, tap)] = sum_c in[pixel][c] * w[k][tap][c]
// (N = K*R*S columns on the tensor cores, every input element read once), followed by a col2im gather
//   out[n][oy][ox][k] = act(bias[k] + sum_{taps whose (oy+pad-r, ox+pad-s) is a multiple of the stride} T[...]) .
// The gather form of the same layers (one MMA chain per output-parity class and tap with N = 16 for 1-3 useful columns)
// is issue-bound at 4-6 % of the HBM roofline (profiles/r02_per_layer_esrf_acoustic.json).
// The mirror image for a conv whose INPUT has one channel (data / weight gradient of that tail): im2col of the single
// plane into [pixel][tap] rows, then plain GEMMs.
// ------------------------------------------------------------------------------------------------------------------
// col2im: T [N*H*W][t_pitch] (columns k*TP + tap, TP = tap stride) -> out [N*P*Q][out_pitch] channels 0..K-1
// one thread per output pixel (32-bit index arithmetic), at most 4 filter rows / columns reach a pixel
template <int KMAX>
__global__ void __launch_bounds__(LT) col2im_taps_kernel(const __nv_bfloat16* __restrict__ T, int t_pitch, int TP, int N, int H, int W,
                                                         int P, int Q, int K, int R, int S, int stride, int pad,
                                                         const float* bias, int act, float slope, void* out, int out_dtype,
                                                         int out_pitch) {
  const uint32_t rows = (uint32_t)N * (uint32_t)P;
  const uint64_t total = (uint64_t)rows * (uint32_t)Q;
  for (uint64_t o = (uint64_t)blockIdx.x * LT + threadIdx.x; o < total; o += (uint64_t)gridDim.x * LT) {
    const uint32_t row = (uint32_t)(o / (uint32_t)Q);
    const int ox = (int)(o - (uint64_t)row * (uint32_t)Q);
    const int n = (int)(row / (uint32_t)P), oy = (int)(row - (uint32_t)n * (uint32_t)P);
    float acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] = (bias && k < K) ? bias[k] : 0.f;
    for (int r = (oy + pad) % stride; r < R; r += stride) {
      const int dy = oy + pad - r;
      if (dy < 0) break;                            // larger r only moves further up
      const int iy = dy / stride;
      if (iy >= H) continue;
      for (int sx = (ox + pad) % stride; sx < S; sx += stride) {
        const int dx = ox + pad - sx;
        if (dx < 0) break;
        const int ix = dx / stride;
        if (ix >= W) continue;
        const __nv_bfloat16* t = T + (((int64_t)n * H + iy) * W + ix) * t_pitch + r * S + sx;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) acc[k] += __bfloat162float(t[k * TP]);
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) icf::st_any(out, out_dtype, (int64_t)o * out_pitch + k, icf::apply_act(acc[k], act, slope));
  }
}

// im2col of channel 0 of src [N*P*Q][src_pitch] into A [N*H*W][a_pitch]: A[n,iy,ix][r*S+s] = src[n, iy*stride-pad+r, ix*stride-pad+s]
// (zero outside, zero in the padding columns R*S .. a_pitch-1); one thread per (pixel, 8-column group) -> 16-byte stores
__global__ void __launch_bounds__(LT) im2col_taps_kernel(const void* src, int src_dtype, int src_pitch, int N, int P, int Q, int H, int W,
                                                         int R, int S, int stride, int pad, __nv_bfloat16* A, int a_pitch) {
  const int groups = a_pitch >> 3;
  const int64_t total = (int64_t)N * H * W * groups;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < total; i += (int64_t)gridDim.x * LT) {
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int ix = (int)(pix % W);
    const int64_t t1 = pix / W;
    const int iy = (int)(t1 % H);
    const int n = (int)(t1 / H);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = g * 8 + j + h;
        v[h] = 0.f;
        if (t < R * S) {
          const int r = t / S, sx = t - r * S;
          const int y = iy * stride - pad + r, x = ix * stride - pad + sx;
          if (y >= 0 && y < P && x >= 0 && x < Q) v[h] = icf::ld_any(src, src_dtype, (((int64_t)n * P + y) * Q + x) * src_pitch);
        }
      }
      __nv_bfloat162 pk = __floats2bfloat162_rn(v[0], v[1]);
      w[j >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(A + pix * a_pitch + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// BatchNorm folded into the convolution that consumes it (no Dropout2d in between, no padding):
//   conv(scale*y + shift; w, b) = conv(y; w * scale[c], b + sum_{t,c} w[k][t][c] * shift[c])
// one block per output channel k of the packed operand [K][T][Cp]
__global__ void __launch_bounds__(LT) bn_fold_weights_kernel(const void* w, int dtype, int T, int Cp, int C, const float* scale,
                                                             const float* shift, const float* bias, void* w_out, float* bias_out) {
  __shared__ float red[32];
  const int k = blockIdx.x;
  const int64_t base = (int64_t)k * T * Cp;
  float s = 0.f;
  for (int i = threadIdx.x; i < T * Cp; i += LT) {
    const int c = i % Cp;
    float v = 0.f;
    if (c < C) {
      const float wv = icf::ld_any(w, dtype, base + i);
      v = wv * scale[c];
      s = fmaf(wv, shift[c], s);
    }
    icf::st_any(w_out, dtype, base + i, v);
  }
  const float t = icf::block_sum(s, red);
  if (threadIdx.x == 0) bias_out[k] = (bias ? bias[k] : 0.f) + t;
}

// weight gradient of that convolution from the gradient computed against y:  dW[k][t][c] = G[k][t][c]*scale[c] + shift[c]*dbias[k]
__global__ void __launch_bounds__(LT) bn_fold_wgrad_kernel(float* dw, int64_t total, int TC, int C, const float* scale,
                                                           const float* shift, const float* dbias) {
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < total; i += (int64_t)gridDim.x * LT) {
    const int c = (int)(i % C);
    const int64_t k = i / TC;
    dw[i] = fmaf(dw[i], scale[c], shift[c] * dbias[k]);
  }
}

inline int grid_for(int64_t total) {
  int64_t b = (total + LT - 1) / LT;
  const int64_t cap = (int64_t)icf::sm_count() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}


// ------------------------------------------------------------------------------------------------
// Gradient-based counterfactual explainer (explain/cf_example.py:130-160): the optimised variables are raw rows; the generator
// sees tanh(raw) (latent code, continuous attributes) or softmax(raw) (categorical attributes).  One warp per (group, row).
// ------------------------------------------------------------------------------------------------
struct ExplainGroups { icf_explain_group g[ICF_EXPLAIN_MAX_GROUPS]; int n; };

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void explain_transform_kernel(const float* __restrict__ raw, float* __restrict__ out, ExplainGroups gs, int64_t rows) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * gs.n) return;
  const int gi = (int)(w / rows);
  const int64_t r = w - (int64_t)gi * rows;
  const icf_explain_group g = gs.g[gi];
  const float* x = raw + g.offset + r * g.width;
  float* y = out + g.offset + r * g.width;
  if (g.mode == ICF_EXPLAIN_SOFTMAX) {
    float mx = -INFINITY;
    for (int j = lane; j < g.width; j += 32) mx = fmaxf(mx, x[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < g.width; j += 32) sum += expf(x[j] - mx);
    sum = warp_sum(sum);
    for (int j = lane; j < g.width; j += 32) y[j] = expf(x[j] - mx) / sum;
  } else if (g.mode == ICF_EXPLAIN_TANH) {
    for (int j = lane; j < g.width; j += 32) y[j] = tanhf(x[j]);
  } else {
    for (int j = lane; j < g.width; j += 32) y[j] = x[j];
  }
}

// draw = dout * d(out)/d(raw) through the saved outputs: tanh' = 1 - y^2;  softmax: y * (g - sum_j g_j y_j)
__global__ void explain_backward_kernel(const float* __restrict__ out, float* __restrict__ draw, ExplainGroups gs, int64_t rows) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * gs.n) return;
  const int gi = (int)(w / rows);
  const int64_t r = w - (int64_t)gi * rows;
  const icf_explain_group g = gs.g[gi];
  const float* y = out + g.offset + r * g.width;
  float* dx = draw + g.offset + r * g.width;
  const float* dy = g.dout ? g.dout + r * g.width : nullptr;
  if (!dy) {
    for (int j = lane; j < g.width; j += 32) dx[j] = 0.f;
  } else if (g.mode == ICF_EXPLAIN_SOFTMAX) {
    float dot = 0.f;
    for (int j = lane; j < g.width; j += 32) dot = fmaf(dy[j], y[j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < g.width; j += 32) dx[j] = y[j] * (dy[j] - dot);
  } else if (g.mode == ICF_EXPLAIN_TANH) {
    for (int j = lane; j < g.width; j += 32) dx[j] = dy[j] * (1.f - y[j] * y[j]);
  } else {
    for (int j = lane; j < g.width; j += 32) dx[j] = dy[j];
  }
}

static int explain_groups(const icf_explain_group* groups, int32_t n_groups, ExplainGroups* gs) {
  if (!groups || n_groups < 0 || n_groups > ICF_EXPLAIN_MAX_GROUPS) return 1;
  gs->n = n_groups;
  for (int i = 0; i < n_groups; ++i) {
    if (groups[i].width <= 0 || groups[i].offset < 0) return 1;
    gs->g[i] = groups[i];
  }
  return 0;
}

}  // namespace

extern "C" {

int icf_mse_loss(const float* x, int64_t target_stride, const void* xr, int32_t xr_dtype, int32_t xr_pitch, int64_t n_img,
                 int64_t pixels_per_image, float weight, const float* extra, float* loss_out, void* dxr, int32_t d_dtype, int32_t d_pitch,
                 void* stream) {
  ICF_REQUIRE(x && xr && n_img > 0 && pixels_per_image > 0 && xr_pitch > 0 && (!dxr || d_pitch > 0) &&
                  (target_stride == 0 || target_stride == pixels_per_image),
              "icf_mse_loss: bad arguments");
  mse_loss_kernel<<<grid_for(n_img * pixels_per_image), LT, 0, icf::as_stream(stream)>>>(
      x, target_stride, xr, xr_dtype, xr_pitch, n_img, pixels_per_image, weight, extra, loss_out, dxr, d_dtype, d_pitch);
  return icf::check_launch("mse_loss");
}

int icf_col_mean(const float* x, int64_t n, int64_t p, float* xbar, float* var_out, void* stream) {
  ICF_REQUIRE(x && xbar && n > 0 && p > 0, "icf_col_mean: bad arguments");
  col_mean_kernel<<<grid_for(p), LT, 0, icf::as_stream(stream)>>>(x, n, p, xbar, var_out);
  return icf::check_launch("col_mean");
}

int icf_latent_l2(const void* z, int32_t z_dtype, int32_t z_pitch, int64_t n, int32_t latent, float weight, float* loss_out,
                  float* dz, int32_t accumulate, void* stream) {
  ICF_REQUIRE(z && n > 0 && latent > 0 && z_pitch >= latent, "icf_latent_l2: bad arguments");
  latent_l2_kernel<<<grid_for(n * latent), LT, 0, icf::as_stream(stream)>>>(z, z_dtype, z_pitch, n, latent, weight, loss_out, dz,
                                                                              accumulate);
  return icf::check_launch("latent_l2");
}

int icf_scm_affine_cf(const icf_scm_affine_args* a, void* stream) {
  ICF_REQUIRE(a && a->value && a->parent && a->n >= 0 && a->hidden >= 0 && a->span != 0.f, "icf_scm_affine_cf: bad arguments");
  ICF_REQUIRE(a->hidden == 0 || (a->w1 && a->b1 && a->w2 && a->b2), "icf_scm_affine_cf: hyper-network weights missing");
  if (a->n == 0) return 0;
  scm_affine_cf_kernel<<<grid_for(a->n), LT, 0, icf::as_stream(stream)>>>(*a);
  return icf::check_launch("scm_affine_cf");
}

int icf_col2im_taps(const void* T, int32_t t_pitch, int32_t TP, int32_t N, int32_t H, int32_t W, int32_t P, int32_t Q, int32_t K,
                    int32_t R, int32_t S, int32_t stride, int32_t pad, const float* bias, int32_t act, float slope, void* out,
                    int32_t out_dtype, int32_t out_pitch, void* stream) {
  ICF_REQUIRE(T && out && N >= 0 && H > 0 && W > 0 && P > 0 && Q > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0 &&
                  TP >= R * S && t_pitch >= K * TP && out_pitch >= K,
              "icf_col2im_taps: bad arguments");
  if (N == 0) return 0;
  ICF_REQUIRE(K <= 8 && (int64_t)N * P < 0x7fffffffLL, "icf_col2im_taps: at most 8 channels");
  const int grid = grid_for((int64_t)N * P * Q);
  const __nv_bfloat16* Tb = reinterpret_cast<const __nv_bfloat16*>(T);
  if (K == 1)
    col2im_taps_kernel<1><<<grid, LT, 0, icf::as_stream(stream)>>>(Tb, t_pitch, TP, N, H, W, P, Q, K, R, S, stride, pad, bias, act,
                                                                   slope, out, out_dtype, out_pitch);
  else if (K <= 2)
    col2im_taps_kernel<2><<<grid, LT, 0, icf::as_stream(stream)>>>(Tb, t_pitch, TP, N, H, W, P, Q, K, R, S, stride, pad, bias, act,
                                                                   slope, out, out_dtype, out_pitch);
  else
    col2im_taps_kernel<8><<<grid, LT, 0, icf::as_stream(stream)>>>(Tb, t_pitch, TP, N, H, W, P, Q, K, R, S, stride, pad, bias, act,
                                                                   slope, out, out_dtype, out_pitch);
  return icf::check_launch("col2im_taps");
}

int icf_im2col_taps(const void* src, int32_t src_dtype, int32_t src_pitch, int32_t N, int32_t P, int32_t Q, int32_t H, int32_t W,
                    int32_t R, int32_t S, int32_t stride, int32_t pad, void* A, int32_t a_pitch, void* stream) {
  ICF_REQUIRE(src && A && N >= 0 && P > 0 && Q > 0 && H > 0 && W > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0 &&
                  a_pitch >= R * S && (a_pitch & 7) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0,
              "icf_im2col_taps: bad arguments");
  if (N == 0) return 0;
  im2col_taps_kernel<<<grid_for((int64_t)N * H * W * (a_pitch >> 3)), LT, 0, icf::as_stream(stream)>>>(
      src, src_dtype, src_pitch, N, P, Q, H, W, R, S, stride, pad, reinterpret_cast<__nv_bfloat16*>(A), a_pitch);
  return icf::check_launch("im2col_taps");
}

int icf_bn_fold_weights(const void* w, int32_t dtype, int32_t K, int32_t T, int32_t Cp, int32_t C, const float* scale,
                        const float* shift, const float* bias, void* w_out, float* bias_out, void* stream) {
  ICF_REQUIRE(w && w_out && bias_out && scale && shift && K > 0 && T > 0 && Cp >= C && C > 0, "icf_bn_fold_weights: bad arguments");
  bn_fold_weights_kernel<<<K, LT, 0, icf::as_stream(stream)>>>(w, dtype, T, Cp, C, scale, shift, bias, w_out, bias_out);
  return icf::check_launch("bn_fold_weights");
}

int icf_bn_fold_wgrad(float* dw, int32_t K, int32_t T, int32_t C, const float* scale, const float* shift, const float* dbias,
                      void* stream) {
  ICF_REQUIRE(dw && scale && shift && dbias && K > 0 && T > 0 && C > 0, "icf_bn_fold_wgrad: bad arguments");
  const int64_t total = (int64_t)K * T * C;
  bn_fold_wgrad_kernel<<<grid_for(total), LT, 0, icf::as_stream(stream)>>>(dw, total, T * C, C, scale, shift, dbias);
  return icf::check_launch("bn_fold_wgrad");
}

int icf_onehot_swap(const void* idx, int32_t idx_is_int64, const uint8_t* mask, int64_t n, int32_t K, float* rows, void* stream) {
  ICF_REQUIRE(idx && rows && n >= 0 && K > 0, "icf_onehot_swap: bad arguments");
  if (n == 0) return 0;
  onehot_swap_kernel<<<grid_for(n * K), LT, 0, icf::as_stream(stream)>>>(idx, idx_is_int64, mask, n, K, rows);
  return icf::check_launch("onehot_swap");
}


int icf_explain_transform(const float* raw, float* out, const icf_explain_group* groups, int32_t n_groups, int64_t rows,
                          void* stream) {
  ExplainGroups gs;
  ICF_REQUIRE(raw && out && rows >= 0 && explain_groups(groups, n_groups, &gs) == 0, "icf_explain_transform: bad arguments");
  if (rows == 0 || n_groups == 0) return 0;
  const int64_t threads = rows * n_groups * 32;
  explain_transform_kernel<<<(unsigned)icf::cdiv(threads, (int64_t)128), 128, 0, icf::as_stream(stream)>>>(raw, out, gs, rows);
  return icf::check_launch("explain_transform");
}

int icf_explain_backward(const float* out, float* draw, const icf_explain_group* groups, int32_t n_groups, int64_t rows,
                         void* stream) {
  ExplainGroups gs;
  ICF_REQUIRE(out && draw && rows >= 0 && explain_groups(groups, n_groups, &gs) == 0, "icf_explain_backward: bad arguments");
  if (rows == 0 || n_groups == 0) return 0;
  const int64_t threads = rows * n_groups * 32;
  explain_backward_kernel<<<(unsigned)icf::cdiv(threads, (int64_t)128), 128, 0, icf::as_stream(stream)>>>(out, draw, gs, rows);
  return icf::check_launch("explain_backward");
}

}  // extern "C"
