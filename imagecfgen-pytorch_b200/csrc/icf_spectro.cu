// Spectrogram front end of the spectrogram families: the step between the waveform and the BiGAN hot path.
//   log-power STFT   audio_mnist.py:59-61,116: torchaudio.transforms.Spectrogram(n_fft=255, win_length=128, pad=96) then
//                    (. + 1e-6).log() — zero padding, reflect centring, periodic Hann window centred in the FFT frame,
//                    one-sided power spectrum; all in ONE pass over the waveform (a direct DFT: 128 window samples x 128 bins per
//                    frame; the frames of a clip share its samples in shared memory)
//   dataset statistics   audio_mnist.py:347-358: per-time-frame mean and mean of squares over (clip, frequency)
//   spect_to_img     audio_mnist.py:361-363: clip((s - mean[t]) / (std[t] + 1e-6), -k, k) / k, written in the engine's dtype
#include "icf_common.cuh"

namespace {

constexpr int ST = 256;
constexpr int FT = 32;       // frames per block

// block = (clip n, tile of FT frames); thread = frequency bin (two halves of the frame tile when ST = 2 * bins)
__global__ void __launch_bounds__(ST) log_spectrogram_kernel(const float* __restrict__ wave, int L, int n_fft, int win, int hop, int pad,
                                                             float eps, float* __restrict__ out, int frames, int bins) {
  extern __shared__ float sm[];
  const int c = n_fft / 2, left = (n_fft - win) / 2;
  const int Lp = L + 2 * pad;
  const int f0 = blockIdx.x * FT, n = blockIdx.y;
  const int nf = min(FT, frames - f0);
  const int seg = (nf - 1) * hop + win;          // windowed samples the tile touches
  float* tw_c = sm;                              // [n_fft] cos(2 pi m / n_fft)
  float* tw_s = tw_c + n_fft;                    // [n_fft] sin
  float* hann = tw_s + n_fft;                    // [win]
  float* xs = hann + win;                        // [seg] reflect-centred, zero-padded samples
  float* ot = xs + ((FT - 1) * hop + win);       // [bins][FT + 1] output tile
  for (int m = threadIdx.x; m < n_fft; m += ST) sincospif(2.f * (float)m / (float)n_fft, &tw_s[m], &tw_c[m]);
  for (int j = threadIdx.x; j < win; j += ST) hann[j] = 0.5f - 0.5f * cospif(2.f * (float)j / (float)win);
  // frame f, window sample j reads the centred signal at  f*hop + left + j - c  (reflect about 0 and Lp-1, zeros in the padding)
  for (int i = threadIdx.x; i < seg; i += ST) {
    int t = f0 * hop + left + i - c;
    if (t < 0) t = -t;
    if (t >= Lp) t = 2 * (Lp - 1) - t;
    const int u = t - pad;
    xs[i] = (u >= 0 && u < L) ? wave[(int64_t)n * L + u] : 0.f;
  }
  __syncthreads();
  const int halves = ST / bins > 0 ? ST / bins : 1;          // 2 for 128 bins
  const int k = threadIdx.x % bins, h = threadIdx.x / bins;
  if (h < halves) {
    for (int f = h; f < nf; f += halves) {
      float re = 0.f, im = 0.f;
      int idx = (int)(((int64_t)k * left) % n_fft);
      const float* x = xs + f * hop;
      for (int j = 0; j < win; ++j) {
        const float v = x[j] * hann[j];
        re = fmaf(v, tw_c[idx], re);
        im = fmaf(-v, tw_s[idx], im);
        idx += k;
        if (idx >= n_fft) idx -= n_fft;
      }
      ot[k * (FT + 1) + f] = logf(re * re + im * im + eps);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins * nf; i += ST) {
    const int kk = i / nf, f = i - kk * nf;
    out[((int64_t)n * bins + kk) * frames + f0 + f] = ot[kk * (FT + 1) + f];
  }
}

// sum[t] += sum over rows of s[row][t], sumsq[t] += sum of squares  (rows = clips x frequency bins)
__global__ void __launch_bounds__(ST) spect_stats_kernel(const float* __restrict__ s, int64_t rows, int T, float* sum, float* sumsq) {
  const int t = blockIdx.x * ST + threadIdx.x;
  if (t >= T) return;
  float a = 0.f, b = 0.f;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float v = s[r * T + t];
    a += v;
    b = fmaf(v, v, b);
  }
  atomicAdd(sum + t, a);
  atomicAdd(sumsq + t, b);
}

__global__ void __launch_bounds__(ST) spect_to_img_kernel(const float* __restrict__ s, const float* __restrict__ mean,
                                                          const float* __restrict__ sd, int64_t total, int T, float k, void* out,
                                                          int out_dtype) {
  for (int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x; i < total; i += (int64_t)gridDim.x * ST) {
    const int t = (int)(i % T);
    float v = (s[i] - mean[t]) / (sd[t] + 1e-6f);
    v = fminf(fmaxf(v, -k), k) / k;
    icf::st_any(out, out_dtype, i, v);
  }
}

}  // namespace

extern "C" {

int icf_log_spectrogram(const float* wave, int64_t N, int32_t L, int32_t n_fft, int32_t win_length, int32_t hop, int32_t pad, float eps,
                        float* out, int32_t frames, void* stream) {
  ICF_REQUIRE(wave && out && N >= 0 && L > 0 && n_fft > 0 && win_length > 0 && win_length <= n_fft && hop > 0 && pad >= 0 && frames > 0,
              "icf_log_spectrogram: bad arguments");
  const int bins = n_fft / 2 + 1;
  ICF_REQUIRE(bins <= ST && frames == 1 + (L + 2 * pad + 2 * (n_fft / 2) - n_fft) / hop && n_fft / 2 < L + 2 * pad,
              "icf_log_spectrogram: %d bins (max %d) / frame count %d does not match torch.stft(center=True)", bins, ST, frames);
  if (N == 0) return 0;
  ICF_REQUIRE(N <= 65535, "icf_log_spectrogram: at most 65535 clips per call");
  const size_t smem = (size_t)(2 * n_fft + win_length + (FT - 1) * hop + win_length + bins * (FT + 1)) * sizeof(float);
  static icf::SmemGuard guard;
  if (smem > 48 * 1024)
    if (int r = guard.ensure(reinterpret_cast<const void*>(log_spectrogram_kernel), smem, "log spectrogram")) return r;
  dim3 grid((unsigned)((frames + FT - 1) / FT), (unsigned)N);
  log_spectrogram_kernel<<<grid, ST, smem, icf::as_stream(stream)>>>(wave, L, n_fft, win_length, hop, pad, eps, out, frames, bins);
  return icf::check_launch("log_spectrogram");
}

int icf_spect_stats(const float* s, int64_t rows, int32_t T, float* sum, float* sumsq, void* stream) {
  ICF_REQUIRE(s && sum && sumsq && rows >= 0 && T > 0, "icf_spect_stats: bad arguments");
  if (rows == 0) return 0;
  int64_t gy = rows < 64 ? rows : 64;
  dim3 grid((unsigned)((T + ST - 1) / ST), (unsigned)gy);
  spect_stats_kernel<<<grid, ST, 0, icf::as_stream(stream)>>>(s, rows, T, sum, sumsq);
  return icf::check_launch("spect_stats");
}

int icf_spect_to_img(const float* s, const float* mean, const float* std_, int64_t rows, int32_t T, float stds_kept, void* out,
                     int32_t out_dtype, void* stream) {
  ICF_REQUIRE(s && mean && std_ && out && rows >= 0 && T > 0 && stds_kept > 0.f, "icf_spect_to_img: bad arguments");
  const int64_t total = rows * T;
  if (total == 0) return 0;
  int64_t blocks = (total + ST - 1) / ST;
  const int64_t cap = (int64_t)icf::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  spect_to_img_kernel<<<(unsigned)blocks, ST, 0, icf::as_stream(stream)>>>(s, mean, std_, total, T, stds_kept, out, out_dtype);
  return icf::check_launch("spect_to_img");
}

}  // extern "C"
