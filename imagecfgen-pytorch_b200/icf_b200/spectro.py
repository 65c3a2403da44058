"""Spectrogram front end on the device: waveform -> log-power spectrogram -> dataset statistics ->
``spect_to_img`` — what the reference does on the host with torchaudio and plain torch before every BiGAN step of the
spectrogram families (audio_mnist.py:59-61,116 ``Spectrogram(n_fft=255, win_length=128, pad=96)`` + ``(. + 1e-6).log()``;
:347-358 statistics pre-pass; :361-366 ``spect_to_img`` / ``img_to_spect``).  The wav / zip readers stay out of scope."""
import torch

from . import lib as _l
from . import ops


class LogSpectrogram:
    """``(Spectrogram(n_fft, win_length, pad)(wave) + eps).log()`` in one kernel; wave (N, L) float32 on the device ->
    (N, n_fft//2 + 1, frames)."""

    def __init__(self, n_fft=255, win_length=128, pad=96, eps=1e-6):
        self.n_fft, self.win_length, self.pad, self.eps = n_fft, win_length, pad, eps
        self.hop = win_length // 2

    def frames(self, L):
        return 1 + (L + 2 * self.pad + 2 * (self.n_fft // 2) - self.n_fft) // self.hop

    def __call__(self, wave: torch.Tensor) -> torch.Tensor:
        ops.require_cuda(wave)
        w = wave.detach().float().contiguous()
        N, L = w.shape
        out = torch.empty((N, self.n_fft // 2 + 1, self.frames(L)), dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            for lo in range(0, N, 65535):
                hi = min(N, lo + 65535)
                ops._launch("icf_log_spectrogram", _l.load().icf_log_spectrogram, ops.ptr(w, lo * L), hi - lo, L, self.n_fft,
                            self.win_length, self.hop, self.pad, self.eps, ops.ptr(out, lo * out.shape[1] * out.shape[2]),
                            out.shape[2])
        return out


class SpectrogramNormalizer:
    """Per-time-frame mean / std accumulated over a stream of (N, F, T) log-spectrograms (audio_mnist.py:347-358: the mean over
    the batches of the per-batch means — batches are weighted equally, as upstream) and ``spect_to_img`` / ``img_to_spect``."""

    def __init__(self, T, device, stds_kept=3.0):
        self.T, self.device, self.k = T, device, float(stds_kept)
        self.mean_sum = torch.zeros(T, dtype=torch.float32, device=device)
        self.sq_sum = torch.zeros(T, dtype=torch.float32, device=device)
        self.batches = 0
        self.mean = self.std = None

    def update(self, s: torch.Tensor):
        ops.require_cuda(s)
        x = s.detach().float().contiguous()
        rows = x.numel() // self.T
        a = torch.zeros(2, self.T, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ops._launch("icf_spect_stats", _l.load().icf_spect_stats, x.data_ptr(), rows, self.T, ops.ptr(a, 0), ops.ptr(a, self.T))
        self.mean_sum += a[0] / rows
        self.sq_sum += a[1] / rows
        self.batches += 1

    def finalize(self):
        self.mean = self.mean_sum / self.batches                      # E[X]
        # E[X^2] - E[X]^2 as upstream (:356-357), clamped at 0: a frame that is constant over the whole dataset (the zero-padded
        # first frame is log(eps) everywhere) leaves a difference of rounding size and either sign; upstream's sqrt returns NaN
        # for the negative ones and poisons every image of spect_to_img — here such a frame gets std 0 (-> divided by 1e-6).
        self.std = torch.sqrt((self.sq_sum / self.batches - self.mean.square()).clamp_min_(0.0))
        return self.mean, self.std

    def to_img(self, s: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        x = s.detach().float().contiguous()
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
        with torch.cuda.device(x.device):
            ops._launch("icf_spect_to_img", _l.load().icf_spect_to_img, x.data_ptr(), self.mean.data_ptr(), self.std.data_ptr(),
                        x.numel() // self.T, self.T, self.k, out.data_ptr(), ops.code_of(out))
        return out

    def to_spect(self, img: torch.Tensor) -> torch.Tensor:
        """img_to_spect (audio_mnist.py:365-366): the inverse affine map (host-side post-processing of generated images)."""
        return img.float() * self.k * (self.std + 1e-6) + self.mean
