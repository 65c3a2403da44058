"""TEST INFRASTRUCTURE ONLY — CPU restatement of the spectrogram front end.

The reference builds its log-spectrograms with third-party **torchaudio** (requirements.txt:2, unpinned; this image has
torchaudio 2.11): ``torchaudio.transforms.Spectrogram(n_fft=255, win_length=128, pad=96)`` (audio_mnist.py:59-61), then
``(spec + 1e-6).log()`` (:116), a statistics pre-pass (:347-358) and ``spect_to_img`` (:361-363).  ``log_spectrogram`` restates
torchaudio.functional.spectrogram through ``torch.stft`` (zero padding by ``pad``, centre=True / reflect, periodic Hann window of
win_length centred in the n_fft frame, power 2); tests/test_oracle_cpu.py pins it to torchaudio's own transform where torchaudio is
importable.  Only tests/ may import this file."""
import torch


def log_spectrogram(wave: torch.Tensor, n_fft=255, win_length=128, pad=96, eps=1e-6) -> torch.Tensor:
    x = torch.nn.functional.pad(wave.double(), (pad, pad))
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    s = torch.stft(x, n_fft, hop_length=win_length // 2, win_length=win_length, window=w, center=True, pad_mode="reflect",
                   normalized=False, onesided=True, return_complex=True)
    return (s.abs().pow(2) + eps).log()


def frame_stats(batches):
    """audio_mnist.py:347-358: mean over the batches of the per-batch mean / mean square over (clip, frequency), per time frame."""
    mean = sum(b.double().mean(dim=(0, 1)) for b in batches) / len(batches)
    ss = sum(b.double().square().mean(dim=(0, 1)) for b in batches) / len(batches)
    return mean, torch.sqrt(ss - mean.square())


def spect_to_img(s, mean, std, stds_kept=3):
    """audio_mnist.py:361-363."""
    return torch.clip((s.double() - mean) / (std + 1e-6), -stds_kept, stds_kept) / float(stds_kept)
