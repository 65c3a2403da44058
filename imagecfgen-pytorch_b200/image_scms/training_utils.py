"""Mirror of image_scms/training_utils.py (reference :6-119): batching generators, ``init_weights``, the
``AdversariallyLearnedInference`` wrapper and its reconstruction losses, so that scripts importing them keep working.
``ssim`` restates the metric the reference imports from the third-party ``pytorch_msssim`` package (training_utils.py:3,
requirements.txt: unpinned, not installed here): it is caller-side loss arithmetic on the networks' outputs, plain torch.
WGAN-GP helpers are out of scope (unused by every BiGAN path)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from icf_b200.modules import init_weights  # noqa: F401  (same semantics as training_utils.py:114-119)


def batchify(*tensors, batch_size=128, device='cpu'):
    """Contiguous slices of every tensor; the last batch may be short (training_utils.py:6-13; ``device`` is
    accepted and unused exactly as upstream)."""
    n = min(map(len, tensors))
    for i in range(0, n, batch_size):
        yield tuple(x[i:i + batch_size] for x in tensors)


def batchify_dict(tensors: dict, batch_size=128, device='cpu'):
    """training_utils.py:16-27."""
    n = min(map(len, tensors.values()))
    for i in range(0, n, batch_size):
        yield {k: v[i:i + batch_size] for k, v in tensors.items()}


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, K=(0.01, 0.03)):
    """Structural similarity as pytorch_msssim.ssim computes it (Wang et al. 2004): separable Gaussian window (11 taps,
    sigma 1.5, 'valid' — no padding), per-channel filtering, C1 = (K1 L)^2, C2 = (K2 L)^2, mean over the map and the
    channels per image; ``size_average`` then averages over the batch.  Inputs (N,C,H,W)."""
    if X.shape != Y.shape:
        raise ValueError(f"Input images should have the same dimensions, but got {X.shape} and {Y.shape}.")
    coords = torch.arange(win_size, dtype=X.dtype, device=X.device) - win_size // 2
    g = torch.exp(-(coords ** 2) / (2 * win_sigma ** 2))
    g = (g / g.sum())
    C_ = X.shape[1]

    def blur(t):
        t = F.conv2d(t, g.reshape(1, 1, -1, 1).repeat(C_, 1, 1, 1), groups=C_) if t.shape[2] >= win_size else t
        return F.conv2d(t, g.reshape(1, 1, 1, -1).repeat(C_, 1, 1, 1), groups=C_) if t.shape[3] >= win_size else t
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = blur(X), blur(Y)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = blur(X * X) - mu1_sq, blur(Y * Y) - mu2_sq, blur(X * Y) - mu12
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    per_image = torch.flatten(ssim_map, 2).mean(-1).mean(1)
    return per_image.mean() if size_average else per_image


def log_loss(score_0, score_1, eps=1e-6):
    """Probability-space ALI loss (training_utils.py:49-51)."""
    return -torch.mean(torch.log(score_1 + eps) + torch.log(1 - score_0 + eps))


class AdversariallyLearnedInference(nn.Module):
    """training_utils.py:54-111: returns (D(G(z),z,a), D(x,E(x),a)); imported but never called by the train loops."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, discriminator: nn.Module):
        super().__init__()
        self.encoder, self.decoder, self.discriminator = encoder, decoder, discriminator

    def __call__(self, x, z, a=None, add_noise=False, noise_scale=0.1):
        extra = () if a is None else (a,)
        ex = self.encoder(x, *extra)
        gz = self.decoder(z, *extra)
        xin = x + torch.normal(0, noise_scale, x.shape).to(x.device) if add_noise else x
        return self.discriminator(gz, z, *extra), self.discriminator(xin, ex, *extra)

    def discriminator_loss(self, x, z, a=None, eps=1e-6, **kwargs):
        dg, de = self(x, z, a=a, **kwargs)
        return log_loss(dg, de, eps)

    def generator_loss(self, x, z, a=None, eps=1e-6, **kwargs):
        dg, de = self(x, z, a=a, **kwargs)
        return log_loss(de, dg, eps)

    def rec_loss(self, x, z=None, a=None, metric='ssim'):
        """training_utils.py:91-111 (default metric 'ssim' as upstream)."""
        if metric == 'mse':
            def loss(Y, X):
                return torch.square(Y - X).mean()
        elif metric == 'ssim':
            def loss(Y, X):
                return 1 - ssim(Y, X, data_range=1.0, size_average=True)
        else:
            raise ValueError(f'Invalid metric {metric}')
        extra = () if a is None else (a,)
        if z is None:
            z = self.encoder(x, *extra)
        return loss(x, self.decoder(z, *extra))
