"""Mirror of image_scms/training_utils.py (reference :6-119): batching generators, ``init_weights`` and the
``AdversariallyLearnedInference`` wrapper, so that scripts importing them keep working.  WGAN-GP helpers and
the SSIM reconstruction metric are out of scope (unused by every BiGAN path; pytorch_msssim is not needed)."""
import torch
import torch.nn as nn

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

    def rec_loss(self, x, z=None, a=None, metric='mse'):
        if metric != 'mse':
            raise ValueError("only metric='mse' is available (pytorch_msssim is outside the hot path)")
        extra = () if a is None else (a,)
        if z is None:
            z = self.encoder(x, *extra)
        return torch.square(x - self.decoder(z, *extra)).mean()
