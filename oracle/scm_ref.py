"""TEST INFRASTRUCTURE ONLY — CPU restatement of the attribute-SCM "intervene" step and of the fine-tune loop body
(SURVEY.md §8f rows N1, N2).  Only tests/ (and bench.py's cpu_baseline legs) may import it; the product never does.

N2.  The reference intervenes with ``CausalModuleGraph.sample_cf`` (attribute_scms/graph.py:144-184): every observed
variable's exogenous noise is recovered (``recover_noise`` :68-90 -> ``nf_inverse`` through the variable's transforms) and
the variables not held by the intervention are regenerated from that noise with the counterfactual parent values
(:166-182 ``generate`` -> ``nf_forward``).  MorphoMNIST's only edge is thickness -> intensity (attribute_scms/mnist.py:48)
with transforms [conditional_affine_autoregressive(1, 1), SigmoidTransform, AffineTransform(i_min, i_max - i_min)]
(:28-33).  The arithmetic lives in third-party **pyro-ppl** (requirements.txt:5, UNPINNED; not installed in this image,
so the reference's attribute_scms package cannot be imported here): restated below from the published algorithm of
pyro.distributions.transforms.ConditionalAffineAutoregressive / AffineAutoregressive (y = mean + exp(clamp(log_scale,
-5, 3)) * x, (mean, log_scale) = arn(x, context)) and pyro.nn.ConditionalAutoRegressiveNN (masked MLP; for a 1-D variable
with a 1-D context the masks leave  out = W2 relu(W1[:, context] c + b1) + b2,  hidden width 10 by default),
torch.distributions SigmoidTransform (inverse clamps to [tiny, 1 - eps]) and AffineTransform.
PARITY UNPINNED for the learned (hyper-network) form: no reference run is possible without pyro.  The closed form is
pinned to the reference's own data-generating SCM (create_train_dataset.py:42-46: intensity = 191*sigmoid(0.5*eps + 2t - 5)
+ 64): abduct-then-regenerate with the factual parent must return the observed intensity, and with the counterfactual
parent must equal generate_i(t_cf, noise=eps) — checked in tests/test_oracle_cpu.py.

N1.  ``finetune_step`` restates finetune_mnist_bigan.py:68-86 / finetune_whale_bigan.py:58-73 over the functional oracle
networks (autograd through G into E, Adam over E only); ``all_pairs`` mirrors the (N,H,W) - (N,1,H,W) broadcast of
finetune_whale_bigan.py:59-65.
"""
from typing import Dict, Optional

import torch

from . import bigan_ref as R


def hyper_net(ctx: torch.Tensor, hyper: Optional[Dict[str, torch.Tensor]], closed, clip=(-5.0, 3.0)):
    """(loc, clamped log scale) of the conditional affine transform for a context value of shape (N,)."""
    if hyper is None:
        loc = closed[0] + closed[1] * ctx
        ls = torch.full_like(ctx, closed[2])
    else:
        h = torch.relu(ctx[:, None] * hyper["w1"][None, :] + hyper["b1"][None, :])       # (N,H)
        o = h @ hyper["w2"].t() + hyper["b2"][None, :]                                     # (N,2): loc, log scale
        loc, ls = o[:, 0], o[:, 1]
    return loc, ls.clamp(clip[0], clip[1])


def affine_sigmoid_cf(value, parent, parent_cf, lo, span, closed=(0.0, 0.0, 0.0), hyper=None, clip=(-5.0, 3.0)):
    """Abduction (graph.py:68-90) and regeneration (:166-182) of one conditional affine -> sigmoid -> affine mechanism.
    -> (value_cf, noise), float64 arithmetic."""
    v, p, pc = (t.detach().double().reshape(-1) for t in (value, parent, parent_cf))
    hy = {k: t.double() for k, t in hyper.items()} if hyper is not None else None
    fi = torch.finfo(torch.float32)
    u = ((v - lo) / span).clamp(fi.tiny, 1.0 - fi.eps)          # AffineTransform^-1, then SigmoidTransform^-1's clamp
    s = u.log() - (-u).log1p()
    loc, ls = hyper_net(p, hy, closed, clip)
    eps = (s - loc) * torch.exp(-ls)                             # AffineAutoregressive^-1
    loc2, ls2 = hyper_net(pc, hy, closed, clip)
    s2 = loc2 + torch.exp(ls2) * eps
    return lo + span * torch.sigmoid(s2), eps


def finetune_step(family, E_sd, G_sd, adam: "R.AdamState", x, c, metric="mse", all_pairs=False):
    """One iteration of finetune_mnist_bigan.py:68-86; E_sd's floating tensors must be the leaves ``adam`` updates.
    -> (rec_loss, latent_loss)."""
    fam = R.FAMILIES[family]
    H, W = fam["image"]
    for p in adam.params:
        p.grad = None
    xi = x.reshape(-1, 1, H, W)
    codes = R.encoder_fwd(family, E_sd, xi, c)
    xr = R.generator_fwd(family, G_sd, codes, c)
    if metric == "ssim":
        import sys, os
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imagecfgen-pytorch_b200"))
        from image_scms.training_utils import ssim
        rec = 1 - ssim(xi, xr, data_range=1.0).mean()
    elif all_pairs:
        rec = torch.square(x.reshape(-1, H, W) - xr).mean()      # (N,H,W) - (N,1,H,W) -> (N,N,H,W)
    else:
        rec = torch.square(xi - xr).mean()
    latent = torch.square(codes).mean()
    (rec + latent).backward()
    adam.step([p.grad for p in adam.params])
    return float(rec.detach()), float(latent.detach())
