"""Gradient-based counterfactual explainers on the device (explain/cf_example.py:17-170, called by
morphomnist_cf_metrics.py:73-92 and mnist_oracle_scores.py:88-107).

``HingeLossCFExplainer.explain`` upstream runs, per image, 30-100 Adam steps on a raw latent code and raw attribute rows: G
forward at batch 1, a classifier, ``loss.backward()`` — which also computes every weight gradient of G and of the classifier
(their parameters require grad; nobody reads or zeroes them) — and a torch Adam step over a handful of tiny tensors: ~100
launches of a few microseconds of work each, one image at a time.  Here

* all raw rows of a batch of images live in one flat buffer: tanh / softmax (``icf_explain_transform``), their backward
  (``icf_explain_backward``) and Adam (``icf_adam_step``) are three launches per step whatever the number of attributes;
* G runs its forward and its DATA-gradient kernels only (``NetExec.generator_backward(grads=None)``): F_G + dgrad instead of
  F_G + dgrad + wgrad, and the classifier is differentiated w.r.t. its input only;
* images are batched: the objective is the SUM of the per-image objectives, Adam is element-wise, so B images optimised
  together follow exactly the B single-image trajectories of upstream (which is the B = 1 case);
* the whole step can be captured once as a CUDA graph and replayed ``steps`` times (``graph=True``) — at batch 1 the step
  is launch-latency, not arithmetic.

The classifier is any torch module on the device (upstream's ``classifiers/`` are out of the hot path); everything of the
generator and the optimiser is libicf_b200.so.  There is no CPU path.
"""
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import ops
from .engine import Act

F32 = ops.F32
COPY, TANH, SOFTMAX = 0, 1, 2


def mse(a: torch.Tensor, b: torch.Tensor):
    """Per-row mean squared difference (cf_example.py:12-14)."""
    d = a - b
    return d.square().mean(dim=list(range(1, d.dim())))


def max_excluding(y: torch.Tensor, c: Union[int, torch.Tensor]):
    """Row-wise maximum over the classes other than ``c`` (cf_example.py:74-79, any batch size)."""
    idx = torch.as_tensor(c, device=y.device, dtype=torch.long).reshape(-1, 1).expand(y.shape[0], 1).contiguous()
    return y.masked_fill(torch.zeros_like(y, dtype=torch.bool).scatter_(1, idx, True), float("-inf")).max(dim=1).values


class DeepCounterfactualExplainer:
    """cf_example.py:17-71: decode the image's code under ``sample_points`` mixtures of the predicted and the target class and
    return the decodings the classifier assigns to the target, ordered by the chosen distance."""

    def __init__(self, encoder, decoder, classifier, target_feature: str):
        self.encoder, self.decoder, self.classifier, self.target_feature = encoder, decoder, classifier, target_feature

    @torch.no_grad()
    def explain(self, x: torch.Tensor, attrs: Dict[str, torch.Tensor], target_class: int, sample_points=100,
                metric="mixture") -> Tuple[torch.Tensor, torch.Tensor]:
        if metric not in ("mixture", "mse", "ssim"):
            raise ValueError(metric)
        ops.require_cuda(x)
        S, dev, tf = sample_points, x.device, self.target_feature
        codes = self.encoder(x, attrs)
        codes = codes.expand(S, *codes.shape[1:]).contiguous()
        original = int(self.classifier(x).argmax(1))
        K = attrs[tf].shape[1]
        cf = {k: v.expand(S, *v.shape[1:]).contiguous() for k, v in attrs.items() if k != tf}
        probs = torch.linspace(0, 1, S, device=dev).reshape(S, 1)
        cf[tf] = torch.zeros(S, K, device=dev)
        cf[tf][:, original] += (1 - probs)[:, 0]
        cf[tf][:, target_class] += probs[:, 0]
        samples = self.decoder(codes, cf)                               # soft one-hots: G takes them as a dense matmul
        preds = self.classifier(samples).argmax(1)
        if metric == "mixture":
            dist = probs
        elif metric == "mse":
            dist = mse(x, samples)
        else:
            from image_scms.training_utils import ssim
            xv = x.expand(S, *x.shape[1:])
            dist = 1 - ssim((xv + 1) / 2, (samples + 1) / 2, data_range=1.0, size_average=False)
        hit = preds == target_class
        if not bool(hit.any()):
            return samples, dist
        dist, samples = dist[hit], samples[hit]
        # upstream indexes with `metric_val.argsort()` as is (:69-71).  For 'mse' / 'ssim' that is a 1-D permutation; for the default
        # 'mixture' metric_val is (S', 1), whose argsort along the last axis is all zeros — upstream then returns S' copies of the
        # first hit with an extra axis.  Kept literally: callers see what the reference returns.
        order = dist.argsort()
        return samples[order], dist[order]


class HingeLossCFExplainer:
    """cf_example.py:82-170.  Same constructor and ``explain`` arguments; ``x`` / ``attrs`` may hold B > 1 images
    (``target_class`` then an int or one class per image), optimised together as B independent problems."""

    def __init__(self, encoder, decoder, classifier, target_feature: str, latent_dim: int,
                 categorical_features: Optional[List[str]] = None, features_to_ignore: Optional[List[str]] = None, c=10.0):
        self.encoder, self.decoder, self.classifier = encoder, decoder, classifier
        self.categorical_features = categorical_features or []
        self.features_to_ignore = features_to_ignore or []
        self.c, self.target_feature, self.latent_dim = c, target_feature, latent_dim
        self.last = None

    # ---- flat parameter layout -------------------------------------------------------------------------------------------
    def _layout(self, attrs, B, train_z):
        specs, off = [], 0
        for k in attrs:                                                  # upstream's dict order (:121-125), then z (:129-130)
            if k in self.features_to_ignore:
                continue
            w = attrs[k].shape[1] if attrs[k].dim() > 1 else 1
            specs.append((k, SOFTMAX if k in self.categorical_features else TANH, w, off))
            off += B * w
            off = (off + 3) // 4 * 4                                     # 16-byte aligned groups (float4 Adam, dz buffer)
        if train_z:
            specs.append(("z", TANH, self.latent_dim, off))
            off += B * self.latent_dim
        return specs, (off + 3) // 4 * 4

    def draw_init(self, attrs, B, train_z, device, generator=None):
        """The starting point upstream draws (:121-130): 0.01*randn per attribute row, randn for z — on the device's generator."""
        init = {k: 0.01 * torch.randn((B, attrs[k].shape[1]), device=device, generator=generator)
                for k in attrs if k not in self.features_to_ignore}
        if train_z:
            init["z"] = torch.randn((B, self.latent_dim), device=device, generator=generator)
        return init

    def _objective(self, x, x_cf, target, original_pred):
        """Sum over the images of  c * hinge + mean|x - x_cf|  (:107-119), and the two parts per image."""
        pred = self.classifier(x_cf)
        if target is not None:
            h = max_excluding(pred, target) - pred.gather(1, target.reshape(-1, 1))[:, 0]
        else:
            h = (pred - original_pred).square().mean(dim=1)
        m = (x - x_cf).abs().flatten(1).mean(dim=1)
        return (self.c * h + m).sum(), h, m

    def explain(self, x: torch.Tensor, attrs: Dict[str, torch.Tensor], target_class=None, train_z=True, steps=30, lr=0.1, *,
                init: Optional[Dict[str, torch.Tensor]] = None, graph=False, history: Optional[list] = None, optimise_z=False):
        """-> x_cf (B,1,H,W).  ``init``: starting raw rows per optimised attribute (+ 'z'), default = upstream's random draw;
        ``graph``: capture one step as a CUDA graph and replay it; ``history``: list receiving (hinge, rec) per step.

        ``train_z`` follows upstream to the letter: params["z"] is created (:129-130) after the loop that sets requires_grad
        (:126-127), so the latent row never receives a gradient and Adam skips it — the image is decoded from a FIXED random
        tanh(z) instead of E(x) (pinned by tests/golden/explain_mnist_s21.pt).  ``optimise_z=True`` trains the row, which is
        what the name promises."""
        ops.require_cuda(x)
        G = self.decoder
        ex = G.engine()
        dev, lat = ex.device, self.latent_dim
        x = x.detach().float().reshape(-1, 1, ex.H, ex.W).contiguous()
        B = x.shape[0]
        with torch.no_grad():
            codes = self.encoder(x, attrs).detach().reshape(B, lat).float().contiguous()       # :98
            original_pred = self.classifier(x).softmax(1)                                       # :100
        target = None
        if target_class is not None:
            target = torch.as_tensor(target_class, device=dev, dtype=torch.long).reshape(-1).expand(B).contiguous()
        specs, n = self._layout(attrs, B, train_z)
        if init is None:
            init = self.draw_init(attrs, B, train_z, dev)
        with torch.cuda.device(dev):
            raw = torch.zeros(n, dtype=torch.float32, device=dev)
            for k, _, w, off in specs:
                raw[off:off + B * w].copy_(init[k].detach().float().reshape(B * w))
            out, draw = torch.zeros_like(raw), torch.zeros_like(raw)
            m1, m2 = torch.zeros_like(raw), torch.zeros_like(raw)
            state = torch.tensor([0, lr, 0.9, 0.999, 1e-8, 1.0, 0, 0], dtype=torch.float32, device=dev)    # torch.optim.Adam defaults (:132)
            view = {k: out[off:off + B * w].view(B, w) for k, _, w, off in specs}
            fixed = {k: attrs[k].detach().float().contiguous() for k in attrs if k in self.features_to_ignore}
            cat_names = [a[0] for a in ex.fam.cat_attrs]
            fwd_groups = ops.explain_groups([(mode, w, off, None) for _, mode, w, off in specs])
            stats = torch.zeros(2, B, dtype=torch.float32, device=dev)

            def decode(save):
                ops.explain_transform(raw.data_ptr(), out.data_ptr(), fwd_groups, B)
                c_cf = {**{k: view[k] for k, *_ in specs if k != "z"}, **fixed}
                z = view["z"] if train_z else codes
                img, st = ex.generator_forward(B, z.data_ptr(), F32, lat, c_cf, save=save)
                x_cf = torch.empty((B, 1, ex.H, ex.W), dtype=torch.float32, device=dev)
                ops.cast(img.ptr, img.code, x_cf.data_ptr(), F32, x_cf.numel())
                return x_cf, st, c_cf

            def step():
                x_cf, st, c_cf = decode(True)
                x_cf.requires_grad_(True)
                with torch.enable_grad():
                    loss, h, m = self._objective(x, x_cf, target, original_pred)
                (g,) = torch.autograd.grad(loss, x_cf)                   # the classifier's own weight gradients are not formed
                stats[0].copy_(h.detach())
                stats[1].copy_(m.detach())
                dz, doh, dco = ex.generator_backward(st, Act(g.reshape(-1, 1).contiguous(), 1), None,
                                                     need_dz=train_z and optimise_z, need_dattr=True)
                cont_names = ex.cont_names(c_cf)
                douts = []
                for k, mode, w, off in specs:
                    if k == "z":
                        d = dz
                    elif k in cat_names:
                        d = doh[cat_names.index(k)]
                    else:
                        d = dco[cont_names.index(k)]
                    douts.append((mode, w, off, d.data_ptr() if d is not None else None))   # no gradient: row stays put
                ops.explain_backward(out.data_ptr(), draw.data_ptr(), ops.explain_groups(douts), B)
                ops.adam_step(raw.data_ptr(), draw.data_ptr(), m1.data_ptr(), m2.data_ptr(), n, state.data_ptr())
                return douts

            if graph and steps > 0:
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):                               # warm-up on a side stream (allocator, autograd, cuBLAS handles)
                    saved = [t.clone() for t in (raw, m1, m2, state)]
                    step()
                    for t, v in zip((raw, m1, m2, state), saved):
                        t.copy_(v)
                torch.cuda.current_stream().wait_stream(s)
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    step()
                for _ in range(steps):                                   # capture records, replay executes
                    g_.replay()
                    if history is not None:
                        history.append(stats.clone())
            else:
                for _ in range(steps):
                    step()
                    if history is not None:
                        history.append(stats.clone())
            x_cf, _, c_cf = decode(False)                                # :160-169 final decode at the optimised point
            self.last = {"raw": {k: raw[off:off + B * w].view(B, w).clone() for k, _, w, off in specs},
                         "attrs_cf": {k: v.clone() for k, v in c_cf.items()}, "grad_raw": draw.clone(), "specs": specs}
        return x_cf
