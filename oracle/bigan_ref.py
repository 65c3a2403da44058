"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's conditional-BiGAN hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / timed CPU baseline. The product
(``imagecfgen-pytorch_b200/``) never imports it and has no CPU fallback.

What it restates (all fp32, CPU): the ``Encoder`` / ``Generator`` / ``Discriminator`` forwards of the four
families in ``/root/reference/image_scms`` written *functionally* over a plain ``dict`` of tensors with the
reference's ``state_dict`` keys, the BCE-with-logits loss, Adam, and the phase A-D train-step body.
The arithmetic of the reference lives in third-party ``torch`` (``requirements.txt:1``, unpinned; this image
has torch 2.11.0): convolution / batch-norm primitives are therefore executed with the same
``torch.nn.functional`` CPU kernels the reference dispatches to, the loss / Adam / dropout / attribute-plane
logic is restated explicitly, and ``oracle/np_ops.py`` restates the primitives themselves in numpy and is
checked against torch in ``tests/test_oracle_cpu.py``.

PARITY PIN: the reference has no golden vectors.  The oracle is pinned against
outputs of the reference itself: ``tests/golden/make_golden.py`` imports ``/root/reference/image_scms``,
runs its modules, its loop body and ``mnist.train()`` on seeded inputs and commits digests under
``tests/golden/``; ``tests/test_oracle_cpu.py`` replays them through this file.

Dropout masks and ``z`` are *inputs* here (the reference draws them from the global RNG):
``masks`` is the list of ``(N,C,1,1)`` tensors ``torch.empty(N,C,1,1).bernoulli_(1-p).div_(1-p)`` in the
order the reference consumes them — dx sites, then dz, then dxz (mnist.py:151-154).
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .arch import FAMILIES

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------
# generic Sequential interpreter
# ---------------------------------------------------------------------------------------------
def bf16_storage(t: Tensor) -> Tensor:
    """Round to bfloat16 and back with a straight-through gradient.  Passed as ``q`` it turns the oracle into the
    reference algorithm *with activations and weight operands stored in bf16* (fp32 accumulation, fp32 BatchNorm
    statistics / losses / Adam) — the arithmetic contract of the CUDA bf16 path.  Needed because LeakyReLU makes
    gradients discontinuous in the activations: a 5e-3 relative perturbation of the pre-activations flips the
    sign of ~0.4 % of them, each flip changes that element's derivative 10x (slope 0.1), i.e. a norm-wise
    gradient difference of ~sqrt(0.004*0.81) = 6 % against an fp32 run for ANY bf16 implementation."""
    return t + (t.detach().bfloat16().float() - t.detach())


_CONTRACTIONS = ("conv", "convT", "linear", "bn")


def run_ops(ops, sd: Dict[str, Tensor], x: Tensor, masks: Optional[List[Tensor]] = None,
            training: bool = True, bn_update: bool = True, q=None, q_final: bool = True,
            signs: Optional[List[Tensor]] = None, record: Optional[list] = None) -> Tensor:
    """Execute an op table from ``oracle/arch.py`` with the semantics of the torch layers it names.
    ``q`` (optional) is applied wherever the CUDA engine stores a tensor: operands (weights), and activations
    after [conv+activation+dropout] and after [BatchNorm+dropout].
    ``signs`` (optional, gradient parity tests): one bool tensor per LeakyReLU of the table, True where the
    implementation under test saw a positive pre-activation.  The LeakyReLU is then evaluated as the LINEAR map
    x * (1 | slope) selected by that mask, in forward and backward alike, so that both sides differentiate the same
    piecewise-linear branch: a pre-activation within rounding distance of zero otherwise flips the branch and changes
    that element's derivative by 1/slope, which no finite-precision implementation can be held to.
    ``record`` (optional, gradient parity tests): every contraction / BatchNorm appends (kind, key, args, input, output)
    with the output's gradient retained, from which ``gradient_rss`` derives the magnitude of the terms each parameter
    gradient sums."""
    qq = q if q is not None else (lambda t: t)
    signs = list(signs) if signs is not None else None

    def rec(kind, key, args, x_in, y):
        if record is not None and y.requires_grad:
            y.retain_grad()
            record.append((kind, key, args, x_in.detach(), y))
    n_ops = len(ops)
    for i, op in enumerate(ops):
        kind = op[0]
        if kind == "conv":
            _, key, stride, pad = op
            x_in = x
            x = F.conv2d(x, qq(sd[key + ".weight"]), sd[key + ".bias"], stride=stride, padding=pad)
            rec("conv", key, (stride, pad), x_in, x)
        elif kind == "convT":
            _, key, stride, pad, opad = op
            x_in = x
            x = F.conv_transpose2d(x, qq(sd[key + ".weight"]), sd[key + ".bias"], stride=stride,
                                   padding=pad, output_padding=opad)
            rec("convT", key, (stride, pad, opad), x_in, x)
        elif kind == "linear":
            x_in = x
            x = F.linear(x, qq(sd[op[1] + ".weight"]), sd[op[1] + ".bias"])
            rec("linear", op[1], (), x_in, x)
        elif kind == "unflatten":
            x = x.reshape(x.shape[0], *op[1])
        elif kind == "lrelu":
            if signs is not None:
                pos = signs.pop(0).reshape(x.shape)
                x = x * torch.where(pos, torch.ones((), dtype=x.dtype), torch.full((), op[1], dtype=x.dtype))
            else:
                x = F.leaky_relu(x, op[1])
        elif kind == "tanh":
            x = torch.tanh(x)
        elif kind == "drop":
            # nn.Dropout2d on a 4-D input: one Bernoulli(1-p)/(1-p) draw per (sample, channel)
            if training:
                m = masks.pop(0)
                assert m.shape == (x.shape[0], x.shape[1], 1, 1), (m.shape, x.shape)
                x = x * m
        elif kind == "bn":
            key = op[1]
            rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
            if training:
                # nn.BatchNorm2d defaults: eps 1e-5, momentum 0.1; batch stats (biased var) normalise,
                # running_var receives the unbiased estimate; num_batches_tracked += 1
                x = F.batch_norm(x, rm if bn_update else rm.clone(), rv if bn_update else rv.clone(),
                                 sd[key + ".weight"], sd[key + ".bias"], True, 0.1, 1e-5)
                if bn_update and (key + ".num_batches_tracked") in sd:
                    sd[key + ".num_batches_tracked"] += 1
            else:
                x = F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], False, 0.1, 1e-5)
            rec("bn", key, (), x, x)
        else:
            raise ValueError(kind)
        if q is not None:
            last = i == n_ops - 1
            if (last and q_final) or (not last and ops[i + 1][0] in _CONTRACTIONS):
                x = q(x)
    return x


def dropout_sites(family: str):
    """[(p, channel count)] of D's Dropout2d sites in RNG-consumption order (dx, dz, dxz)."""
    fam = FAMILIES[family]
    sites = []
    if family != "mnist":
        return sites
    chans = {"Dx": [5, 32, 64, 128, 256], "Dz": [512, 512], "Dxz": [1024, 1024, 1024]}
    for part in ("Dx", "Dz", "Dxz"):
        ps = [op[1] for op in fam[part] if op[0] == "drop"]
        sites += list(zip(ps, chans[part]))
    return sites


def draw_masks(family: str, n: int, generator: Optional[torch.Generator] = None, device=None) -> List[Tensor]:
    """Masks for ONE Discriminator forward, drawn exactly like torch's Dropout2d (feature_dropout)."""
    out = []
    for p, c in dropout_sites(family):
        out.append(torch.empty(n, c, 1, 1, device=device).bernoulli_(1 - p, generator=generator).div_(1 - p))
    return out


# ---------------------------------------------------------------------------------------------
# attribute handling
# ---------------------------------------------------------------------------------------------
def _embedding_plane(table: Tensor, onehot: Tensor, size=None, scale=None) -> Tensor:
    """nn.Sequential(Embedding(K,256), Unflatten(1,(1,16,16)), Upsample(nearest), Tanh) applied to
    ``onehot.argmax(1)`` (mnist.py:24-29,52; audio_mnist.py:177-185). argmax = first maximal index."""
    idx = onehot.argmax(1)
    e = F.embedding(idx, table).reshape(-1, 1, 16, 16)
    if size is not None:
        e = F.interpolate(e, size=size, mode="nearest")
    else:
        e = F.interpolate(e, scale_factor=scale, mode="nearest")
    return torch.tanh(e)


def _const_plane(v: Tensor, size) -> Tensor:
    """continuous_feature_map (mnist.py:17-18)."""
    return v.reshape(v.size(0), 1, 1, 1).repeat(1, 1, *size)


def image_features(family: str, sd: Dict[str, Tensor], X: Tensor, c: Dict[str, Tensor]) -> Tensor:
    """The channel stack fed to E.layers / D.dx."""
    fam = FAMILIES[family]
    H, W = fam["image"]
    if family == "mnist":
        # mnist.py:47-55 — [X, digit plane] + continuous planes in sorted key order
        cont = {k: _const_plane(v, (H, W)) for k, v in c.items() if k != "digit"}
        dig = _embedding_plane(sd["digit_embedding.0.weight"], c["digit"], size=(H, W))
        return torch.cat([X, dig] + [cont[k] for k in sorted(cont)], dim=1)
    X = X.reshape(-1, 1, H, W)
    if family == "esrf_acoustic":
        # esrf_acoustic.py:166-170
        hb = _embedding_plane(sd["has_boat_embedding.0.weight"], c["has_boat"], scale=fam["upsample"])
        cb = _const_plane(c["closest_boat"].reshape(-1, 1), (H, W))
        return torch.cat([X, hb, cb], dim=1)
    # audio_mnist.py:204-210, whalecalls.py:265-271 — one plane per attribute, sorted key order
    planes = [_embedding_plane(sd[f"embedding_dict.{k}.0.weight"], c[k], scale=fam["upsample"])
              for k in sorted(fam["attribute_dims"])]
    return torch.cat([X] + planes, dim=1)


def latent_features(family: str, sd: Dict[str, Tensor], z: Tensor, c: Dict[str, Tensor]) -> Tensor:
    """The vector fed to G.layers: z ++ soft embeddings (dense matmul, so mixtures are legal) ++ continuous."""
    fam = FAMILIES[family]
    if family == "mnist":
        # mnist.py:77-85
        dig = c["digit"].matmul(sd["digit_embedding.weight"]).reshape(-1, 256, 1, 1)
        cont = {k: _const_plane(v, (1, 1)) for k, v in c.items() if k != "digit"}
        return torch.cat([z, dig] + [cont[k] for k in sorted(cont)], dim=1)
    z = z.reshape(-1, fam["latent"])
    if family == "esrf_acoustic":
        # esrf_acoustic.py:201-205
        hb = c["has_boat"].matmul(sd["has_boat_embedding.weight"])
        return torch.cat([z, hb, c["closest_boat"].reshape(-1, 1)], dim=1)
    # audio_mnist.py:250-256, whalecalls.py:314-321
    embs = [c[k].float().matmul(sd[f"embedding_dict.{k}.weight"]) for k in sorted(fam["attribute_dims"])]
    return torch.cat([z] + embs, dim=1)


# ---------------------------------------------------------------------------------------------
# network forwards
# ---------------------------------------------------------------------------------------------
def _q(q, t):
    return q(t) if q is not None else t


def encoder_fwd(family: str, sd, X, c, q=None, signs=None, record=None) -> Tensor:
    """Encoder.forward (mnist.py:46-56 etc.) -> (N,512,1,1)."""
    return run_ops(FAMILIES[family]["E"], sd, _q(q, image_features(family, sd, X, c)), q=q, signs=signs, record=record)


def generator_fwd(family: str, sd, z, c, q=None, signs=None, record=None) -> Tensor:
    """Generator.forward (mnist.py:76-86 etc.) -> (N,1,H,W)."""
    return run_ops(FAMILIES[family]["G"], sd, _q(q, latent_features(family, sd, z, c)), q=q, signs=signs, record=record)


def discriminator_fwd(family: str, sd, X, z, c, masks=None, training=True, bn_update=True, q=None, signs=None,
                      record=None) -> Tensor:
    """Discriminator.forward (mnist.py:142-154 etc.) -> logits (N,1).  ``signs``: {"Dx": [...], "Dz": [...], "Dxz": [...]}."""
    sg = signs or {}
    fam = FAMILIES[family]
    masks = list(masks) if masks is not None else None
    if training and dropout_sites(family) and masks is None:
        raise ValueError("training-mode MNIST discriminator needs explicit dropout masks")
    feats = image_features(family, sd, X, c)
    zin = z.reshape(-1, fam["latent"], 1, 1)
    if q is not None and not (training and dropout_sites(family)):
        feats, zin = q(feats), q(zin)          # with input dropout the engine rounds after the mask (first op)
    dx = run_ops(fam["Dx"], sd, feats, masks, training, bn_update, q=q, signs=sg.get("Dx"), record=record)
    dz = run_ops(fam["Dz"], sd, zin, masks, training, bn_update, q=q, signs=sg.get("Dz"), record=record)
    out = run_ops(fam["Dxz"], sd, torch.cat([dx, dz], dim=1), masks, training, bn_update, q=q, q_final=False,
                  signs=sg.get("Dxz"), record=record)
    return out.reshape(-1, 1)


def gradient_rss(record, sd) -> Dict[str, Tensor]:
    """Root-sum-square of the TERMS every parameter gradient sums (after ``backward``), from the records ``run_ops`` kept:
    a weight gradient is dW = sum_{n,p,q} dY * X, a bias gradient sum dY, BatchNorm's sum dY * xhat / sum dY.  When those
    terms cancel (a conv bias in front of a BatchNorm, a near input-independent network at the as-shipped init scale) the
    sum is far smaller than its terms and its relative error has no meaning; the error of ANY summation scales with the
    terms.  Parity tests therefore bound |g - g_ref| by tol * max(|g_ref|, rss) norm-wise per tensor.  The squares are
    pushed through the same linear operator: sum (dY*X)^2 = d/dW0 <op(X^2, W0), dY^2>."""
    acc: Dict[str, Tensor] = {}

    def add(k, v):
        acc[k] = acc[k] + v if k in acc else v
    for kind, key, args, x_in, y in record:
        if y.grad is None:
            continue
        gy2 = y.grad.detach().double() ** 2
        if kind == "bn":
            g, b = sd[key + ".weight"].detach().double(), sd[key + ".bias"].detach().double()
            xhat = (y.detach().double() - b.reshape(1, -1, 1, 1)) / g.reshape(1, -1, 1, 1)
            add(key + ".weight", (gy2 * xhat ** 2).sum(dim=(0, 2, 3)))
            add(key + ".bias", gy2.sum(dim=(0, 2, 3)))
            continue
        w0 = torch.zeros_like(sd[key + ".weight"], dtype=torch.float64).requires_grad_()
        x2 = x_in.double() ** 2
        if kind == "conv":
            y2 = F.conv2d(x2, w0, None, stride=args[0], padding=args[1])
        elif kind == "convT":
            y2 = F.conv_transpose2d(x2, w0, None, stride=args[0], padding=args[1], output_padding=args[2])
        else:
            y2 = F.linear(x2, w0)
        (gw,) = torch.autograd.grad(y2, w0, gy2)
        add(key + ".weight", gw)
        add(key + ".bias", gy2.sum(dim=[d for d in range(gy2.dim()) if d != 1]))
    return {k: v.clamp_min(0).sqrt().float() for k, v in acc.items()}


def counterfactual(family: str, E_sd, G_sd, x, c, c_cf) -> Tensor:
    """G(E(x, c), c_cf) under no_grad (mnist_gan_counterfactuals.py:71)."""
    with torch.no_grad():
        return generator_fwd(family, G_sd, encoder_fwd(family, E_sd, x, c), c_cf)


# ---------------------------------------------------------------------------------------------
# loss and optimiser
# ---------------------------------------------------------------------------------------------
def bce_with_logits(logits: Tensor, target: Tensor) -> Tensor:
    """nn.BCEWithLogitsLoss() (mean): max(l,0) - l*t + log1p(exp(-|l|))  (mnist.py:181)."""
    l = logits
    return (l.clamp(min=0) - l * target + torch.log1p(torch.exp(-l.abs()))).mean()


class AdamState:
    """torch.optim.Adam(lr, betas, eps=1e-8, weight_decay=0, amsgrad=False) restated (mnist.py:176-179)."""

    def __init__(self, params: List[Tensor], lr=1e-4, betas=(0.5, 0.999), eps=1e-8):
        self.params = params
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    def step(self, grads: List[Optional[Tensor]]):
        b1, b2 = self.betas
        self.step_count += 1
        t = self.step_count
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                if g is None:
                    continue
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-self.lr / bc1)


def _leaf_params(sd):
    """Trainable tensors of a state dict, in state_dict order (== module.parameters() order)."""
    return [(k, v) for k, v in sd.items()
            if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]


class BiGANOracle:
    """Holds E/G/D state dicts + two Adam states and replays the reference loop body."""

    def __init__(self, family: str, E_sd, G_sd, D_sd, lr=1e-4, betas=None, q=None):
        self.family = family
        self.q = q                  # None = the fp32 reference; bf16_storage = bf16-storage arithmetic contract
        fam = FAMILIES[family]
        betas = betas if betas is not None else fam["adam_betas"]

        def own(sd):
            return {k: v.detach().clone().float() if v.is_floating_point() else v.detach().clone()
                    for k, v in sd.items()}
        self.E, self.G, self.D = own(E_sd), own(G_sd), own(D_sd)
        for sd in (self.E, self.G, self.D):
            for k, v in _leaf_params(sd):
                v.requires_grad_(True)
        # mnist.py:176-179: optimizer_E covers E.parameters() + G.parameters()
        self.pE = [v for _, v in _leaf_params(self.E)] + [v for _, v in _leaf_params(self.G)]
        self.pD = [v for _, v in _leaf_params(self.D)]
        self.optE = AdamState(self.pE, lr, betas)
        self.optD = AdamState(self.pD, lr, betas)

    def _zero(self, params):
        for p in params:
            p.grad = None

    def grads(self, which):
        sd = getattr(self, which)
        return {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in _leaf_params(sd)}

    def train_step(self, images, c, z, masks6, phase_a=True, keep_grads=False):
        """One iteration of the hot loop, mnist.py:220-248 (audio_mnist.py:396-420 is identical).

        ``images`` already scaled to [-1,1]; ``c`` already scaled; ``masks6`` = six mask lists, one per
        Discriminator forward in the order the loop performs them (A-valid, A-fake, B, C, D-fake, D-valid).
        Returns dict(loss_EG, loss_D_valid, loss_D_fake, DG_mean, DE_mean [, grads_*]).
        """
        fam = self.family
        n = images.shape[0]
        valid = torch.ones(n, 1, device=images.device)
        fake = torch.zeros(n, 1, device=images.device)
        masks6 = [list(m) if m is not None else None for m in masks6]
        out = {}
        # Phase A — E+G update (mnist.py:224-230)
        if phase_a:
            self._zero(self.pE)
            self._zero(self.pD)
            D_valid = discriminator_fwd(fam, self.D, images, encoder_fwd(fam, self.E, images, c, q=self.q), c, masks6[0], q=self.q)
            D_fake = discriminator_fwd(fam, self.D, generator_fwd(fam, self.G, z, c, q=self.q), z, c, masks6[1], q=self.q)
            loss_EG = (bce_with_logits(D_valid, fake) + bce_with_logits(D_fake, valid)) / 2
            loss_EG.backward()
            if keep_grads:
                out["grads_A_E"], out["grads_A_G"] = self.grads("E"), self.grads("G")
            self.optE.step([p.grad for p in self.pE])
            out["loss_EG"] = float(loss_EG)
        # Phase B — D on real pairs (mnist.py:232-236)
        self._zero(self.pD)
        D_valid = discriminator_fwd(fam, self.D, images, encoder_fwd(fam, self.E, images, c, q=self.q), c, masks6[2], q=self.q)
        loss_D = bce_with_logits(D_valid, valid)
        loss_D.backward()
        if keep_grads:
            out["grads_B_D"] = self.grads("D")
        self.optD.step([p.grad for p in self.pD])
        out["loss_D_valid"] = float(loss_D)
        # Phase C — D on generated pairs (mnist.py:237-241)
        self._zero(self.pD)
        D_fake = discriminator_fwd(fam, self.D, generator_fwd(fam, self.G, z, c, q=self.q), z, c, masks6[3], q=self.q)
        loss_D = bce_with_logits(D_fake, fake)
        loss_D.backward()
        if keep_grads:
            out["grads_C_D"] = self.grads("D")
        self.optD.step([p.grad for p in self.pD])
        out["loss_D_fake"] = float(loss_D)
        # Phase D — scores (mnist.py:243-248); D stays in train mode (dropout + BN stat updates)
        with torch.no_grad():
            Gz = generator_fwd(fam, self.G, z, c, q=self.q)
            EX = encoder_fwd(fam, self.E, images, c, q=self.q)
            DG = discriminator_fwd(fam, self.D, Gz, z, c, masks6[4], q=self.q).sigmoid()
            DE = discriminator_fwd(fam, self.D, images, EX, c, masks6[5], q=self.q).sigmoid()
        out["DG_mean"] = float(DG.mean())
        out["DE_mean"] = float(DE.mean())
        self._zero(self.pE)
        self._zero(self.pD)
        return out


def scale_images_mnist(images_u8: Tensor) -> Tensor:
    """mnist.py:204."""
    return 2 * images_u8.reshape(-1, 1, 28, 28).float() / 255 - 1


def scale_attrs_mnist(attrs: Dict[str, Tensor], stats: Dict[str, tuple]) -> Dict[str, Tensor]:
    """mnist.py:205-209: min-max scale continuous attributes to [-1,1]; digit passes through."""
    c = {k: 2 * (attrs[k] - stats[k][0]) / (stats[k][1] - stats[k][0]) - 1 for k in stats}
    c["digit"] = attrs["digit"]
    return c


def digest(t: Tensor) -> dict:
    """Order-sensitive fingerprint of a tensor in float64 (what tests/golden/ stores)."""
    t = t.detach().double().reshape(-1)
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64)
    w = (w % 97 + 1) / 97.0
    return {"n": int(t.numel()), "sum": float(t.sum()), "abs": float(t.abs().sum()),
            "wsum": float((t * w).sum()), "l2": float(t.square().sum().sqrt()),
            "head": [float(v) for v in t[:4]]}


# ---------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by tests/golden/make_golden.py and the parity tests)
# ---------------------------------------------------------------------------------------------
def synth_state_dict(family: str, net: str, seed: int, std: float = 0.05) -> Dict[str, Tensor]:
    """A full state_dict (reference keys/shapes) drawn from a seeded CPU generator:
    conv/linear weights N(0,std), biases N(0,0.02), embeddings N(0,1), BN gamma 1+N(0,0.1), beta N(0,0.1),
    running stats at their initial values.  'Trained-like' scale so that parity is well conditioned (§4)."""
    from .arch import param_shapes
    g = torch.Generator().manual_seed(seed * 7919 + {"E": 1, "G": 2, "D": 3}[net])
    sd = {}
    for key, shape in param_shapes(family, net):
        leaf = key.rsplit(".", 1)[1]
        is_bn = len(shape) == 1 and any(k == key.rsplit(".", 1)[0] + ".running_mean" for k, _ in param_shapes(family, net))
        if "embedding" in key:
            t = torch.randn(shape, generator=g)
        elif leaf == "running_mean":
            t = torch.zeros(shape)
        elif leaf == "running_var":
            t = torch.ones(shape)
        elif is_bn and leaf == "weight":
            t = 1 + 0.1 * torch.randn(shape, generator=g)
        elif is_bn and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight":
            t = std * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[key] = t
        if leaf == "running_var":
            sd[key.rsplit(".", 1)[0] + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd
