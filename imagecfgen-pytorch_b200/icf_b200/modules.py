"""nn.Module boundary of the hot path: drop-in ``Encoder`` / ``Generator`` / ``Discriminator``.

The classes built here keep the reference's constructor-free interface, ``forward`` signatures, ``.device``
property, parameter names / shapes / registration order (``state_dict`` layout of SURVEY.md App. A.5) and
default initialisation, but own *only parameters*: their ``forward`` hands raw device pointers to
libicf_b200.so through ``icf_b200.engine.NetExec`` inside a ``torch.autograd.Function`` (one per network),
whose ``backward`` runs the hand-written dgrad / wgrad kernels.  Reference: image_scms/mnist.py:21-154,
audio_mnist.py:173-318, whalecalls.py:230-387, esrf_acoustic.py:134-260.
"""
import math
import os
from typing import Dict

import torch
import torch.nn as nn

from . import ops
from .arch import FAMILIES, Family, L
from .engine import Act, NetExec, dtype_code

DEFAULT_DTYPE = os.environ.get("ICF_DTYPE", "fp32")


# ---------------------------------------------------------------------------------------------------
# parameter holders (class names matter: init_weights() keys on names starting with "Conv",
# training_utils.py:114-119)
# ---------------------------------------------------------------------------------------------------
def _uniform_fan_in(weight, bias, fan_in):
    bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
    with torch.no_grad():
        weight.uniform_(-bound, bound)     # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        if bias is not None:
            bias.uniform_(-bound, bound)


class Conv2dParams(nn.Module):
    """weight [Cout,Cin,k,k] + bias [Cout] of an nn.Conv2d (torch default init)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.empty(cout))
        _uniform_fan_in(self.weight, self.bias, cin * k * k)


class ConvTranspose2dParams(nn.Module):
    """weight [Cin,Cout,k,k] + bias [Cout] of an nn.ConvTranspose2d (torch default init: fan_in = Cout*k*k)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cin, cout, k, k))
        self.bias = nn.Parameter(torch.empty(cout))
        _uniform_fan_in(self.weight, self.bias, cout * k * k)


class LinearParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.empty(cout))
        _uniform_fan_in(self.weight, self.bias, cin)


class EmbeddingParams(nn.Module):
    def __init__(self, k, dim=256):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(k, dim))


class BatchNorm2dParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class Slots(nn.Module):
    """Name-space node: children are registered under the reference's Sequential / ModuleDict names."""


def _attach(root: nn.Module, dotted: str, leaf: nn.Module):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, Slots())
        node = node._modules[p]
    node.add_module(parts[-1], leaf)


def _layer_params(l: L) -> nn.Module:
    if l.kind == "conv":
        return Conv2dParams(l.cin, l.cout, l.k)
    if l.kind == "convT":
        return ConvTranspose2dParams(l.cin, l.cout, l.k)
    return LinearParams(l.cin, l.cout)


# ---------------------------------------------------------------------------------------------------
# autograd bridges
# ---------------------------------------------------------------------------------------------------
def _as_image(X, H, W):
    ops.require_cuda(X)
    if X.dtype not in (torch.float32, torch.bfloat16):
        X = X.float()
    X = X.contiguous()
    if X.numel() % (H * W) != 0:
        raise ValueError(f"image tensor of shape {tuple(X.shape)} is not a batch of {H}x{W} images")
    return X, X.numel() // (H * W)


def _as_latent(z, latent):
    ops.require_cuda(z)
    if z.dtype not in (torch.float32, torch.bfloat16):
        z = z.float()
    z = z.contiguous()
    if z.numel() % latent != 0:
        raise ValueError(f"latent tensor of shape {tuple(z.shape)} is not a batch of {latent}-vectors")
    return z, z.numel() // latent


def _to_f32(act: Act, shape):
    out = torch.empty(shape, dtype=torch.float32, device=act.t.device)
    assert act.off == 0 and act.pitch == act.C
    ops.cast(act.ptr, act.code, out.data_ptr(), ops.F32, out.numel())
    return out


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex: NetExec, keys, X, *rest):
        nc = len(keys)
        c = dict(zip(keys, rest[:nc]))
        Xc, N = _as_image(X, ex.H, ex.W)
        save = any(ctx.needs_input_grad)
        out, st = ex.encoder_forward(N, Xc.data_ptr(), ops.code_of(Xc), 1, c, save=save)
        ctx.ex, ctx.st, ctx.nc, ctx.xshape = ex, st, nc, X.shape
        return _to_f32(out, (N, ex.fam.latent, 1, 1))

    @staticmethod
    def backward(ctx, gout):
        ex, st, nc = ctx.ex, ctx.st, ctx.nc
        need_p = any(ctx.needs_input_grad[3 + nc:])
        grads = ex.new_grads() if need_p else None
        g = gout.contiguous().float().reshape(st["N"], ex.fam.latent)
        dX = ex.encoder_backward(st, Act(g, ex.fam.latent), grads, need_dX=ctx.needs_input_grad[2])
        ctx.st = None
        pg = [grads[k] for k, _ in ex.module.named_parameters()] if need_p else [None] * len(ctx.needs_input_grad[3 + nc:])
        return (None, None, dX.reshape(ctx.xshape) if dX is not None else None, *([None] * nc), *pg)


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex: NetExec, keys, z, *rest):
        nc = len(keys)
        c = dict(zip(keys, rest[:nc]))
        zc, N = _as_latent(z, ex.fam.latent)
        save = any(ctx.needs_input_grad)
        out, st = ex.generator_forward(N, zc.data_ptr(), ops.code_of(zc), ex.fam.latent, c, save=save)
        ctx.ex, ctx.st, ctx.nc, ctx.zshape, ctx.keys = ex, st, nc, z.shape, keys
        ctx.cmeta = [(rest[i].shape, rest[i].dtype) for i in range(nc)]
        return _to_f32(out, (N, 1, ex.H, ex.W))

    @staticmethod
    def backward(ctx, gout):
        ex, st, nc = ctx.ex, ctx.st, ctx.nc
        need_p = any(ctx.needs_input_grad[3 + nc:])
        need_attr = any(ctx.needs_input_grad[3:3 + nc])
        grads = ex.new_grads() if need_p else None
        g = gout.contiguous().float().reshape(-1, 1)
        dz, doh, dco = ex.generator_backward(st, Act(g, 1), grads, need_dz=ctx.needs_input_grad[2],
                                             need_dattr=need_attr)
        ctx.st = None
        cg = [None] * nc
        if need_attr:
            cat_names = [a[0] for a in ex.fam.cat_attrs]
            cont_names = ex.cont_names(dict.fromkeys(ctx.keys))
            for i, k in enumerate(ctx.keys):
                if not ctx.needs_input_grad[3 + i]:
                    continue
                shape, dt = ctx.cmeta[i]
                if k in cat_names:
                    cg[i] = doh[cat_names.index(k)].reshape(shape).to(dt)
                elif k in cont_names:
                    cg[i] = dco[cont_names.index(k)].reshape(shape).to(dt)
        pg = [grads[k] for k, _ in ex.module.named_parameters()] if need_p else [None] * len(ctx.needs_input_grad[3 + nc:])
        return (None, None, dz.reshape(ctx.zshape) if dz is not None else None, *cg, *pg)


class _DiscriminatorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex: NetExec, keys, training, masks, X, z, *rest):
        nc = len(keys)
        c = dict(zip(keys, rest[:nc]))
        Xc, N = _as_image(X, ex.H, ex.W)
        zc, Nz = _as_latent(z, ex.fam.latent)
        if N != Nz:
            raise ValueError(f"batch mismatch: {N} images vs {Nz} latent codes")
        save = any(ctx.needs_input_grad)
        logits, st = ex.discriminator_forward(N, Xc.data_ptr(), ops.code_of(Xc), 1, zc.data_ptr(), ops.code_of(zc),
                                              ex.fam.latent, c, masks=masks, training=training, save=save)
        ctx.ex, ctx.st, ctx.nc, ctx.xshape, ctx.zshape = ex, st, nc, X.shape, z.shape
        return logits.t.reshape(N, 1)

    @staticmethod
    def backward(ctx, gout):
        ex, st, nc = ctx.ex, ctx.st, ctx.nc
        need_p = any(ctx.needs_input_grad[6 + nc:])
        grads = ex.new_grads() if need_p else None
        g = gout.contiguous().float().reshape(st["N"], 1)
        dX, dz = ex.discriminator_backward(st, Act(g, 1), grads, need_dX=ctx.needs_input_grad[4],
                                           need_dz=ctx.needs_input_grad[5])
        ctx.st = None
        pg = [grads[k] for k, _ in ex.module.named_parameters()] if need_p else [None] * len(ctx.needs_input_grad[6 + nc:])
        return (None, None, None, None, dX.reshape(ctx.xshape) if dX is not None else None,
                dz.reshape(ctx.zshape) if dz is not None else None, *([None] * nc), *pg)


# ---------------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------------
class _BiGANNet(nn.Module):
    """Common base: parameter construction from the family table + engine cache."""
    FAMILY: str = ""
    ROLE: str = ""
    # class-level defaults: a whole-module pickle written by the REFERENCE (train_mnist_image_scm.py:61-67,
    # train_esrf_bigan.py:31-35) unpickles into these classes without running __init__, i.e. without the two attributes
    compute_dtype: str = DEFAULT_DTYPE
    _exec = None

    def __init__(self):
        super().__init__()
        fam: Family = FAMILIES[self.FAMILY]
        role = self.ROLE
        emb_idx = 3 if role == "G" else 2
        by_name = {a[0]: a for a in fam.cat_attrs}
        for name in fam.emb_decl:                      # registration order of the reference constructor
            a = by_name[name]
            key = a[emb_idx]
            assert key.endswith(".weight")
            _attach(self, key[:-len(".weight")], EmbeddingParams(a[1]))
        towers = {"E": ("E",), "G": ("G",), "D": fam.d_decl}[role]   # mnist.py:98-136 declares dz first, ESRF dx first
        for t in towers:
            for l in getattr(fam, t):
                _attach(self, l.key, _layer_params(l))
                if l.bn:
                    _attach(self, l.bn, BatchNorm2dParams(l.cout))
        self.compute_dtype = DEFAULT_DTYPE
        self._exec = None

    # ---- engine cache (never pickled) -------------------------------------------------------------
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_exec"] = None
        return d

    def set_compute_dtype(self, name: str):
        """'fp32' (default, 1e-3 parity with the reference) or 'bf16' (tensor-core path, 2e-2)."""
        dtype_code(name)
        self.compute_dtype = name
        self._exec = None
        return self

    def engine(self) -> NetExec:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__}: the hot path runs only in the sm_100a CUDA extension; parameters are on "
                f"{dev}. Move the module to a CUDA device (there is no CPU / eager fallback).")
        code = dtype_code(self.compute_dtype)
        ex = self._exec
        if ex is None or ex.device != dev or ex.code != code:
            with torch.cuda.device(dev):
                ex = NetExec(FAMILIES[self.FAMILY], self.ROLE, self, code, dev)
            self._exec = ex
        return ex

    @property
    def device(self):
        return next(self.parameters()).device

    def _params(self):
        return [p for _, p in self.named_parameters()]

    @staticmethod
    def _split(c: Dict[str, torch.Tensor], fam: Family):
        """Keep the attribute entries the network consumes (extra keys such as 'audio' are ignored,
        audio_mnist.py:205-208); MorphoMNIST consumes every key (mnist.py:47-55)."""
        if fam.name == "mnist":
            keys = tuple(c.keys())
        else:
            keys = tuple(a[0] for a in fam.cat_attrs) + tuple(fam.cont_attrs)
            for k in keys:
                if k not in c:
                    raise KeyError(f"attribute {k!r} missing from the attribute dict")
        return keys, [c[k] for k in keys]


class EncoderBase(_BiGANNet):
    ROLE = "E"

    def forward(self, X: torch.Tensor, c: Dict[str, torch.Tensor]):
        ex = self.engine()
        keys, vals = self._split(c, ex.fam)
        with torch.cuda.device(ex.device):
            return _EncoderFn.apply(ex, keys, X, *vals, *self._params())


class GeneratorBase(_BiGANNet):
    ROLE = "G"

    def forward(self, z: torch.Tensor, c: Dict[str, torch.Tensor]):
        ex = self.engine()
        keys, vals = self._split(c, ex.fam)
        with torch.cuda.device(ex.device):
            return _GeneratorFn.apply(ex, keys, z, *vals, *self._params())


class DiscriminatorBase(_BiGANNet):
    ROLE = "D"

    def forward(self, X: torch.Tensor, z: torch.Tensor, c: Dict[str, torch.Tensor], masks=None):
        """``masks`` (optional, testing): the Dropout2d masks of this forward in RNG order; drawn from the
        device's default generator like nn.Dropout2d when omitted."""
        ex = self.engine()
        keys, vals = self._split(c, ex.fam)
        with torch.cuda.device(ex.device):
            return _DiscriminatorFn.apply(ex, keys, self.training, masks, X, z, *vals, *self._params())


def init_weights(layer, std=0.01):
    """training_utils.py:114-119: N(0,std) weights and zero bias for every ``Conv*`` layer."""
    name = layer.__class__.__name__
    if name.startswith("Conv"):
        torch.nn.init.normal_(layer.weight, mean=0, std=std)
        if getattr(layer, "bias", None) is not None:
            torch.nn.init.constant_(layer.bias, 0)
