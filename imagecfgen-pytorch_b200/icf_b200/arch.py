"""Layer programs of the four conditional-BiGAN families, in the *fused* form the B200 engine executes.

One ``L`` = one convolution-shaped contraction together with everything the engine fuses around it:
the Dropout2d in front of it, the activation + Dropout2d in its epilogue, and the BatchNorm2d (+Dropout2d)
that follows.  ``key`` is the reference's ``state_dict`` prefix so checkpoints stay interchangeable
(SURVEY.md App. A.5).  Geometry sources: image_scms/mnist.py:21-136, audio_mnist.py:173-303,
whalecalls.py:230-371, esrf_acoustic.py:134-247 of the reference.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class L:
    key: str
    kind: str                     # "conv" | "convT" | "linear"
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    pad: int = 0
    opad: int = 0
    act: str = "none"             # "none" | "lrelu" | "tanh"
    slope: float = 0.0
    in_drop: float = 0.0          # Dropout2d on the tensor this layer reads
    out_drop: float = 0.0         # Dropout2d right after the activation
    bn: Optional[str] = None      # BatchNorm2d (state_dict prefix) on the layer output
    bn_drop: float = 0.0          # Dropout2d after that BatchNorm
    unflatten: Optional[Tuple[int, int, int]] = None   # linear only: (C, H, W) of nn.Unflatten


@dataclass(frozen=True)
class Family:
    name: str
    image: Tuple[int, int]
    latent: int
    # categorical attributes rendered as tanh(upsampled embedding) planes, in channel order, with the
    # state_dict prefix of their table in E/D and in G; continuous attributes rendered as constant planes
    cat_attrs: Tuple[Tuple[str, int, str, str], ...]      # (name, K, E/D table key, G table key)
    cont_attrs: Tuple[str, ...]
    E: Tuple[L, ...]
    G: Tuple[L, ...]
    Dx: Tuple[L, ...]
    Dz: Tuple[L, ...]
    Dxz: Tuple[L, ...]
    init_std: float
    adam_betas: Tuple[float, float]
    default_batch: int
    emb_decl: Tuple[str, ...] = ()     # order in which the reference constructor registers the embeddings
    d_decl: Tuple[str, ...] = ("Dz", "Dx", "Dxz")   # order in which the reference Discriminator registers its towers


def _lr(s):
    return dict(act="lrelu", slope=s)


MNIST = Family(
    name="mnist", image=(28, 28), latent=512,
    cat_attrs=(("digit", 10, "digit_embedding.0.weight", "digit_embedding.weight"),),
    cont_attrs=("intensity", "slant", "thickness"),       # sorted(keys != "digit"), mnist.py:53-55
    E=(L("layers.0", "conv", 5, 64, 3, 2, 1, **_lr(0.2)),
       L("layers.2", "conv", 64, 128, 4, 2, 1, **_lr(0.2)),
       L("layers.4", "conv", 128, 256, 4, 2, 1, **_lr(0.2)),
       L("layers.6", "conv", 256, 512, 4, 2, 1, **_lr(0.2)),
       L("layers.8", "conv", 512, 512, 1, 2, 0)),
    G=(L("layers.0", "convT", 771, 512, 3, 1, 0, **_lr(0.2)),
       L("layers.2", "convT", 512, 256, 3, 2, 0, **_lr(0.2)),
       L("layers.4", "convT", 256, 128, 3, 2, 1, **_lr(0.2)),
       L("layers.6", "convT", 128, 64, 3, 2, 1, **_lr(0.2)),
       L("layers.8", "convT", 64, 1, 4, 1, 0, act="tanh")),
    Dx=(L("dx.1", "conv", 5, 32, 5, 1, 0, **_lr(0.1), in_drop=0.2, out_drop=0.2, bn="dx.4"),
        L("dx.5", "conv", 32, 64, 4, 2, 0, **_lr(0.1), bn="dx.7", bn_drop=0.5),
        L("dx.9", "conv", 64, 128, 4, 1, 0, **_lr(0.1), bn="dx.11", bn_drop=0.5),
        L("dx.13", "conv", 128, 256, 4, 2, 0, **_lr(0.1), bn="dx.15", bn_drop=0.5),
        L("dx.17", "conv", 256, 512, 3, 1, 0, **_lr(0.1))),
    Dz=(L("dz.1", "conv", 512, 512, 1, 1, 0, **_lr(0.1), in_drop=0.2, out_drop=0.5),
        L("dz.4", "conv", 512, 512, 1, 1, 0, **_lr(0.1))),
    Dxz=(L("dxz.1", "conv", 1024, 1024, 1, 1, 0, **_lr(0.1), in_drop=0.2, out_drop=0.2),
         L("dxz.4", "conv", 1024, 1024, 1, 1, 0, **_lr(0.1), out_drop=0.2),
         L("dxz.7", "conv", 1024, 1, 1, 1, 0)),
    init_std=0.01, adam_betas=(0.5, 0.999), default_batch=64, emb_decl=("digit",))


def _tower(prefix, chans, start=0, last_act=False):
    out = []
    for i in range(len(chans) - 1):
        act = _lr(0.2) if (i < len(chans) - 2 or last_act) else {}
        out.append(L(f"{prefix}.{start + 2 * i}", "conv", chans[i], chans[i + 1], 5, 2, 1, **act))
    return tuple(out)


def _gen(in_dim, chans):
    out = [L("layers.0", "linear", in_dim, 16384, act="lrelu", slope=0.2, unflatten=(1024, 4, 4))]
    for i in range(len(chans) - 1):
        last = i == len(chans) - 2
        out.append(L(f"layers.{3 + 2 * i}", "convT", chans[i], chans[i + 1], 5, 2, 2, 1,
                     **({"act": "tanh"} if last else _lr(0.2))))
    return tuple(out)


_DZ = (L("dz.0", "conv", 512, 512, 1, 1, 0, **_lr(0.2)), L("dz.2", "conv", 512, 512, 1, 1, 0, **_lr(0.2)))
_DXZ = (L("dxz.0", "conv", 1024, 1024, 1, 1, 0, **_lr(0.2)), L("dxz.2", "conv", 1024, 1024, 1, 1, 0, **_lr(0.2)),
        L("dxz.4", "conv", 1024, 1, 1, 1, 0))
_d = 64

_AUDIO_ATTRS = {"country_of_origin": 13, "native_speaker": 2, "accent": 15, "digit": 10, "age": 5, "gender": 2}
AUDIO_MNIST = Family(
    name="audio_mnist", image=(128, 128), latent=512,
    cat_attrs=tuple((k, _AUDIO_ATTRS[k], f"embedding_dict.{k}.0.weight", f"embedding_dict.{k}.weight")
                    for k in sorted(_AUDIO_ATTRS)),       # forward uses sorted(ATTRIBUTE_DIMS), audio_mnist.py:205
    cont_attrs=(),
    E=_tower("layers", [7, _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 512]),
    G=_gen(512 + 256 * 6, [16 * _d, 8 * _d, 4 * _d, 2 * _d, _d, 1]),
    Dx=_tower("dx", [7, _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 512]),
    Dz=_DZ, Dxz=_DXZ, init_std=0.001, adam_betas=(0.5, 0.9), default_batch=128, emb_decl=tuple(_AUDIO_ATTRS))

WHALE = Family(
    name="whalecalls", image=(256, 256), latent=512,
    cat_attrs=(("call_type", 3, "embedding_dict.call_type.0.weight", "embedding_dict.call_type.weight"),),
    cont_attrs=(),
    E=_tower("layers", [2, _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 16 * _d, 512]),
    G=_gen(512 + 256, [16 * _d, 16 * _d, 8 * _d, 4 * _d, 2 * _d, _d, 1]),
    Dx=_tower("dx", [2, _d, 2 * _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 512]),
    Dz=_DZ, Dxz=_DXZ, init_std=0.001, adam_betas=(0.5, 0.9), default_batch=32, emb_decl=("call_type",))

ESRF = Family(
    name="esrf_acoustic", image=(512, 512), latent=512,
    cat_attrs=(("has_boat", 2, "has_boat_embedding.0.weight", "has_boat_embedding.weight"),),
    cont_attrs=("closest_boat",),
    E=_tower("layers", [3, _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 32 * _d, 64 * _d, 512]),
    G=_gen(512 + 257, [16 * _d, 16 * _d, 8 * _d, 4 * _d, 2 * _d, _d, _d, 1]),
    Dx=_tower("dx", [3, _d, 2 * _d, 4 * _d, 8 * _d, 16 * _d, 32 * _d, 64 * _d, 512]),
    Dz=_DZ, Dxz=_DXZ, init_std=0.001, adam_betas=(0.5, 0.9), default_batch=64, emb_decl=("has_boat",),
    d_decl=("Dx", "Dz", "Dxz"))                            # esrf_acoustic.py:218-241 declares dx before dz

FAMILIES: Dict[str, Family] = {f.name: f for f in (MNIST, AUDIO_MNIST, WHALE, ESRF)}


def conv_out(h, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1


def convT_out(h, k, stride, pad, opad):
    return (h - 1) * stride - 2 * pad + k + opad


def pad8(c):
    return (c + 7) // 8 * 8


def forward_flops_per_image(family: Family) -> Dict[str, float]:
    """2 x valid-tap MACs of one forward pass per network (SURVEY.md §8(d) definition)."""
    def taps_1d(h_in, h_out, k, s, p, transposed):
        if not transposed:   # output positions x kernel taps that land inside the un-padded input
            return sum(1 for o in range(h_out) for r in range(k) if 0 <= o * s - p + r < h_in)
        return sum(1 for i in range(h_in) for r in range(k) if 0 <= i * s - p + r < h_out)

    def tower(layers: Tuple[L, ...], h):
        macs = 0
        for l in layers:
            if l.kind == "linear":
                macs += l.cin * l.cout
                h = l.unflatten[1]
            elif l.kind == "conv":
                ho = conv_out(h, l.k, l.stride, l.pad)
                t = taps_1d(h, ho, l.k, l.stride, l.pad, False)
                macs += l.cin * l.cout * t * t
                h = ho
            else:
                ho = convT_out(h, l.k, l.stride, l.pad, l.opad)
                t = taps_1d(h, ho, l.k, l.stride, l.pad, True)
                macs += l.cin * l.cout * t * t
                h = ho
        return macs
    H = family.image[0]
    fe = tower(family.E, H)
    fg = tower(family.G, 1)
    fd = tower(family.Dx, H) + tower(family.Dz, 1) + tower(family.Dxz, 1)
    return {"E": 2.0 * fe, "G": 2.0 * fg, "D": 2.0 * fd}
