"""TEST INFRASTRUCTURE ONLY — architecture tables of the reference BiGAN families.

This file belongs to ``oracle/``: a CPU restatement of the reference's algorithm used as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs. Nothing under ``imagecfgen-pytorch_b200/`` may import it.

Each table restates, op by op, an ``nn.Sequential`` of the reference (file:line cited per table).
Op tuples:
  ("conv",  key, stride, pad)            nn.Conv2d           weight [Cout,Cin,kh,kw]
  ("convT", key, stride, pad, out_pad)   nn.ConvTranspose2d  weight [Cin,Cout,kh,kw]
  ("linear", key)                        nn.Linear           weight [out,in]
  ("unflatten", (C,H,W))                 nn.Unflatten(1, ...)
  ("lrelu", slope) / ("tanh",)
  ("bn", key)                            nn.BatchNorm2d (eps 1e-5, momentum 0.1, affine)
  ("drop", p)                            nn.Dropout2d(p)
``key`` is the state_dict prefix inside the owning module (e.g. "layers.0", "dx.4").
"""


def _conv_stack(prefix, n, stride, pad, slope, last_act=False, start=0):
    """[conv, lrelu]*n with Sequential indices start, start+2, ... (no act after the last unless asked)."""
    ops = []
    for i in range(n):
        ops.append(("conv", f"{prefix}.{start + 2 * i}", stride, pad))
        if i < n - 1 or last_act:
            ops.append(("lrelu", slope))
    return ops


# --------------------------------------------------------------------------------------------
# MorphoMNIST — image_scms/mnist.py
# --------------------------------------------------------------------------------------------
MNIST = {
    "name": "mnist",
    "image": (28, 28),
    "latent": 512,
    # Encoder.layers, mnist.py:30-40
    "E": [("conv", "layers.0", 2, 1), ("lrelu", 0.2),
          ("conv", "layers.2", 2, 1), ("lrelu", 0.2),
          ("conv", "layers.4", 2, 1), ("lrelu", 0.2),
          ("conv", "layers.6", 2, 1), ("lrelu", 0.2),
          ("conv", "layers.8", 2, 0)],
    # Generator.layers, mnist.py:63-74
    "G": [("convT", "layers.0", 1, 0, 0), ("lrelu", 0.2),
          ("convT", "layers.2", 2, 0, 0), ("lrelu", 0.2),
          ("convT", "layers.4", 2, 1, 0), ("lrelu", 0.2),
          ("convT", "layers.6", 2, 1, 0), ("lrelu", 0.2),
          ("convT", "layers.8", 1, 0, 0), ("tanh",)],
    # Discriminator.dx, mnist.py:106-126
    "Dx": [("drop", 0.2), ("conv", "dx.1", 1, 0), ("lrelu", 0.1), ("drop", 0.2), ("bn", "dx.4"),
           ("conv", "dx.5", 2, 0), ("lrelu", 0.1), ("bn", "dx.7"), ("drop", 0.5),
           ("conv", "dx.9", 1, 0), ("lrelu", 0.1), ("bn", "dx.11"), ("drop", 0.5),
           ("conv", "dx.13", 2, 0), ("lrelu", 0.1), ("bn", "dx.15"), ("drop", 0.5),
           ("conv", "dx.17", 1, 0), ("lrelu", 0.1)],
    # Discriminator.dz, mnist.py:98-105
    "Dz": [("drop", 0.2), ("conv", "dz.1", 1, 0), ("lrelu", 0.1),
           ("drop", 0.5), ("conv", "dz.4", 1, 0), ("lrelu", 0.1)],
    # Discriminator.dxz, mnist.py:127-136
    "Dxz": [("drop", 0.2), ("conv", "dxz.1", 1, 0), ("lrelu", 0.1),
            ("drop", 0.2), ("conv", "dxz.4", 1, 0), ("lrelu", 0.1),
            ("drop", 0.2), ("conv", "dxz.7", 1, 0)],
    "init_std": 0.01,              # training_utils.py:114
    "adam_betas": (0.5, 0.999),    # mnist.py:176-179
}


def _spectro_family(name, image, n_enc, dx_keys_same_as_E, n_gen_convT, init_std=0.001):
    """Spectrogram families share one shape: k5 s2 p1 conv towers, Linear + k5 s2 p2 op1 ConvT towers.

    audio_mnist.py:187-197 / :224-243 / :272-303; whalecalls.py:245-257 / :286-307 / :340-371;
    esrf_acoustic.py:144-160 / :181-198 / :218-247.
    """
    E = _conv_stack("layers", n_enc, 2, 1, 0.2)
    Dx = _conv_stack("dx", n_enc, 2, 1, 0.2)
    G = [("linear", "layers.0"), ("unflatten", (1024, 4, 4)), ("lrelu", 0.2)]
    for i in range(n_gen_convT):
        G.append(("convT", f"layers.{3 + 2 * i}", 2, 2, 1))
        G.append(("lrelu", 0.2) if i < n_gen_convT - 1 else ("tanh",))
    Dz = [("conv", "dz.0", 1, 0), ("lrelu", 0.2), ("conv", "dz.2", 1, 0), ("lrelu", 0.2)]
    Dxz = [("conv", "dxz.0", 1, 0), ("lrelu", 0.2), ("conv", "dxz.2", 1, 0), ("lrelu", 0.2),
           ("conv", "dxz.4", 1, 0)]
    return {"name": name, "image": image, "latent": 512, "E": E, "G": G, "Dx": Dx, "Dz": Dz,
            "Dxz": Dxz, "init_std": init_std, "adam_betas": (0.5, 0.9)}


AUDIO_MNIST = _spectro_family("audio_mnist", (128, 128), 6, True, 5)
# audio_mnist.py:23-30 (insertion order = embedding_dict order; forward uses sorted() order, :205-208)
AUDIO_MNIST["attribute_dims"] = {"country_of_origin": 13, "native_speaker": 2, "accent": 15,
                                 "digit": 10, "age": 5, "gender": 2}
AUDIO_MNIST["upsample"] = 8

WHALE = _spectro_family("whalecalls", (256, 256), 7, True, 6)
WHALE["attribute_dims"] = {"call_type": 3}    # whalecalls.py:30-36 minus "time"/"path"
WHALE["upsample"] = 16

ESRF = _spectro_family("esrf_acoustic", (512, 512), 8, True, 7)
ESRF["attribute_dims"] = {"closest_boat": 1, "has_boat": 2}   # esrf_acoustic.py:17-20
ESRF["upsample"] = 32

FAMILIES = {"mnist": MNIST, "audio_mnist": AUDIO_MNIST, "whalecalls": WHALE, "esrf_acoustic": ESRF}


# Parameter shapes, in state_dict order. Used to synthesise weights without
# importing either implementation.
def _conv_shapes(prefix, chans, k, start=0, transposed=False):
    out = []
    for i in range(len(chans) - 1):
        cin, cout = chans[i], chans[i + 1]
        w = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        out.append((f"{prefix}.{start + 2 * i}.weight", w))
        out.append((f"{prefix}.{start + 2 * i}.bias", (cout,)))
    return out


def param_shapes(family, net):
    """List of (state_dict key, shape) for floating-point parameters AND BN buffers of one network."""
    if family == "mnist":
        if net == "E":
            return [("digit_embedding.0.weight", (10, 256)),
                    ("layers.0.weight", (64, 5, 3, 3)), ("layers.0.bias", (64,)),
                    ("layers.2.weight", (128, 64, 4, 4)), ("layers.2.bias", (128,)),
                    ("layers.4.weight", (256, 128, 4, 4)), ("layers.4.bias", (256,)),
                    ("layers.6.weight", (512, 256, 4, 4)), ("layers.6.bias", (512,)),
                    ("layers.8.weight", (512, 512, 1, 1)), ("layers.8.bias", (512,))]
        if net == "G":
            return [("digit_embedding.weight", (10, 256)),
                    ("layers.0.weight", (771, 512, 3, 3)), ("layers.0.bias", (512,)),
                    ("layers.2.weight", (512, 256, 3, 3)), ("layers.2.bias", (256,)),
                    ("layers.4.weight", (256, 128, 3, 3)), ("layers.4.bias", (128,)),
                    ("layers.6.weight", (128, 64, 3, 3)), ("layers.6.bias", (64,)),
                    ("layers.8.weight", (64, 1, 4, 4)), ("layers.8.bias", (1,))]
        if net == "D":
            out = [("digit_embedding.0.weight", (10, 256)),
                   ("dz.1.weight", (512, 512, 1, 1)), ("dz.1.bias", (512,)),
                   ("dz.4.weight", (512, 512, 1, 1)), ("dz.4.bias", (512,)),
                   ("dx.1.weight", (32, 5, 5, 5)), ("dx.1.bias", (32,))]

            def bn(key, c):
                return [(f"{key}.weight", (c,)), (f"{key}.bias", (c,)),
                        (f"{key}.running_mean", (c,)), (f"{key}.running_var", (c,))]
            out += bn("dx.4", 32)
            out += [("dx.5.weight", (64, 32, 4, 4)), ("dx.5.bias", (64,))] + bn("dx.7", 64)
            out += [("dx.9.weight", (128, 64, 4, 4)), ("dx.9.bias", (128,))] + bn("dx.11", 128)
            out += [("dx.13.weight", (256, 128, 4, 4)), ("dx.13.bias", (256,))] + bn("dx.15", 256)
            out += [("dx.17.weight", (512, 256, 3, 3)), ("dx.17.bias", (512,)),
                    ("dxz.1.weight", (1024, 1024, 1, 1)), ("dxz.1.bias", (1024,)),
                    ("dxz.4.weight", (1024, 1024, 1, 1)), ("dxz.4.bias", (1024,)),
                    ("dxz.7.weight", (1, 1024, 1, 1)), ("dxz.7.bias", (1,))]
            return out
    fam = FAMILIES[family]
    d = 64
    enc = {"audio_mnist": [None, d, 2 * d, 4 * d, 8 * d, 16 * d, 512],
           "whalecalls": [None, d, 2 * d, 4 * d, 8 * d, 16 * d, 16 * d, 512],
           "esrf_acoustic": [None, d, 2 * d, 4 * d, 8 * d, 16 * d, 32 * d, 64 * d, 512]}[family]
    dxc = {"audio_mnist": enc, "whalecalls": [None, d, 2 * d, 2 * d, 4 * d, 8 * d, 16 * d, 512],
           "esrf_acoustic": enc}[family]
    gen = {"audio_mnist": [16 * d, 8 * d, 4 * d, 2 * d, d, 1],
           "whalecalls": [16 * d, 16 * d, 8 * d, 4 * d, 2 * d, d, 1],
           "esrf_acoustic": [16 * d, 16 * d, 8 * d, 4 * d, 2 * d, d, d, 1]}[family]
    if family == "esrf_acoustic":
        cin0, gin = 3, 512 + 257
        emb_ed = [("has_boat_embedding.0.weight", (2, 256))]
        emb_g = [("has_boat_embedding.weight", (2, 256))]
    else:
        dims = fam["attribute_dims"]
        cin0, gin = 1 + len(dims), 512 + 256 * len(dims)
        emb_ed = [(f"embedding_dict.{k}.0.weight", (v, 256)) for k, v in dims.items()]
        emb_g = [(f"embedding_dict.{k}.weight", (v, 256)) for k, v in dims.items()]
    if net == "E":
        return emb_ed + _conv_shapes("layers", [cin0] + enc[1:], 5)
    if net == "G":
        return (emb_g + [("layers.0.weight", (16384, gin)), ("layers.0.bias", (16384,))]
                + _conv_shapes("layers", gen, 5, start=3, transposed=True))
    if net == "D":
        dz = _conv_shapes("dz", [512, 512, 512], 1)
        dx = _conv_shapes("dx", [cin0] + dxc[1:], 5)
        # registration order of the reference constructors: dz, dx, dxz (audio_mnist.py:272-297, whalecalls.py:338-365)
        # except ESRF, which declares dx first (esrf_acoustic.py:218-241)
        towers = dx + dz if family == "esrf_acoustic" else dz + dx
        return emb_ed + towers + _conv_shapes("dxz", [1024, 1024, 1024, 1], 1)
    raise KeyError((family, net))
