"""Shared helpers of the parity tests."""
import torch

from icf_b200 import synth


def golden_inputs(family, n, seed):
    """Exactly the inputs tests/golden/make_golden.py fed to the reference."""
    if family == "mnist":
        x, a, z = synth.mnist_batch(n, seed)
        stats = synth.mnist_attr_stats()
        images, c = synth.mnist_scale(x, a, stats)
        _, c_cf = synth.mnist_scale(x, synth.intervene_mnist(a), stats)
        return images, c, z, c_cf
    x, c, z = synth.spectro_batch(family, n, seed)
    c_cf = dict(c)
    k = sorted(synth.ATTR_DIMS[family])[0]
    c_cf[k] = torch.roll(c[k], 1, dims=1)
    return x, c, z, c_cf


def rel_err(a, b):
    """||a-b||_2 / ||b||_2 (norm-wise)."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def digest_close(d, ref, tol):
    if d["n"] != ref["n"]:
        return False
    scale = max(ref["l2"], 1e-30)
    return (abs(d["l2"] - ref["l2"]) <= tol * scale and abs(d["sum"] - ref["sum"]) <= tol * max(ref["abs"], 1e-30)
            and abs(d["wsum"] - ref["wsum"]) <= tol * max(ref["abs"], 1e-30))
