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


def lrelu_signs(ex, tower, saved, n):
    """Per LeakyReLU layer of one tower: bool (N,C,H,W) tensor (CPU), True where the CUDA forward's stored activation is
    positive — the branch its backward differentiates (icf_common.cuh act_grad_from_output).  ``saved`` is the list of
    per-layer records the engine keeps for the backward pass."""
    out = []
    for le, rec in zip(ex.towers[tower].layers, saved):
        if le.spec.act != "lrelu":
            continue
        y = rec["y"]
        t = y.t[:, y.off:y.off + le.Cout].float().cpu()
        out.append((t.reshape(n, le.Hout, le.Wout, le.Cout).permute(0, 3, 1, 2) > 0).contiguous())
    return out
