#!/usr/bin/env python
"""Layer-by-layer determinism probe of one Discriminator forward (run on the GPU box): repeats the same forward on fixed
inputs / masks and reports, per saved tensor, whether it is bitwise reproducible and where the first differences sit.
usage: python tools/determinism_probe2.py [n] [runs] [which: fake|real]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "imagecfgen-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from icf_b200 import ops, synth  # noqa: E402
from image_scms import mnist  # noqa: E402
from oracle import bigan_ref as R  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
which = sys.argv[3] if len(sys.argv) > 3 else "fake"
seed, std = 3, 0.05
dev = torch.device("cuda:0")
x, a, z = synth.mnist_batch(n, seed)
images, c = synth.mnist_scale(x, a, synth.mnist_attr_stats())
sds = {k: R.synth_state_dict("mnist", k, seed, std) for k in "EGD"}
torch.manual_seed(0)
masks6 = [R.draw_masks("mnist", n) for _ in range(6)]
cd = {k: v.to(dev) for k, v in c.items()}
xd, zd = images.to(dev), z.to(dev)
md = [[m.to(dev) for m in ms] for ms in masks6]
N = {}
for k, cls in (("E", mnist.Encoder), ("G", mnist.Generator), ("D", mnist.Discriminator)):
    m = cls()
    m.load_state_dict(sds[k])
    N[k] = m.to(dev).set_compute_dtype("bf16")
with torch.no_grad():
    e = N["E"](xd, cd).clone()
    g = N["G"](zd, cd).clone()
ex = N["D"].engine()
X, Z, M = (g, zd, md[1]) if which == "fake" else (xd, e, md[0])


def snapshot():
    logits, st = ex.discriminator_forward(n, X.data_ptr(), ops.code_of(X), 1, Z.data_ptr(), ops.code_of(Z), 512, cd, masks=M,
                                          training=True, save=True)
    torch.cuda.synchronize()
    out = {}
    for t in ("Dx", "Dz", "Dxz"):
        for i, rec in enumerate(st[t]):
            out[f"{t}.{i}.x"] = rec["x"].t.clone()
            out[f"{t}.{i}.y"] = rec["y"].t.clone()
            if rec["bn"] is not None:
                out[f"{t}.{i}.bn_ss"] = rec["bn"]["ss"].clone()
    out["logits"] = logits.t.clone()
    return out


ref = snapshot()
for r in range(1, runs):
    cur = snapshot()
    bad = []
    for k in ref:
        if not torch.equal(cur[k], ref[k]):
            d = (cur[k].float() - ref[k].float())
            idx = d.nonzero()
            bad.append((k, tuple(ref[k].shape), int(idx.shape[0]), idx[:6].tolist(), float(d.abs().max()), float(ref[k].float().abs().max())))
    print(f"run {r}: {len(bad)} tensors differ")
    for b in bad[:8]:
        print("   ", b)
