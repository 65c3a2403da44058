#!/usr/bin/env python
"""Run-to-run determinism probe of the bf16 train step (run on the GPU box).

Repeats, from identical state and inputs: (1) E / G / D forwards (bitwise), (2) phase-A gradients (max relative
difference per tensor between runs), (3) the whole fused step (the five outputs per run), and prints each
run's outputs against the fp32 oracle and the oracle under the bf16-storage contract, so that a swing between
runs can be attributed to a stage.  usage: python tools/determinism_probe.py [n] [runs] [dtype]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "imagecfgen-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from icf_b200 import synth  # noqa: E402
from icf_b200.trainer import BiGANTrainer  # noqa: E402
from image_scms import mnist  # noqa: E402
from oracle import bigan_ref as R  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
seed, std = 3, 0.05
dev = torch.device("cuda:0")
x, a, z = synth.mnist_batch(n, seed)
stats = synth.mnist_attr_stats()
images, c = synth.mnist_scale(x, a, stats)
sds = {k: R.synth_state_dict("mnist", k, seed, std) for k in "EGD"}
torch.manual_seed(0)
masks6 = [R.draw_masks("mnist", n) for _ in range(6)]
cd = {k: v.to(dev) for k, v in c.items()}
xd, zd = images.to(dev), z.to(dev)
md = [[m.to(dev) for m in ms] for ms in masks6]


def nets():
    out = {}
    for k, cls in (("E", mnist.Encoder), ("G", mnist.Generator), ("D", mnist.Discriminator)):
        m = cls()
        m.load_state_dict(sds[k])
        out[k] = m.to(dev).set_compute_dtype(dtype)
    return out


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---- (1) forwards ------------------------------------------------------------------------------------------
N = nets()
with torch.no_grad():
    ref = None
    for r in range(runs):
        e = N["E"](xd, cd)
        g = N["G"](zd, cd)
        d1 = N["D"](xd, e, cd, masks=md[0])
        d2 = N["D"](g, zd, cd, masks=md[1])
        cur = [t.clone() for t in (e, g, d1, d2)]
        if ref is None:
            ref = cur
        else:
            same = [bool(torch.equal(p, q)) for p, q in zip(cur, ref)]
            diff = [rel(p, q) for p, q in zip(cur, ref)]
            print(f"forward run {r}: bitwise equal E,G,D(x,E),D(G,z) = {same}  rel {['%.1e' % v for v in diff]}")

# ---- (2) phase-A gradients through autograd -----------------------------------------------------------------
bce = torch.nn.BCEWithLogitsLoss()
first = None
for r in range(runs):
    N = nets()
    E, G, D = N["E"], N["G"], N["D"]
    Dv = D(xd, E(xd, cd), cd, masks=md[0])
    Df = D(G(zd, cd), zd, cd, masks=md[1])
    loss = (bce(Dv, torch.zeros(n, 1, device=dev)) + bce(Df, torch.ones(n, 1, device=dev))) / 2
    loss.backward()
    cur = {f"{nm}.{k}": p.grad.clone() for nm, net in N.items() for k, p in net.named_parameters() if p.grad is not None}
    if first is None:
        first = cur
    else:
        worst = sorted(((rel(cur[k], first[k]), k) for k in cur), reverse=True)[:4]
        print(f"grad run {r}: loss {float(loss):.7f}; worst run-to-run rel diffs {[(k, '%.1e' % v) for v, k in worst]}")

# ---- (3) whole step ---------------------------------------------------------------------------------------------
o = R.BiGANOracle("mnist", sds["E"], sds["G"], sds["D"])
want = o.train_step(images, c, z, masks6)
want = [want["loss_EG"], want["loss_D_valid"], want["loss_D_fake"], want["DG_mean"], want["DE_mean"]]
o16 = R.BiGANOracle("mnist", sds["E"], sds["G"], sds["D"], q=R.bf16_storage)
w16 = o16.train_step(images, c, z, masks6)
w16 = [w16["loss_EG"], w16["loss_D_valid"], w16["loss_D_fake"], w16["DG_mean"], w16["DE_mean"]]
print("oracle fp32      ", ["%.6f" % v for v in want])
print("oracle bf16-store", ["%.6f" % v for v in w16], "rel vs fp32", ["%.2e" % (abs(p - q) / max(abs(q), 0.1)) for p, q in zip(w16, want)])
for r in range(runs):
    N = nets()
    tr = BiGANTrainer(N["E"], N["G"], N["D"])
    out = tr.step(xd, cd, zd, md)
    torch.cuda.synchronize()
    got = out[:5].tolist()
    print(f"step run {r}:", ["%.6f" % v for v in got], "rel vs fp32", ["%.2e" % (abs(p - q) / max(abs(q), 0.1)) for p, q in zip(got, want)],
          "vs bf16-store", ["%.2e" % (abs(p - q) / max(abs(q), 0.1)) for p, q in zip(got, w16)])
