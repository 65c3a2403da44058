"""Latency of the gradient-based explainer on one GPU (development tool): seconds per explain() call of
``steps`` Adam steps for B images — the fused device path eager and as a replayed CUDA graph, and upstream's formulation
(torch autograd through the drop-in modules, torch.optim.Adam, G's and the classifier's weight gradients included) on the same GPU.

usage: python tools/explain_bench.py [--steps 30] [--batches 1,16,256]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imagecfgen-pytorch_b200"))

import torch  # noqa: E402

from icf_b200 import synth  # noqa: E402
from icf_b200.explain import HingeLossCFExplainer, max_excluding  # noqa: E402
from image_scms import mnist  # noqa: E402


def upstream_style(E, G, clf, x, attrs, target, steps, lr, c=10.0):
    """explain/cf_example.py:96-170 as written upstream (batch 1), on the drop-in modules."""
    codes = E(x, attrs).detach()
    params = {k: (0.01 * torch.randn((1, attrs[k].shape[1]), device=x.device)).requires_grad_(True) for k in attrs
              if k not in ("slant", "intensity")}
    z = torch.randn(codes.shape, device=x.device)
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    for _ in range(steps):
        opt.zero_grad()
        a = {k: (params[k].softmax(1) if k == "digit" else params[k].tanh()) if k in params else attrs[k] for k in attrs}
        x_cf = G(z.tanh(), a)
        pred = clf(x_cf)
        loss = c * (max_excluding(pred, target) - pred[:, target]).mean() + (x - x_cf).abs().mean()
        loss.backward()
        opt.step()
    return x_cf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batches", default="1,16,256")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    E, G = mnist.Encoder().to(dev).set_compute_dtype("bf16"), mnist.Generator().to(dev).set_compute_dtype("bf16")
    clf = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(784, 10)).to(dev)
    ex = HingeLossCFExplainer(E, G, clf, "digit", 512, categorical_features=["digit"], features_to_ignore=["slant", "intensity"])

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    for B in [int(b) for b in args.batches.split(",")]:
        xu, a, _ = synth.mnist_batch(B, 5)
        x, c = synth.mnist_scale(xu, a, synth.mnist_attr_stats())
        x = x.reshape(B, 1, 28, 28).to(dev)
        c = {k: v.to(dev) for k, v in c.items()}
        t_eager = timed(lambda: ex.explain(x, c, target_class=3, steps=args.steps))
        t_graph = timed(lambda: ex.explain(x, c, target_class=3, steps=args.steps, graph=True))
        line = f"B={B:4d} steps={args.steps}: fused eager {1e3 * t_eager:8.1f} ms, fused CUDA graph {1e3 * t_graph:8.1f} ms"
        if B == 1:
            t_up = timed(lambda: upstream_style(E, G, clf, x, c, 3, args.steps, 0.1))
            line += f", upstream formulation through the modules' autograd {1e3 * t_up:8.1f} ms"
        print(line + f"  (best: {1e3 * min(t_eager, t_graph) / B:.2f} ms per image; the graph is captured inside every explain() call)")


if __name__ == "__main__":
    main()
