"""Micro-benchmark of the streaming BatchNorm / activation kernels through the C-ABI (development tool, GPU box).

Shapes are the BatchNorm'd layers of the MorphoMNIST step at batch 4096 (channels, pixels per image).  Each kernel
runs `--iters` times over two alternating buffer sets (the working set of one launch already exceeds the 126 MB L2
for the big shapes) between two CUDA events.  Prints ms, algorithmic GB/s and the fraction of the measured HBM peak.

usage: python tools/ew_bench.py [--batch 4096] [--iters 10]      (tuning aids: ICF_EW_VU, ICF_EW_CAP)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imagecfgen-pytorch_b200"))

import torch  # noqa: E402

from icf_b200 import ops  # noqa: E402

SHAPES = [(32, 576), (64, 121), (128, 64), (256, 9), (64, 625), (128, 169), (256, 49), (64, 196), (128, 49)]


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
    except Exception:
        return 6.65e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = hbm_peak()
    N = a.batch
    total = 0.0
    for C, pps in SHAPES:
        pix = N * pps
        sets = []
        for _ in range(2):
            y = torch.randn(pix, C, device=dev).to(torch.bfloat16)
            g = torch.randn(pix, C, device=dev).to(torch.bfloat16)
            o = torch.empty(pix, C, device=dev, dtype=torch.bfloat16)
            sets.append((y, g, o))
        mask = (torch.rand(N, C, device=dev) > 0.2).float() * 1.25
        scale = torch.rand(C, device=dev) + 0.5
        shift = torch.randn(C, device=dev)
        mean = torch.randn(C, device=dev)
        invstd = torch.rand(C, device=dev) + 0.5
        gamma = torch.rand(C, device=dev) + 0.5
        sums = torch.zeros(2 * C, device=dev)
        dgamma = torch.zeros(C, device=dev)
        dbeta = torch.zeros(C, device=dev)
        P = ops.ptr

        def ssm(i):
            y, g, o = sets[i & 1]
            ops.scale_shift_mask(P(y), ops.BF16, C, P(o), ops.BF16, C, pix, pps, C, P(scale), P(shift), P(mask), C)

        def bbr(i):
            y, g, o = sets[i & 1]
            ops.bn_bwd_reduce(P(g), ops.BF16, C, P(y), ops.BF16, C, pix, pps, C, P(mask), C, P(mean), P(invstd), P(sums))

        def actb(i):
            y, g, o = sets[i & 1]
            ops.act_backward(P(g), ops.BF16, C, P(y), ops.BF16, C, P(o), ops.BF16, C, pix, pps, C, "lrelu", 0.1,
                             bn_sums=P(sums), bn_mask=P(mask), bn_mask_pitch=C, bn_gamma=P(gamma), bn_mean=P(mean),
                             bn_invstd=P(invstd), bn_dgamma=P(dgamma), bn_dbeta=P(dbeta))

        for name, fn, nb in (("scale_shift_mask", ssm, 2), ("bn_bwd_reduce", bbr, 2), ("act_backward", actb, 3)):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            nbytes = pix * C * 2 * nb
            total += ms
            print(f"{name:18s} C{C:<4d} pps{pps:<4d} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  frac {nbytes / (ms * 1e-3) / peak:.3f}")
    print(f"sum {total:.3f} ms")


if __name__ == "__main__":
    main()
