"""Per-layer micro-benchmark of the conv kernels through the C-ABI (development tool, runs on the GPU box).

For every conv-shaped layer of a family it times forward, dgrad and wgrad in isolation: `--iters` back-to-back
launches between two CUDA events, rotating over enough distinct input/output buffers that the working set
exceeds the 126 MB L2.  Prints one row per (layer, pass): ms, valid-tap TFLOP/s, algorithmic GB/s and the
fraction of max(FLOPs/peak_tensor, bytes/peak_hbm) — the per-layer roofline of SURVEY.md §8(d).

usage: python tools/layer_bench.py [--family mnist] [--batch 4096] [--iters 10] [--only substr] [--json out]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "imagecfgen-pytorch_b200"))

import torch  # noqa: E402

from icf_b200 import ops  # noqa: E402
from icf_b200.arch import FAMILIES, pad8  # noqa: E402
from icf_b200.engine import Act, LayerExec  # noqa: E402


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return d["bf16_tflops"] * 1e12, d["hbm_gbs"] * 1e9
    except Exception:
        return 1.59e15, 6.65e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", default="mnist")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--passes", default="fwd,dgrad,wgrad")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    fam = FAMILIES[a.family]
    N = a.batch
    code = ops.BF16
    tf_peak, hbm_peak = peaks()
    rows = []
    H, W = fam.image
    for tname, specs, h0 in (("E", fam.E, H), ("G", fam.G, 1), ("Dx", fam.Dx, H), ("Dz", fam.Dz, 1), ("Dxz", fam.Dxz, 1)):
        h = w = h0
        for i, sp in enumerate(specs):
            fold = (i == 0 and tname in ("E", "Dx"))
            le = LayerExec(sp, h, w, code, dev, fold=fold)
            name = f"{tname}.{sp.key} {sp.kind} {sp.cin}x{h}x{w}->{sp.cout}x{le.Hout}x{le.Wout} k{sp.k}s{sp.stride}p{sp.pad}"
            hin, win = h, w
            h, w = le.Hout, le.Wout
            if a.only and a.only not in name:
                continue
            wt = torch.randn(le.spec.cout * le.spec.cin * le.taps if sp.kind != "linear" else sp.cout * sp.cin, device=dev) * 0.05
            if sp.kind == "conv":
                wt = wt.view(sp.cout, sp.cin, sp.k, sp.k)
            elif sp.kind == "convT":
                wt = wt.view(sp.cin, sp.cout, sp.k, sp.k)
            else:
                wt = wt.view(sp.cout, sp.cin)
            bias = torch.randn(sp.cout, device=dev) * 0.1
            le.repack(wt, bias)
            fpad = le.pad if le.fold else 0
            in_rows = N * (hin + 2 * fpad) * (win + 2 * fpad) + (win + 2 * fpad if le.fold else 0)
            in_pitch = pad8(le.Cin)
            out_rows = N * le.Hout * le.Wout
            out_pitch = pad8(le.Cout)
            in_bytes, out_bytes = in_rows * in_pitch * 2, out_rows * out_pitch * 2
            nbuf = max(2, min(8, int(300e6 // max(in_bytes + out_bytes, 1)) + 1))
            xs = [Act(torch.randn(in_rows, in_pitch, device=dev).to(torch.bfloat16), le.Cin) for _ in range(nbuf)]
            ys = [Act(torch.empty(out_rows, out_pitch, device=dev, dtype=torch.bfloat16), le.Cout) for _ in range(nbuf)]
            gs = [Act(torch.randn(out_rows, out_pitch, device=dev).to(torch.bfloat16), le.Cout) for _ in range(nbuf)]
            dxs = [Act(torch.empty(N * hin * win, in_pitch, device=dev, dtype=torch.bfloat16), le.Cin) for _ in range(nbuf)]
            gw = torch.zeros_like(wt)
            scratch = torch.empty(le.wgrad_elems, dtype=torch.float32, device=dev)
            flops = le.alg_flops_img * N
            byts = 2.0 * (N * (hin * win * le.Cin + le.Hout * le.Wout * le.Cout) + le.Kout * le.Cin * le.taps)

            def timed(fn):
                for j in range(2):
                    fn(j % nbuf)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for j in range(a.iters):
                    fn(j % nbuf)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / a.iters

            passes = {}
            if "fwd" in a.passes:
                passes["fwd"] = lambda j: le.forward(N, xs[j], ys[j])
            if "dgrad" in a.passes:
                passes["dgrad"] = lambda j: le.dgrad(N, gs[j], dxs[j])
            if "wgrad" in a.passes:
                passes["wgrad"] = lambda j: le.wgrad(N, gs[j], xs[j], gw, scratch)
            for pname, fn in passes.items():
                ms = timed(fn)
                t_bound = max(flops / tf_peak, byts / hbm_peak)
                row = {"layer": name, "pass": pname, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                       "gbs": round(byts / ms / 1e6, 1), "roofline_frac": round(t_bound * 1e3 / ms, 3),
                       "bound": "tensor" if flops / tf_peak > byts / hbm_peak else "hbm"}
                rows.append(row)
                print(f"{name:58s} {pname:5s} {ms:8.4f} ms {row['tflops']:7.1f} TF/s {row['gbs']:7.1f} GB/s "
                      f"frac {row['roofline_frac']:.3f} ({row['bound']})", flush=True)
            del xs, ys, gs, dxs
            torch.cuda.empty_cache()
    tot = sum(r["ms"] for r in rows)
    print(f"sum of single launches: {tot:.3f} ms")
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
