#!/usr/bin/env python
"""bench.py — BiGAN train-step images/s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): MorphoMNIST conditional BiGAN train step (phases A-D of
image_scms/mnist.py:220-248), batch 4096 per GPU, bf16 activations / fp32 accumulation, synthetic
MorphoMNIST-shaped data, weights from init_weights (std 0.01).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "imagecfgen-pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--family", default="mnist")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=128)
    ap.add_argument("--cf-batch", type=int, default=65536)
    ap.add_argument("--skip-cf", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sust": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sust": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# -----------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port of image_scms/mnist.py:220-248)
# -----------------------------------------------------------------------------------------------------------
def cpu_reference_rate(family, cpu_batch, steps, warmup, seed=42):
    """images/s of the oracle's train step (same torch CPU kernels the reference dispatches to) on all host cores."""
    from icf_b200 import synth
    from oracle import bigan_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds = {k: R.synth_state_dict(family, k, seed, 0.01) for k in "EGD"}
    o = R.BiGANOracle(family, sds["E"], sds["G"], sds["D"])
    x, a, z = synth.mnist_batch(cpu_batch, seed)
    images, c = synth.mnist_scale(x, a, synth.mnist_attr_stats())
    times = []
    for i in range(warmup + steps):
        masks6 = [R.draw_masks(family, cpu_batch) for _ in range(6)]
        t0 = time.perf_counter()
        o.train_step(images, c, z, masks6)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return cpu_batch / sec, sec, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    rate, sec, cores = cpu_reference_rate(args.family, args.cpu_batch, steps, warm)
    line = {"impl": "reference", "metric": "bigan_train_step_images_per_s", "value": rate, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"MorphoMNIST conditional BiGAN train step, batch {args.batch} per GPU "
                                   f"(image_scms/mnist.py:220-248)", "family": args.family,
                       "batch_per_gpu": args.batch},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} steps of {args.cpu_batch} images of the batch (oracle/bigan_ref.py, "
                                       f"torch {torch.__version__} CPU kernels, fp32)"},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------
def run_ours(args):
    if os.environ.get("ICF_WATCHDOG"):          # development aid: dump every thread's stack and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["ICF_WATCHDOG"]), exit=True)
    import torch.distributed as dist
    from icf_b200 import ops, synth
    from icf_b200.arch import FAMILIES, forward_flops_per_image
    from icf_b200.trainer import BiGANTrainer, counterfactual
    from image_scms import mnist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    B = args.batch
    torch.manual_seed(42)
    E, G, D = mnist.Encoder().to(dev), mnist.Generator().to(dev), mnist.Discriminator().to(dev)
    for m in (E, G, D):
        m.apply(mnist.init_weights)
    if world > 1:                                     # identical replicas: broadcast rank 0's weights
        for m in (E, G, D):
            for t in list(m.parameters()) + list(m.buffers()):
                dist.broadcast(t.data, 0)
    tr = BiGANTrainer(E, G, D, lr=1e-4, dtype=args.dtype, process_group=pg)
    # synthetic MorphoMNIST-shaped batch, pinned host copies for the end-to-end leg
    x, a, z = synth.mnist_batch(B, 42 + rank)
    stats = synth.mnist_attr_stats()
    h_x = x.pin_memory()
    h_a = {k: v.pin_memory() for k, v in a.items()}
    cont = sorted(k for k in a if k != "digit")
    lo = {k: stats[k][0].to(dev) for k in cont}
    span = {k: (stats[k][1] - stats[k][0]).to(dev) for k in cont}
    d_x = torch.empty((B, 28, 28), device=dev)
    d_a = {k: torch.empty_like(v, device=dev) for k, v in a.items()}

    def upload_and_scale():
        """host->device copy of the raw batch + the scaling of mnist.py:204-209 (into the static step inputs)."""
        d_x.copy_(h_x, non_blocking=True)
        for k in d_a:
            d_a[k].copy_(h_a[k], non_blocking=True)
        images = 2 * d_x.reshape(-1, 1, 28, 28) / 255 - 1
        c = {k: 2 * (d_a[k] - lo[k]) / span[k] - 1 for k in cont}
        c["digit"] = d_a["digit"]
        return images, c

    images, c = upload_and_scale()
    h2d = h_x.numel() * 4 + sum(v.numel() * 4 for v in h_a.values())
    use_graph = not args.no_graph
    n0 = ops.LAUNCHES
    tr.step(images, c)                                 # eager step: warms allocator, counts launches
    launches_per_step = ops.LAUNCHES - n0
    torch.cuda.synchronize()
    if use_graph:
        tr.capture(images, c, warmup=1)

    def one_step(imgs=None, cc=None):
        if use_graph:
            return tr.replay(imgs, cc)
        return tr.step(images if imgs is None else imgs, c if cc is None else cc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(3, args.warmup)):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step = timed(one_step, args.steps)              # inputs resident in HBM
    h_out = torch.empty(8).pin_memory()

    def e2e_step():
        imgs, cc = upload_and_scale()
        out = one_step(imgs, cc)
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the step's result (losses, scores) is read on the host

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing (eager replica of the step with CUDA events around every C-ABI launch) ----------
    roof = None
    kern = {}
    if world > 1 and rank != 0:
        tr.step(images, c)                             # the profiled replica runs collectives: every rank takes part
        torch.cuda.synchronize()
    if rank == 0:
        ops.PROFILE = []
        tr.step(images, c)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        layers = {}
        for name, e0, e1, fl, nb, det in prof:
            ms = e0.elapsed_time(e1)
            k = kern.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
            k["ms"] += ms
            k["n"] += 1
            k["flops"] += fl
            k["bytes"] += nb
            if det:
                d = layers.setdefault(det, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
                d["ms"] += ms
                d["n"] += 1
                d["flops"] += fl
                d["bytes"] += nb
        pk0 = peaks()
        per_layer = []
        for det, d in sorted(layers.items(), key=lambda kv: -kv[1]["ms"]):
            sec = d["ms"] * 1e-3
            t_roof = max(d["flops"] / (pk0["tf_sust"] * 1e12), d["bytes"] / (pk0["hbm"] * 1e9))
            per_layer.append({"layer": det, "n": d["n"], "ms": round(d["ms"], 4),
                              "tflops": round(d["flops"] / sec / 1e12, 2), "gbs": round(d["bytes"] / sec / 1e9, 1),
                              "roofline_frac": round(t_roof / sec, 4)})
        with open(os.path.join(ROOT, "gpurun_out", "per_layer.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else os.devnull, "w") as f:
            json.dump(per_layer, f, indent=1)
        total = sum(k["ms"] for k in kern.values())
        pk = peaks()
        conv = {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0}
        for nm in ("icf_conv_forward", "icf_conv_wgrad"):
            for f in conv:
                conv[f] += kern.get(nm, conv)[f] if nm in kern else 0
        if conv["ms"] > 0:
            ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
            peak = pk["tf_sust"]
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "kernel": "implicit-GEMM conv (icf_conv_forward + icf_conv_wgrad, all layers)",
                    "launches": conv["n"], "avg_launch_ms": conv["ms"] / conv["n"], "share_of_step": conv["ms"] / total,
                    "peak_source": f"bf16_tflops_sustained, {pk['src']}",
                    "hbm_frac_same_launches": conv["bytes"] / (conv["ms"] * 1e-3) / 1e9 / pk["hbm"]}
            # the single most expensive conv layer of the step against ITS roofline (max of tensor / HBM time); `traffic` =
            # DRAM bytes per launch of that layer from the committed `ncu --set full` capture when one exists
            convs = [(det, d) for det, d in layers.items() if det.startswith(("gather", "transp", "wgrad"))]
            if convs:
                det, d = max(convs, key=lambda kv: kv[1]["ms"])
                sec1 = d["ms"] * 1e-3 / d["n"]
                fl1, by1 = d["flops"] / d["n"], d["bytes"] / d["n"]
                hbm_bound = by1 / (pk["hbm"] * 1e9) > fl1 / (pk["tf_sust"] * 1e12)
                traffic = None
                try:
                    with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as tf:
                        traffic = json.load(tf).get(det)
                except Exception:
                    traffic = None
                roof["dominant_layer"] = {
                    "layer": det, "launches": d["n"], "avg_launch_ms": d["ms"] / d["n"], "bound": "hbm" if hbm_bound else "tensor",
                    "achieved": by1 / sec1 / 1e9 if hbm_bound else fl1 / sec1 / 1e12,
                    "peak": pk["hbm"] if hbm_bound else pk["tf_sust"], "unit": "GB/s" if hbm_bound else "TFLOP/s",
                    "frac": (by1 / sec1 / 1e9 / pk["hbm"]) if hbm_bound else (fl1 / sec1 / 1e12 / pk["tf_sust"]),
                    "algorithmic_bytes_per_launch": by1, "traffic": traffic}
        kern = {k: {"ms": round(v["ms"], 4), "n": v["n"]} for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- counterfactual pipeline (BASELINE.json configs[2]) --------------------------------------------------
    cf = None
    if not args.skip_cf:
        try:
            nb = args.cf_batch
            xc, ac, _ = synth.mnist_batch(nb, 7 + rank)
            imgs_cf, c0 = synth.mnist_scale(xc, ac, stats)
            _, c1 = synth.mnist_scale(xc, synth.intervene_mnist(ac), stats)
            imgs_cf = imgs_cf.to(dev)
            c0 = {k: v.to(dev) for k, v in c0.items()}
            c1 = {k: v.to(dev) for k, v in c1.items()}
            outb = torch.empty((nb, 1, 28, 28), device=dev)
            for _ in range(2):
                counterfactual(E, G, imgs_cf, c0, c1, out=outb)
            ms_cf = timed(lambda: counterfactual(E, G, imgs_cf, c0, c1, out=outb), 5)
            fl = forward_flops_per_image(FAMILIES["mnist"])
            cf = {"metric": "counterfactual_images_per_s", "value": world * nb / (ms_cf * 1e-3), "unit": "images/s",
                  "batch_per_gpu": nb, "ms": ms_cf,
                  "tensor_frac": (fl["E"] + fl["G"]) * nb / (ms_cf * 1e-3) / 1e12 / peaks()["tf_sust"]}
            del imgs_cf, outb
        except torch.cuda.OutOfMemoryError:
            cf = {"error": "out of memory"}

    def shutdown():
        """Tear NCCL down without risking a wedge: captured graphs that hold collectives are released first, and a
        watchdog ends the process if destroy_process_group() still does not return."""
        if world <= 1:
            return
        import gc
        import threading
        barrier()
        tr.graph = None
        gc.collect()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)

    if rank != 0:
        shutdown()
        return
    cpu = None
    if not args.skip_cpu and world == 1:
        rate, sec, cores = cpu_reference_rate(args.family, args.cpu_batch, 4, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"4 steps of {args.cpu_batch} images (oracle/bigan_ref.py train_step, torch CPU fp32)"}
    fl = forward_flops_per_image(FAMILIES["mnist"])
    step_flops = (4 * fl["E"] + 4 * fl["G"] + 12 * fl["D"]) * B
    line = {"metric": "bigan_train_step_images_per_s", "value": world * B / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"MorphoMNIST conditional BiGAN train step, batch {B} per GPU "
                                   "(image_scms/mnist.py:220-248)", "family": "mnist", "batch_per_gpu": B,
                       "global_batch": world * B, "parallelism": f"dp{world}", "cuda_graph": use_graph,
                       "l2": "working set per step (>1 GB of activations) exceeds the 126 MB L2",
                       "sync_bn": False},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roof,
            "step_tensor_frac": step_flops / (ms_step * 1e-3) / 1e12 / peaks()["tf_sust"],
            "cpu_baseline": cpu,
            "kernels_ms_per_step": kern,
            "counterfactual": cf}
    print(json.dumps(line), flush=True)
    sys.stdout.flush()
    shutdown()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
