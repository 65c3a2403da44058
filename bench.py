#!/usr/bin/env python
"""bench.py — BiGAN train-step images/s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores
    python bench.py --impl torch_gpu ...      # informational: stock torch (cuDNN/cuBLAS) running the reference's loop body

Workload (BASELINE.json configs[1], the default): MorphoMNIST conditional BiGAN train step (phases A-D of
image_scms/mnist.py:220-248), batch 4096 per GPU, bf16 activations / fp32 accumulation, synthetic MorphoMNIST-shaped
data, weights from init_weights (std 0.01).  ``--family audio_mnist|whalecalls|esrf_acoustic`` runs the same step of
the spectrogram families at the reference's default batch (128 / 32 / 64 per GPU, configs[3-4]).
One JSON line on stdout (rank 0).
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

DEFAULT_BATCH = {"mnist": 4096, "audio_mnist": 128, "whalecalls": 32, "esrf_acoustic": 64}
CF_BATCH = {"mnist": 65536, "audio_mnist": 512, "whalecalls": 128, "esrf_acoustic": 64}
NAMES = {"mnist": "MorphoMNIST", "audio_mnist": "AudioMNIST", "whalecalls": "whale-call", "esrf_acoustic": "ESRF"}
LOOP = {"mnist": "image_scms/mnist.py:220-248", "audio_mnist": "image_scms/audio_mnist.py:384-420",
        "whalecalls": "image_scms/whalecalls.py:453-499", "esrf_acoustic": "image_scms/esrf_acoustic.py:336-377"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the family's default)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--family", default="mnist", choices=list(DEFAULT_BATCH))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="images per CPU-baseline step (0: sized to the time budget)")
    ap.add_argument("--cf-batch", type=int, default=0)
    ap.add_argument("--skip-cf", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-torch", action="store_true")
    a = ap.parse_args()
    a.batch = a.batch or DEFAULT_BATCH[a.family]
    a.cf_batch = a.cf_batch or CF_BATCH[a.family]
    return a


def workload_config(args, world):
    """The `config` object of the JSON line — identical for our arm and the reference arm."""
    return {"workload": f"{NAMES[args.family]} conditional BiGAN train step, batch {args.batch} per GPU ({LOOP[args.family]})",
            "family": args.family, "batch_per_gpu": args.batch, "global_batch": world * args.batch,
            "parallelism": f"dp{world}"}


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


def family_batch(family, n, seed):
    """-> (images (n,1,H,W) in [-1,1], scaled attribute dict, z (n,512,1,1)); CPU fp32."""
    from icf_b200 import synth
    if family == "mnist":
        x, a, z = synth.mnist_batch(n, seed)
        images, c = synth.mnist_scale(x, a, synth.mnist_attr_stats())
        return images, c, z
    return synth.spectro_batch(family, n, seed)


def init_std(family):
    return 0.01 if family == "mnist" else 0.001


# -----------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port of the loop body) on the host cores
# -----------------------------------------------------------------------------------------------------------
class CpuReference:
    """The oracle's train step / counterfactual (the same torch CPU kernels the reference dispatches to), all host cores."""

    def __init__(self, family, seed=42):
        from oracle import bigan_ref as R
        self.R, self.family = R, family
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        sds = {k: R.synth_state_dict(family, k, seed, init_std(family)) for k in "EGD"}
        self.oracle = R.BiGANOracle(family, sds["E"], sds["G"], sds["D"])
        self.seed = seed

    def step_seconds(self, n):
        images, c, z = family_batch(self.family, n, self.seed)
        masks6 = [self.R.draw_masks(self.family, n) for _ in range(6)]
        t0 = time.perf_counter()
        self.oracle.train_step(images, c, z, masks6)
        return time.perf_counter() - t0

    def sample_size(self, batch, steps_total, budget_s):
        """Images per step so that `steps_total` steps fit the time budget (a bounded sample of the batch)."""
        probe = min(batch, 32 if self.family == "mnist" else 1)
        self.step_seconds(probe)                                   # first call pays one-time costs
        per_img = self.step_seconds(probe) / probe
        s = int(budget_s / max(steps_total, 1) / per_img)
        s = max(1, min(batch, s))
        if s >= 8:
            s = 1 << (s.bit_length() - 1)                          # a power of two
        return s

    def train_rate(self, n, steps, warmup):
        times = []
        for i in range(warmup + steps):
            sec = self.step_seconds(n)
            if i >= warmup:
                times.append(sec)
        sec = sum(times) / len(times)
        return n / sec, sec

    def cf_rate(self, n, reps=2):
        from icf_b200 import synth
        images, c, _ = family_batch(self.family, n, 7)
        if self.family == "mnist":
            x, a, _ = synth.mnist_batch(n, 7)
            _, c_cf = synth.mnist_scale(x, synth.intervene_mnist(a), synth.mnist_attr_stats())
        else:
            k = sorted(synth.ATTR_DIMS[self.family])[0]
            c_cf = dict(c)
            c_cf[k] = torch.roll(c[k], 1, dims=1)
        o = self.oracle
        self.R.counterfactual(self.family, o.E, o.G, images[:8], {k: v[:8] for k, v in c.items()}, {k: v[:8] for k, v in c_cf.items()})
        t0 = time.perf_counter()
        for _ in range(reps):
            self.R.counterfactual(self.family, o.E, o.G, images, c, c_cf)
        sec = (time.perf_counter() - t0) / reps
        return n / sec, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.family)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    n = args.cpu_batch or ref.sample_size(args.batch, steps + warm, 150.0)
    rate, sec = ref.train_rate(n, steps, warm)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {"impl": "reference", "metric": "bigan_train_step_images_per_s", "value": rate, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": ref.cores, "kind": "port",
                             "sample": f"every step = the loop body on a bounded sample of {n} images of the {args.batch}-image batch "
                                       f"(oracle/bigan_ref.py train_step: the reference's torch {torch.__version__} CPU kernels, fp32); "
                                       "images/s = sample / step time"},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------------
# informational arm: stock torch on the same GPU (the "existing Blackwell kernel" bar)
# -----------------------------------------------------------------------------------------------------------
def torch_gpu_rate(family, B, steps, warmup, mode, dev, seed=42):
    """images/s of stock torch (cuDNN / cuBLAS sm_100 kernels) executing the reference's loop body as the reference
    executes it (7 F_E + 7 F_G + 14 F_D, nn.BCEWithLogitsLoss, torch.optim.Adam) on the same GPU.  ``mode``: 'fp32'
    (TF32 off, the reference as shipped), 'tf32', or 'bf16_cl' (torch.autocast(bfloat16) + channels_last weights).  The
    layer programs are the oracle's functional restatement of the reference modules (the reference itself cannot travel to
    the GPU box); more generous than the reference in two places: z is drawn on the device (the reference draws it on the
    host, mnist.py:220-221) and the scores stay on the device (no per-step .item())."""
    import torch.nn.functional as F
    from oracle import bigan_ref as R
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    sds = {}
    for k in "EGD":
        sd = {n: (v.to(dev).float() if v.is_floating_point() else v.to(dev)) for n, v in R.synth_state_dict(family, k, seed, init_std(family)).items()}
        for n, v in sd.items():
            if v.is_floating_point() and not n.endswith(("running_mean", "running_var")):
                if mode == "bf16_cl" and v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
                sd[n] = v.requires_grad_(True)
        sds[k] = sd
    leaf = lambda sd: [v for v in sd.values() if v.requires_grad]
    betas = (0.5, 0.999) if family == "mnist" else (0.5, 0.9)
    optE = torch.optim.Adam(leaf(sds["E"]) + leaf(sds["G"]), lr=1e-4, betas=betas)
    optD = torch.optim.Adam(leaf(sds["D"]), lr=1e-4, betas=betas)
    images, c, _ = family_batch(family, B, seed)
    images = images.to(dev)
    c = {k: v.to(dev) for k, v in c.items()}
    valid, fake = torch.ones(B, 1, device=dev), torch.zeros(B, 1, device=dev)
    bce = F.binary_cross_entropy_with_logits
    score = torch.zeros(2, device=dev)
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "bf16_cl" else (lambda: torch.autocast("cuda", enabled=False))

    def E(xx):
        return R.encoder_fwd(family, sds["E"], xx, c)

    def G(zz):
        return R.generator_fwd(family, sds["G"], zz, c)

    def D(xx, zz):
        return R.discriminator_fwd(family, sds["D"], xx, zz, c, R.draw_masks(family, B, device=dev) or None)

    def step():
        z = torch.randn(B, 512, 1, 1, device=dev)
        with ctx():
            optE.zero_grad()
            loss = (bce(D(images, E(images)).float(), fake) + bce(D(G(z), z).float(), valid)) / 2
        loss.backward()
        optE.step()
        with ctx():
            optD.zero_grad()
            loss = bce(D(images, E(images)).float(), valid)
        loss.backward()
        optD.step()
        with ctx():
            optD.zero_grad()
            loss = bce(D(G(z), z).float(), fake)
        loss.backward()
        optD.step()
        with ctx():
            Gz, EX = G(z).detach(), E(images).detach()
            score[0] += torch.sigmoid(D(Gz, z).float()).mean()
            score[1] += torch.sigmoid(D(images, EX).float()).mean()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del optE, optD, sds
    torch.cuda.empty_cache()
    return B / (ms * 1e-3), ms


def torch_gpu_modes(family, B, steps, warmup, dev):
    out = {}
    for mode in ("fp32", "tf32", "bf16_cl"):
        try:
            rate, ms = torch_gpu_rate(family, B, steps, warmup, mode, dev)
            out[mode] = {"value": rate, "unit": "images/s", "ms_per_step": ms}
        except torch.cuda.OutOfMemoryError:
            out[mode] = {"error": "out of memory"}
            torch.cuda.empty_cache()
    out["what"] = ("stock torch (cuDNN/cuBLAS) executing the reference's loop body as the reference executes it "
                   "(7F_E+7F_G+14F_D, torch.optim.Adam) on this GPU, same batch; fp32 = as shipped (TF32 off), tf32, "
                   "bf16_cl = autocast(bfloat16) + channels_last")
    return out


def run_torch_gpu(args):
    """Informational arm: stock torch on the same GPU (not the reference arm, not the product)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    out = torch_gpu_modes(args.family, args.batch, args.steps, max(3, args.warmup), dev)
    print(json.dumps({"impl": "torch_gpu", "metric": "bigan_train_step_images_per_s", "family": args.family,
                      "batch_per_gpu": args.batch, "steps": args.steps, "modes": out}), flush=True)


# -----------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------
def run_ours(args):
    if os.environ.get("ICF_WATCHDOG"):          # development aid: dump every thread's stack and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["ICF_WATCHDOG"]), exit=True)
    import importlib
    import torch.distributed as dist
    from icf_b200 import ops, synth
    from icf_b200.arch import FAMILIES, forward_flops_per_image
    from icf_b200.trainer import BiGANTrainer, counterfactual, counterfactual_stream
    fam = args.family
    mod = importlib.import_module(f"image_scms.{fam}")
    H, W = FAMILIES[fam].image

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
    E, G, D = mod.Encoder().to(dev), mod.Generator().to(dev), mod.Discriminator().to(dev)
    for m in (E, G, D):
        m.apply(mod.init_weights)
    # BiGANTrainer broadcasts rank 0's parameters / buffers and offsets the ranks' random streams
    tr = BiGANTrainer(E, G, D, lr=1e-4, dtype=args.dtype, process_group=pg, sync_bn=args.sync_bn)
    # synthetic batch of the family's shape (each rank its own), pinned host copies for the end-to-end leg
    if fam == "mnist":
        x, a, _ = synth.mnist_batch(B, 42 + rank)                      # raw bytes + raw attributes, scaled on the device
        stats = synth.mnist_attr_stats()
        cont = sorted(k for k in a if k != "digit")
        lo = {k: stats[k][0].to(dev) for k in cont}
        span = {k: (stats[k][1] - stats[k][0]).to(dev) for k in cont}
    else:
        x, a, _ = synth.spectro_batch(fam, B, 42 + rank)
    h_x = x.pin_memory()
    h_a = {k: v.pin_memory() for k, v in a.items()}
    d_x = torch.empty(x.shape, device=dev)
    d_a = {k: torch.empty_like(v, device=dev) for k, v in a.items()}

    def upload_and_scale():
        """host->device copy of the raw batch + the scaling of mnist.py:204-209 (into the static step inputs)."""
        d_x.copy_(h_x, non_blocking=True)
        for k in d_a:
            d_a[k].copy_(h_a[k], non_blocking=True)
        if fam != "mnist":
            return d_x, d_a
        images = 2 * d_x.reshape(-1, 1, H, W) / 255 - 1
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
    checksum = [round(v, 6) for v in (tr.static["out"] if use_graph else tr.step(images, c))[:5].tolist()]

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
        out_dir = os.path.join(ROOT, "gpurun_out")
        with open(os.path.join(out_dir, f"per_layer_{fam}.json") if os.path.isdir(out_dir) else os.devnull, "w") as f:
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
            traffic_tab = {}
            try:
                with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as tf:
                    traffic_tab = json.load(tf)
            except Exception:
                traffic_tab = {}
            conv_rows = [r for r in per_layer if r["layer"].startswith(("gather", "transp", "wgrad"))]
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic_tab.get("conv_launch_avg_dram_bytes"),
                    "kernel": "implicit-GEMM conv (icf_conv_forward + icf_conv_wgrad, all layers)",
                    "launches": conv["n"], "avg_launch_ms": conv["ms"] / conv["n"], "share_of_step": conv["ms"] / total,
                    "algorithmic_flops_per_launch": conv["flops"] / conv["n"],
                    "algorithmic_bytes_per_launch": conv["bytes"] / conv["n"],
                    "peak_source": f"bf16_tflops_sustained, {pk['src']}",
                    "hbm_frac_same_launches": conv["bytes"] / (conv["ms"] * 1e-3) / 1e9 / pk["hbm"],
                    "layer_rows_at_or_above_half_roofline": sum(1 for r in conv_rows if r["roofline_frac"] >= 0.5),
                    "layer_rows": len(conv_rows)}
            # the single most expensive conv layer of the step against ITS roofline (max of tensor / HBM time); `traffic` =
            # DRAM bytes per launch of that layer from the committed `ncu --set full` capture when one exists
            convs = [(det, d) for det, d in layers.items() if det.startswith(("gather", "transp", "wgrad"))]
            if convs:
                det, d = max(convs, key=lambda kv: kv[1]["ms"])
                sec1 = d["ms"] * 1e-3 / d["n"]
                fl1, by1 = d["flops"] / d["n"], d["bytes"] / d["n"]
                hbm_bound = by1 / (pk["hbm"] * 1e9) > fl1 / (pk["tf_sust"] * 1e12)
                roof["dominant_layer"] = {
                    "layer": det, "launches": d["n"], "avg_launch_ms": d["ms"] / d["n"], "bound": "hbm" if hbm_bound else "tensor",
                    "achieved": by1 / sec1 / 1e9 if hbm_bound else fl1 / sec1 / 1e12,
                    "peak": pk["hbm"] if hbm_bound else pk["tf_sust"], "unit": "GB/s" if hbm_bound else "TFLOP/s",
                    "frac": (by1 / sec1 / 1e9 / pk["hbm"]) if hbm_bound else (fl1 / sec1 / 1e12 / pk["tf_sust"]),
                    "algorithmic_bytes_per_launch": by1, "traffic": traffic_tab.get(det)}
        kern = {k: {"ms": round(v["ms"], 4), "n": v["n"]} for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- counterfactual pipeline (BASELINE.json configs[2]) --------------------------------------------------
    cf = None
    fl = forward_flops_per_image(FAMILIES[fam])
    if not args.skip_cf:
        try:
            nb = args.cf_batch
            if fam == "mnist":
                xc, ac, _ = synth.mnist_batch(nb, 7 + rank)
                imgs_cf, c0 = synth.mnist_scale(xc, ac, stats)
                _, c1 = synth.mnist_scale(xc, synth.intervene_mnist(ac), stats)
            else:
                imgs_cf, c0, _ = synth.spectro_batch(fam, nb, 7 + rank)
                k0 = sorted(synth.ATTR_DIMS[fam])[0]
                c1 = dict(c0)
                c1[k0] = torch.roll(c0[k0], 1, dims=1)
            h_img = imgs_cf.pin_memory()
            h_c0 = {k: v.pin_memory() for k, v in c0.items()}
            h_c1 = {k: v.pin_memory() for k, v in c1.items()}
            d_img = h_img.to(dev)
            d_c0 = {k: v.to(dev) for k, v in h_c0.items()}
            d_c1 = {k: v.to(dev) for k, v in h_c1.items()}
            outb = torch.empty((nb, 1, H, W), device=dev)
            h_res = torch.empty((nb, 1, H, W)).pin_memory()
            for _ in range(2):
                counterfactual(E, G, d_img, d_c0, d_c1, out=outb)
            ms_cf = timed(lambda: counterfactual(E, G, d_img, d_c0, d_c1, out=outb), 5)

            def cf_e2e():
                # host batch -> counterfactual images on the host through the public streamed call: chunks of 8192 images,
                # H2D / kernels / D2H on three streams (every byte still crosses inside the timed region)
                counterfactual_stream(E, G, h_img, h_c0, h_c1, out=h_res, chunk=8192)
                torch.cuda.current_stream().synchronize()

            cf_e2e()
            ms_cf_e2e = timed(cf_e2e, 3)
            cf_h2d = h_img.numel() * 4 + sum(v.numel() * 4 for v in h_c0.values()) * 2
            cf_flops = (fl["E"] + fl["G"]) * nb
            tf = cf_flops / (ms_cf * 1e-3) / 1e12
            cf = {"metric": "counterfactual_images_per_s", "value": world * nb / (ms_cf * 1e-3), "unit": "images/s",
                  "batch_per_gpu": nb, "ms": ms_cf,
                  "e2e": {"value": world * nb / (ms_cf_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": cf_h2d,
                          "d2h_bytes_per_step": h_res.numel() * 4, "ms": ms_cf_e2e,
                          "api": "icf_b200.trainer.counterfactual_stream (8192-image chunks, copies overlapped)"},
                  "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks()["tf_sust"], "unit": "TFLOP/s",
                               "frac": tf / peaks()["tf_sust"], "traffic": None,
                               "algorithmic_flops": cf_flops, "what": "F_E + F_G valid-tap FLOPs of the whole pipeline / its time"},
                  "checksum": float(outb.double().sum())}
            del d_img, outb, h_res, h_img
        except torch.cuda.OutOfMemoryError:
            cf = {"error": "out of memory"}

    def shutdown():
        """Tear NCCL down without risking a wedge: captured graphs that hold collectives are released first, and a
        watchdog ends the process if destroy_process_group() still does not return."""
        if world <= 1:
            return
        import gc
        barrier()
        tr.graph = None
        gc.collect()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)

    if rank != 0:
        shutdown()
        return
    cudnn = None
    if not args.skip_torch and world == 1:
        tr.graph = None
        torch.cuda.empty_cache()
        cudnn = torch_gpu_modes(fam, B, 4, 3, dev)
    cpu = None
    if not args.skip_cpu and world == 1:
        ref = CpuReference(fam)
        n_cpu = args.cpu_batch or ref.sample_size(B, 4, 20.0)
        rate, sec = ref.train_rate(n_cpu, 3, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": ref.cores, "kind": "port",
               "sample": f"3 steps (after 1 warm-up) of the loop body on {n_cpu} images of the {B}-image batch "
                         "(oracle/bigan_ref.py train_step, torch CPU fp32)"}
        if cf and "value" in cf:
            n_cf = min(args.cf_batch, 4096 if fam == "mnist" else max(1, DEFAULT_BATCH[fam] // 8))
            cf_rate, cf_sec = ref.cf_rate(n_cf)
            cf["cpu_baseline"] = {"value": cf_rate, "unit": "images/s", "cores": ref.cores, "kind": "port",
                                  "sample": f"G(E(x,c),c_cf) under no_grad on {n_cf} images (oracle/bigan_ref.py counterfactual)"}
    step_flops = (4 * fl["E"] + 4 * fl["G"] + 12 * fl["D"]) * B
    cfg = workload_config(args, world)
    cfg.update({"cuda_graph": use_graph, "sync_bn": bool(args.sync_bn),
                "l2": "working set per step exceeds the 126 MB L2 (activations of one D forward alone are >0.6 GB at this batch)"})
    line = {"metric": "bigan_train_step_images_per_s", "value": world * B / (ms_step * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roof,
            "step_tensor_frac": step_flops / (ms_step * 1e-3) / 1e12 / peaks()["tf_sust"],
            "cpu_baseline": cpu,
            "cudnn_baseline": cudnn,
            "checksum": {"losses_scores_accumulated": checksum},
            "kernels_ms_per_step": kern,
            "counterfactual": cf}
    print(json.dumps(line), flush=True)
    sys.stdout.flush()
    shutdown()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "torch_gpu":
        run_torch_gpu(a)
    else:
        run_ours(a)
