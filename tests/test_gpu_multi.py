"""Multi-GPU parity: N ranks with SyncBN on a batch split over the ranks reproduce the 1-rank step on the
whole batch.  Needs >= 2 GPUs (skipped otherwise); launched as two processes over NCCL on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "imagecfgen-pytorch_b200")); sys.path.insert(0, os.path.join(root, "tests"))
from helpers import golden_inputs
from oracle import bigan_ref as R
from icf_b200.trainer import BiGANTrainer
from image_scms import mnist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
fam, n, seed, std, dtype = "mnist", 64, 16, 0.05, sys.argv[2]
images, c, z, _ = golden_inputs(fam, n, seed)
torch.manual_seed(9)
masks6 = [R.draw_masks(fam, n) for _ in range(6)]
def nets(device, perturb):
    out = {}
    for k, cls in (("E", mnist.Encoder), ("G", mnist.Generator), ("D", mnist.Discriminator)):
        m = cls(); m.load_state_dict(R.synth_state_dict(fam, k, seed, std))
        if perturb:                      # rank 1 starts from different weights: the trainer's broadcast must repair that
            with torch.no_grad():
                for p in m.parameters(): p.add_(0.01)
        out[k] = m.to(device)
    return out
# data-parallel step: rank r owns samples [r*n/world, (r+1)*n/world)
lo, hi = rank * n // world, (rank + 1) * n // world
N = nets(dev, perturb=(rank == 1))
tr = BiGANTrainer(N["E"], N["G"], N["D"], dtype=dtype, process_group=dist.group.WORLD, sync_bn=True)
sl = lambda t: t[lo:hi].to(dev)
out = tr.step(sl(images), {k: sl(v) for k, v in c.items()}, sl(z), [[sl(m) for m in ms] for ms in masks6])
out = tr.reduce_scores(out)
res = {"out": out[:5].tolist()}
if rank == 0:
    # single-GPU reference on the whole batch, same process
    M = nets(dev, perturb=False)
    tr1 = BiGANTrainer(M["E"], M["G"], M["D"], dtype=dtype)
    full = lambda t: t.to(dev)
    o1 = tr1.step(full(images), {k: full(v) for k, v in c.items()}, full(z), [[full(m) for m in ms] for ms in masks6])
    res["single"] = o1[:5].tolist()
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    res["w_err"] = max(rel(tr.gEG.flat, tr1.gEG.flat), rel(tr.gD.flat, tr1.gD.flat))
    res["bn_err"] = max(rel(a.float(), b.float()) for a, b in zip(tr.D.buffers(), tr1.D.buffers()))
# replicas stay identical
w = tr.gD.flat.clone(); dist.broadcast(w, 0)
res["replica_diff"] = float((w - tr.gD.flat).abs().max())
print("RESULT " + json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_two_ranks_with_sync_bn_equal_one_rank_on_the_whole_batch(tmp_path, dtype):
    import json
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, dtype], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = [json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][-1][7:]) for o in outs]
    print(res)
    tol = 1e-3 if dtype == "fp32" else 2e-2
    r0 = res[0]
    for a, b in zip(r0["out"], r0["single"]):
        assert abs(a - b) <= tol * max(abs(b), 0.1), r0
    assert r0["w_err"] < (1e-4 if dtype == "fp32" else 5e-3) and r0["bn_err"] < tol, r0
    assert all(r["replica_diff"] == 0.0 for r in res), res
