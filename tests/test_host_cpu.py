"""CPU-side checks: the C-ABI library loads and exports every symbol include/icf.h declares, the drop-in modules
keep the reference's state_dict layout, host logic (batching, mask order, FLOP accounting, flat parameter
groups, 2-rank gloo gradient averaging) — no kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from icf_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "icf.h")).read()
    declared = set(re.findall(r"\b(icf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"icf_conv_args", "icf_wgrad_args"}
    assert os.path.exists(lib.LIB_PATH), "build with python imagecfgen-pytorch_b200/build.py"
    h = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(h, s)]
    assert not missing, missing
    assert set(lib.EXPORTED_SYMBOLS) == declared
    assert lib.load().icf_version() == 1
    # no entry point needs a hidden workspace (SURVEY §8b query): 0 for every exported kernel entry, -1 for an unknown name
    assert lib.load().icf_workspace_bytes(b"icf_conv_forward", None) == 0 and lib.load().icf_workspace_bytes(b"nope", None) == -1


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the argument structs must have the C sizes (checked against a tiny gcc program)."""
    from icf_b200 import lib
    src = r'''
#include <stdio.h>
#include "icf.h"
int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(icf_conv_args), sizeof(icf_wgrad_args), sizeof(icf_perm),
 sizeof(icf_imgfeat_args), sizeof(icf_latfeat_args), sizeof(icf_actbwd_args), sizeof(icf_explain_group), sizeof(icf_scm_affine_args),
 sizeof(icf_pack_job));return 0;}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    mine = [ctypes.sizeof(t) for t in (lib.ConvArgs, lib.WgradArgs, lib.Perm, lib.ImgFeatArgs, lib.LatFeatArgs,
                                       lib.ActBwdArgs, lib.ExplainGroup, lib.ScmAffineArgs, lib.PackJob)]
    assert mine == sizes, (mine, sizes)


@pytest.mark.parametrize("fam", ["mnist", "audio_mnist", "whalecalls", "esrf_acoustic"])
def test_state_dict_layout_matches_reference_tables(fam):
    import importlib
    from oracle.arch import param_shapes
    m = importlib.import_module(f"image_scms.{fam}")
    with torch.device("meta"):
        nets = {"E": m.Encoder(), "G": m.Generator(), "D": m.Discriminator()}
    for k, net in nets.items():
        sd = net.state_dict()
        want = param_shapes(fam, k)
        got = [(n, tuple(v.shape)) for n, v in sd.items() if not n.endswith("num_batches_tracked")]
        assert got == [(n, tuple(s)) for n, s in want], (fam, k)
        assert type(net).__module__ == f"image_scms.{fam}"


def test_default_init_and_init_weights():
    from image_scms import mnist
    from image_scms.training_utils import init_weights
    torch.manual_seed(0)
    D = mnist.Discriminator()
    w = D.dxz._modules["1"].weight
    bound = 1 / (1024 ** 0.5)
    assert float(w.abs().max()) <= bound and float(w.std()) > 0.5 * bound / 3 ** 0.5
    emb0 = D.digit_embedding._modules["0"].weight.clone()
    bn_w = D.dx._modules["4"].weight.clone()
    D.apply(init_weights)
    assert abs(float(D.dxz._modules["1"].weight.std()) - 0.01) < 1e-3
    assert float(D.dxz._modules["1"].bias.abs().max()) == 0.0
    assert torch.equal(emb0, D.digit_embedding._modules["0"].weight)      # Embedding / BN keep torch defaults
    assert torch.equal(bn_w, D.dx._modules["4"].weight)


def test_cpu_tensors_are_refused():
    from image_scms import mnist
    E = mnist.Encoder()
    c = {"digit": torch.eye(10)[:2], "thickness": torch.zeros(2, 1), "intensity": torch.zeros(2, 1),
         "slant": torch.zeros(2, 1)}
    with pytest.raises(RuntimeError, match="no CPU"):
        E(torch.zeros(2, 1, 28, 28), c)


def test_batchify_and_mask_sites():
    from image_scms.training_utils import batchify, batchify_dict
    x, y = torch.arange(10), torch.arange(10) * 2
    got = list(batchify(x, y, batch_size=4))
    assert [len(b[0]) for b in got] == [4, 4, 2] and torch.equal(got[2][1], torch.tensor([16, 18]))
    gd = list(batchify_dict({"a": x, "b": y}, batch_size=3))
    assert [len(b["a"]) for b in gd] == [3, 3, 3, 1]
    from icf_b200.arch import FAMILIES
    from icf_b200.engine import mask_sites
    sites = [(p, c) for (_, _, _, p, c) in mask_sites(FAMILIES["mnist"])]
    assert sites == [(0.2, 5), (0.2, 32), (0.5, 64), (0.5, 128), (0.5, 256), (0.2, 512), (0.5, 512), (0.2, 1024),
                     (0.2, 1024), (0.2, 1024)]
    from oracle.bigan_ref import dropout_sites
    assert sites == dropout_sites("mnist")
    assert mask_sites(FAMILIES["audio_mnist"]) == []


def test_flop_accounting_matches_survey():
    """per-image forward FLOPs (valid taps)."""
    from icf_b200.arch import FAMILIES, forward_flops_per_image
    f = forward_flops_per_image(FAMILIES["mnist"])
    assert abs(f["E"] / 1e6 - 22.96) < 0.01 and abs(f["G"] / 1e6 - 75.71) < 0.01 and abs(f["D"] / 1e6 - 46.36) < 0.01
    f = forward_flops_per_image(FAMILIES["audio_mnist"])
    assert abs(f["E"] / 1e6 - 1293.2) < 0.5 and abs(f["G"] / 1e6 - 1534.3) < 0.5 and abs(f["D"] / 1e6 - 1298.5) < 0.5


def _ws_plan(a):
    """(header, classes, schedule words) of the row-streaming kernel for ConvArgs `a`, or None if it declines."""
    from icf_b200 import lib
    words = 16 + 4 * 48 + 4096
    out = (ctypes.c_int32 * words)()
    rc = lib.load().icf_ws_plan(ctypes.byref(a), out, words)
    if rc == -1:
        return None
    assert rc == 0, lib.load().icf_last_error()
    o = list(out)
    classes = []
    for c in range(o[0]):
        b = 16 + 48 * c
        cl = dict(zip(("Pi", "Qj", "py", "px", "ylo", "yhi", "dymax", "ngroups", "ntaps", "tiles_x", "cta_begin", "cta_count"),
                      o[b:b + 12]))
        cl["prog_off"] = o[b + 12:b + 12 + o[8] + 1]
        cl["groups"] = [tuple(o[b + 16 + 4 * g:b + 20 + 4 * g]) for g in range(cl["ngroups"])]
        classes.append(cl)
    return o[:16], classes, o[16 + 4 * 48:16 + 4 * 48 + o[9]]


def _check_ws_schedule(hdr, classes, prog):
    """Replay every issuer's schedule of every class and check the accumulator protocol the kernel relies on."""
    n_acc, sstep, issuers = hdr[4], hdr[6], hdr[8]
    assert hdr[1] * hdr[2] == 128 and hdr[3] >= 2 and sum(c["cta_count"] for c in classes) == hdr[10] <= 148
    for cl in classes:
        Pi, ylo, yhi = cl["Pi"], cl["ylo"], cl["yhi"]
        dys = [g[0] for g in cl["groups"]]
        # ground truth: which (source row, group) pairs feed which output row
        feeds = {}
        for y in range(ylo, yhi + 1):
            for gi, dy in enumerate(dys):
                num = y - dy
                if num < 0 or num % sstep:
                    continue
                i = num // sstep
                if i < Pi:
                    feeds.setdefault(i, []).append((y, gi))
        assert sorted(feeds) == list(range(Pi)), "every output row needs a source row"
        seen, opened, completed = {}, {}, {}
        for wi in range(issuers):
            words = prog[cl["prog_off"][wi]:cl["prog_off"][wi + 1]]
            assert (words[-1] >> 18) & 3 == 2, "spare word at the end of an issuer's list"
            y, k = ylo, 0
            while y <= yhi:
                w = words[k]
                k += 1
                i, first, cnt, kind = w & 0xFF, (w >> 8) & 31, (w >> 13) & 31, (w >> 18) & 3
                if kind == 0:
                    assert i % issuers == wi, "an accumulator is fed by one issuer only"
                    gi = next(g for g, grp in enumerate(cl["groups"]) if grp[2] == first and grp[3] == cnt)
                    assert (y, gi) in feeds[i] and (i, y, gi) not in seen
                    seen[(i, y, gi)] = True
                    assert i not in completed, "no chain after the accumulator was handed to the epilogue"
                    if (w >> 20) & 1:
                        assert i not in opened
                        opened[i] = y
                    else:
                        assert i in opened, "first chain of an output row must clear the accumulator"
                elif kind == 1:
                    assert i % issuers == wi and i in opened and i not in completed
                    completed[i] = y
                if (w >> 21) & 1:
                    y += 1
            assert k == len(words) - 1
        assert len(seen) == sum(len(v) for v in feeds.values()), "every contribution issued exactly once"
        for i in range(Pi):
            assert opened[i] == min(y for y, _ in feeds[i]) and completed[i] >= max(y for y, _ in feeds[i])
        # output rows in flight at any source row never exceed the accumulator ring minus the one being drained
        for y in range(ylo, yhi + 1):
            live = sum(1 for i in range(Pi) if opened[i] <= y <= completed[i])
            assert live <= n_acc - 1, (live, n_acc)


@pytest.mark.parametrize("fam", ["mnist", "audio_mnist", "whalecalls", "esrf_acoustic"])
def test_row_streaming_schedules_replay(fam):
    """Host logic of the weight-stationary kernel (icf_ws_plan, no launch): for every conv layer of the family, forward
    and data-gradient form, the issuer schedules must issue every (source row, filter row) contribution exactly once,
    from the issuer that owns the output row, clear each accumulator on its first chain and hand it over after its last."""
    from icf_b200 import lib
    from icf_b200.arch import FAMILIES

    def pad8(c):
        return (c + 7) // 8 * 8

    f = FAMILIES[fam]
    N, served = 256, 0
    for tower, h0 in (("E", f.image), ("G", (1, 1)), ("Dx", f.image)):
        h, w = h0
        for li, sp in enumerate(getattr(f, tower)):
            if sp.kind not in ("conv", "convT"):
                continue
            if sp.kind == "conv":
                P, Q = (h + 2 * sp.pad - sp.k) // sp.stride + 1, (w + 2 * sp.pad - sp.k) // sp.stride + 1
                form_f, form_b = lib.FORM_GATHER, lib.FORM_TRANSPOSED
            else:
                P, Q = (h - 1) * sp.stride - 2 * sp.pad + sp.k, (w - 1) * sp.stride - 2 * sp.pad + sp.k
                form_f, form_b = lib.FORM_TRANSPOSED, lib.FORM_GATHER
            cases = [(form_f, h, w, sp.cin, P, Q, sp.cout, sp.k, sp.k, sp.stride, sp.pad, 0),
                     (form_b, P, Q, sp.cout, h, w, sp.cin, sp.k, sp.k, sp.stride, sp.pad, 0)]
            if li == 0 and tower in ("E", "Dx"):     # folded first layer: filter columns live in the channel dimension
                cases.append((form_f, h + 2 * sp.pad, w + 2 * sp.pad, sp.k * 8, P, Q, sp.cout, sp.k, 1, sp.stride, 0, sp.k))
            for form, H, W, Cc, Po, Qo, K, R, S, stride, pad, win in cases:
                a = lib.ConvArgs(lib.BF16, form, N, H, W, Cc, 8 if win else pad8(Cc), Po, Qo, K, pad8(K), R, S, stride, pad,
                                 K, pad8(Cc), 0, 0.0, 0, 0, 0, win, 0x10000, 0x20000, None, 0x30000, None, None)
                plan = _ws_plan(a)
                if plan is not None:
                    _check_ws_schedule(*plan)
                    served += 1
            h, w = P, Q
    assert served > 0 or fam == "esrf_acoustic"


@pytest.mark.parametrize("fam", ["mnist", "audio_mnist", "whalecalls", "esrf_acoustic"])
@pytest.mark.parametrize("batch", [64, 4096])
def test_wgrad_plans_fit_the_machine(fam, batch):
    """Host logic of the tensor-core weight-gradient kernel (icf_wgrad_plan, no launch): tap groups cover the taps, the
    accumulators fit TMEM, the stage ring fits shared memory at the assumed CTAs per SM, and the grid is one resident wave
    (a ragged second wave cost 20-30 %)."""
    from icf_b200 import lib
    from icf_b200.arch import FAMILIES

    def pad8(c):
        return (c + 7) // 8 * 8

    f = FAMILIES[fam]
    served = 0
    for tower, h0 in (("E", f.image), ("G", (1, 1)), ("Dx", f.image)):
        h, w = h0
        for sp in getattr(f, tower):
            if sp.kind not in ("conv", "convT"):
                continue
            if sp.kind == "conv":
                P, Q = (h + 2 * sp.pad - sp.k) // sp.stride + 1, (w + 2 * sp.pad - sp.k) // sp.stride + 1
                small, big = (P, Q, sp.cout), (h, w, sp.cin)          # dY is the small operand, X the strided one
            else:
                P, Q = (h - 1) * sp.stride - 2 * sp.pad + sp.k, (w - 1) * sp.stride - 2 * sp.pad + sp.k
                small, big = (h, w, sp.cin), (P, Q, sp.cout)          # transposed conv: roles swap
            a = lib.WgradArgs(lib.BF16, batch, small[0], small[1], small[2], pad8(small[2]), big[0], big[1], big[2],
                              pad8(big[2]), sp.k, sp.k, sp.stride, sp.pad, 0x10000, 0x20000, 0x30000, 0)
            out = (ctypes.c_int32 * 16)()
            rc = lib.load().icf_wgrad_plan(ctypes.byref(a), out, 16)
            h, w = P, Q
            if rc == -1:
                continue
            assert rc == 0, lib.load().icf_last_error()
            (tile_n, tp, groups, stages, stage_bytes, cols, smem, ctas, tiles, splits, bq, bp, bn, rows, n_blocks,
             taps) = list(out)
            served += 1
            assert taps == sp.k * sp.k and (groups - 1) * tp < taps <= groups * tp
            assert cols >= tp * tile_n and cols & (cols - 1) == 0 and ctas * cols <= 512
            assert stages >= 2 and smem == stages * stage_bytes + 1280 and ctas * smem <= 227 * 1024
            assert stage_bytes == (2 + tp * tile_n // 64) * 8192
            assert rows == bq * bp * bn and rows in (16, 32, 48, 64)
            assert n_blocks == -(-small[1] // bq) * -(-small[0] // bp) * -(-batch // bn)
            assert 1 <= splits <= n_blocks
            assert tiles * splits <= 148 * ctas or splits == 1, "one resident wave"
            if tiles <= 148 * ctas and n_blocks >= 148 * ctas:      # floor(capacity / tiles) keeps more than half the slots busy
                assert tiles * splits > 148 * ctas // 2, "grid should fill at least half the machine"
    assert served > 0


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from icf_b200.trainer import _FlatGroup
from image_scms import mnist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
E, G = mnist.Encoder(), mnist.Generator()
named = [("E." + n, p) for n, p in E.named_parameters()] + [("G." + n, p) for n, p in G.named_parameters()]
grp = _FlatGroup(named, torch.device("cpu"))
n_E = len(list(E.named_parameters()))
assert all(p.data_ptr() >= grp.flat.data_ptr() for p in grp.params)          # parameters became views
for name, v in grp.grad_views.items():
    v.fill_(float(rank + 1))
# two buckets (E first, then G), as BiGANTrainer.step launches them
for lo_i, hi_i in ((0, n_E), (n_E, len(grp.params))):
    lo, hi = grp.segment(lo_i, hi_i)
    dist.all_reduce(grp.grad[lo:hi])
avg = grp.grad / world
want = sum(range(1, world + 1)) / world
ok = all(bool((avg[o:o + p.numel()] == want).all()) for p, o in zip(grp.params, grp.offsets))
dist.barrier()
print("OK" if ok else "BAD", flush=True)
dist.destroy_process_group()
'''


def test_two_rank_gloo_gradient_buckets(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), os.path.join(ROOT, "imagecfgen-pytorch_b200")],
                              env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("OK" in o for o in outs), outs


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from icf_b200 import dp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = dp.Group(dist.group.WORLD)
assert (g.rank, g.world) == (rank, world)
# replicas: every rank ends with rank 0's parameters and buffers
torch.manual_seed(100 + rank)
flat, buf, nbt = torch.randn(1000), torch.randn(7), torch.tensor(rank, dtype=torch.long)
g.broadcast_state([flat, buf, nbt])
torch.manual_seed(100)
assert torch.equal(flat, torch.randn(1000)) and torch.equal(buf, torch.randn(7)) and int(nbt) == 0
# rank-offset random streams: rank 0 keeps the default stream, the others leave it
torch.manual_seed(5)
g.seed_offset("cpu")
mine = torch.rand(4)
torch.manual_seed(5)
base = torch.rand(4)
assert torch.equal(mine, base) == (rank == 0)
# sharded epoch: same step count on every rank, disjoint batches, drawn by rank 0
import numpy as np
np.random.seed(1000 + rank)
batches = dp.shard_permutation(103, 8, g)
assert len(batches) == 103 // 16 and all(len(b) == 8 for b in batches)
mine = torch.cat(batches)
both = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(both, mine)
allidx = torch.cat(both)
assert allidx.unique().numel() == allidx.numel() == 96
# SyncBN statistics: sum / sum of squares of the shards = those of the whole batch
torch.manual_seed(7)
x = torch.randn(2 * 6, 5, 3, 3)
shard = x[rank * 6:(rank + 1) * 6]
stats = torch.cat([shard.sum(dim=(0, 2, 3)), shard.square().sum(dim=(0, 2, 3))])
dp.sync_stats(g, stats)
cnt = 6 * 9 * world
mean, var = stats[:5] / cnt, stats[5:] / cnt - (stats[:5] / cnt) ** 2
assert torch.allclose(mean, x.mean(dim=(0, 2, 3)), atol=1e-5) and torch.allclose(var, x.var(dim=(0, 2, 3), unbiased=False), atol=1e-5)
dist.barrier()
print("OK", flush=True)
dist.destroy_process_group()
'''


def test_two_rank_gloo_data_parallel_plumbing(tmp_path):
    """icf_b200.dp with world_size 2 over gloo: replica broadcast, rank-offset RNG, sharded permutation, SyncBN statistics."""
    script = tmp_path / "dpw.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), os.path.join(ROOT, "imagecfgen-pytorch_b200")],
                              env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("OK" in o for o in outs), outs


def test_gradient_buckets_and_single_rank_sharding():
    from icf_b200 import dp
    offs = [0, 10, 30, 31, 100, 260]
    r = dp.bucket_ranges(offs, 64)
    assert r[0][1] == 260 and r[-1][0] == 0                       # from the end of the buffer to its start
    assert all(a[0] == b[1] for a, b in zip(r, r[1:]))            # contiguous, no gaps
    assert all(hi - lo >= 64 for lo, hi in r[:-1]) and all(lo in offs and hi in offs for lo, hi in r)
    g = dp.Group(None)
    import numpy as np
    np.random.seed(3)
    b = dp.shard_permutation(10, 4, g)
    assert [len(t) for t in b] == [4, 4, 2] and sorted(torch.cat(b).tolist()) == list(range(10))    # batchify semantics


def test_train_signatures_match_the_reference():
    """Positional parameters of the four train() entry points (mnist.py:157-167, audio_mnist.py:321-327, whalecalls.py:390-399,
    esrf_acoustic.py:263-272): a reference caller (train_esrf_bigan.py:24 passes path_to_wavs, path_to_labels positionally)
    must bind the same way; our additions are keyword-only."""
    import importlib
    import inspect
    want = {"mnist": ["x_train", "a_train", "x_test", "a_test", "n_epochs", "l_rate", "device", "save_images_every",
                      "image_output_path", "batch_size", "d_updates_per_g_update"],
            "audio_mnist": ["path_to_zip", "n_epochs", "l_rate", "device", "save_images_every", "batch_size", "image_output_path"],
            "whalecalls": ["nocall_directory", "gunshot_directory", "upcall_directory", "n_epochs", "l_rate", "device",
                           "save_images_every", "batch_size", "image_output_path", "filter_length"],
            "esrf_acoustic": ["path_to_wavs", "path_to_labels", "n_epochs", "l_rate", "device", "save_images_every", "batch_size",
                              "image_output_path", "validation_split", "start_model_path"]}
    for fam, names in want.items():
        sig = inspect.signature(importlib.import_module(f"image_scms.{fam}").train)
        pos = [p.name for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
        assert pos == names, (fam, pos)
        assert all(p.kind == p.KEYWORD_ONLY for p in sig.parameters.values() if p.name not in names)


def test_explainer_signatures_and_layout_match_the_reference():
    """explain/cf_example.py:18-34,83-103: constructor / explain() positional parameters of the two explainers under the reference's
    import path (our additions are keyword-only), and the flat raw-row layout of the hinge explainer: upstream's dict order, the
    ignored attributes absent, the latent row last, 16-byte aligned groups."""
    import inspect
    from explain.cf_example import DeepCounterfactualExplainer, HingeLossCFExplainer, hinge, max_excluding, mse  # noqa: F401
    from icf_b200 import explain as X

    def pos(fn):
        ps = list(inspect.signature(fn).parameters.values())
        assert all(p.kind == p.KEYWORD_ONLY for p in ps if p.kind != p.POSITIONAL_OR_KEYWORD)
        return [p.name for p in ps if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert pos(DeepCounterfactualExplainer.__init__) == ["self", "encoder", "decoder", "classifier", "target_feature"]
    assert pos(DeepCounterfactualExplainer.explain) == ["self", "x", "attrs", "target_class", "sample_points", "metric"]
    assert pos(HingeLossCFExplainer.__init__) == ["self", "encoder", "decoder", "classifier", "target_feature", "latent_dim",
                                                  "categorical_features", "features_to_ignore", "c"]
    assert pos(HingeLossCFExplainer.explain) == ["self", "x", "attrs", "target_class", "train_z", "steps", "lr"]
    d = inspect.signature(HingeLossCFExplainer.explain).parameters
    assert (d["target_class"].default, d["train_z"].default, d["steps"].default, d["lr"].default) == (None, True, 30, 0.1)
    assert inspect.signature(HingeLossCFExplainer.__init__).parameters["c"].default == 10.0
    ex = HingeLossCFExplainer(None, None, None, "digit", 512, categorical_features=["digit"], features_to_ignore=["slant"])
    attrs = {"thickness": torch.zeros(3, 1), "intensity": torch.zeros(3, 1), "slant": torch.zeros(3, 1), "digit": torch.zeros(3, 10)}
    specs, n = ex._layout(attrs, 3, True)
    assert [(k, m, w) for k, m, w, _ in specs] == [("thickness", X.TANH, 1), ("intensity", X.TANH, 1), ("digit", X.SOFTMAX, 10),
                                                  ("z", X.TANH, 512)]
    offs = [o for *_, o in specs]
    assert offs == [0, 4, 8, 40] and all(o % 4 == 0 for o in offs) and n == 40 + 3 * 512
    assert float(max_excluding(torch.tensor([[1.0, 5.0, 3.0], [9.0, 2.0, 4.0]]), torch.tensor([1, 0]))[0]) == 3.0
    assert torch.equal(max_excluding(torch.tensor([[1.0, 5.0, 3.0], [9.0, 2.0, 4.0]]), 0), torch.tensor([5.0, 4.0]))


class _TinyNet(torch.nn.Module):
    def __init__(self, v):
        super().__init__()
        self.w = torch.nn.Parameter(torch.full((3,), float(v)))


def test_esrf_warm_start_is_what_gets_trained(monkeypatch, tmp_path):
    """esrf_acoustic.py:276-284: init_weights first, THEN the networks of start_model_path replace E/G/D — the loop must
    receive the checkpoint's modules untouched (round 1 re-initialised them inside the loop)."""
    from image_scms import esrf_acoustic as m
    Tiny = _TinyNet
    path = str(tmp_path / "ck.tar")
    torch.save({"E": Tiny(1), "G": Tiny(2), "D": Tiny(3)}, path)
    seen = {}

    def fake_loop(E, G, D, *a, **k):
        seen["nets"] = (E, G, D)
        return E, G, D, None, None

    monkeypatch.setattr(m, "_fresh", lambda device: (Tiny(0), Tiny(0), Tiny(0)))
    monkeypatch.setattr(m, "train_stream", fake_loop)
    E, G, D, _, _ = m.train("wavs", "labels", 1, 1e-4, "cpu", 2, 2, "", 0.2, path, data=object())
    assert [float(n.w[0]) for n in seen["nets"]] == [1.0, 2.0, 3.0]
    m.train("wavs", "labels", 1, data=object())
    assert [float(n.w[0]) for n in seen["nets"]] == [0.0, 0.0, 0.0]
    with pytest.raises(NotImplementedError, match="dataset reader"):
        m.train("wavs", "labels")
