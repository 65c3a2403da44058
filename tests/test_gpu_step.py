"""The BENCHMARKED configuration held to the oracle: CUDA-graph capture / replay, the B=4096 bf16 train step, the B=65536
counterfactual batch, the train() entry points of the four families and the reference-written whole-module pickle."""
import io
import os

import pytest
import torch
import torch.nn as nn

from helpers import golden_inputs, rel_err
from oracle import bigan_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def family_module(fam):
    import importlib
    return importlib.import_module(f"image_scms.{fam}")


def build(fam, seed, std, dtype):
    m = family_module(fam)
    nets = {}
    for k, cls in (("E", m.Encoder), ("G", m.Generator), ("D", m.Discriminator)):
        net = cls()
        net.load_state_dict(R.synth_state_dict(fam, k, seed, std))
        nets[k] = net.to(DEV).set_compute_dtype(dtype)
    return nets


def to_dev(c):
    return {k: v.to(DEV) for k, v in c.items()}


def _snapshot(tr):
    return [t.clone() for t in (tr.gEG.flat, tr.gEG.exp_avg, tr.gEG.exp_avg_sq, tr.gD.flat, tr.gD.exp_avg, tr.gD.exp_avg_sq,
                                tr.stateEG, tr.stateD, tr.stateG)] + [b.clone() for b in tr.D.buffers()]


def _restore(tr, snap):
    dst = [tr.gEG.flat, tr.gEG.exp_avg, tr.gEG.exp_avg_sq, tr.gD.flat, tr.gD.exp_avg, tr.gD.exp_avg_sq, tr.stateEG,
           tr.stateD, tr.stateG] + list(tr.D.buffers())
    for d, s in zip(dst, snap):
        d.copy_(s)
    for ex in (tr.exE, tr.exG, tr.exD):
        ex.repack(force=True)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_graph_replay_matches_eager_and_oracle(dtype):
    """capture()/replay() is what bench.py times.  From the same parameters, Adam state and generator seed, (1) the eager
    step with z and the Dropout2d masks drawn inside the step equals the oracle fed the same draws (the masks follow torch's
    stream, so they can be re-drawn here), and (2) a replay of the captured graph equals the eager step — twice, so that the
    second replay proves the graph advances the generator and the optimiser state like two eager steps do."""
    from icf_b200.arch import FAMILIES
    from icf_b200.engine import draw_masks
    from icf_b200.trainer import BiGANTrainer
    fam, n, seed, std = "mnist", 64, 16, 0.05
    tol = 1e-3 if dtype == "fp32" else 2e-2
    images, c, _, _ = golden_inputs(fam, n, seed)
    nets = build(fam, seed, std, dtype)
    tr = BiGANTrainer(nets["E"], nets["G"], nets["D"], lr=1e-4)
    x, cd = images.to(DEV), to_dev(c)
    snap = _snapshot(tr)

    def draws(s):
        torch.manual_seed(s)
        z = torch.randn(n, 512, 1, 1, device=DEV)
        return z, [draw_masks(FAMILIES[fam], n, torch.device(DEV)) for _ in range(6)]

    z1, m1 = draws(1234)
    z2, m2 = draws(1234)
    assert torch.equal(z1, z2) and all(torch.equal(a, b) for A, B in zip(m1, m2) for a, b in zip(A, B))
    # two steps: the second starts where the generator stands after the first
    torch.manual_seed(1234)
    za = torch.randn(n, 512, 1, 1, device=DEV)
    ma = [draw_masks(FAMILIES[fam], n, torch.device(DEV)) for _ in range(6)]
    zb = torch.randn(n, 512, 1, 1, device=DEV)
    mb = [draw_masks(FAMILIES[fam], n, torch.device(DEV)) for _ in range(6)]
    o = R.BiGANOracle(fam, *(R.synth_state_dict(fam, k, seed, std) for k in "EGD"))
    want = []
    for zz, mm in ((za, ma), (zb, mb)):
        r = o.train_step(images, c, zz.cpu(), [[t.cpu() for t in ms] for ms in mm])
        want.append([r["loss_EG"], r["loss_D_valid"], r["loss_D_fake"], r["DG_mean"], r["DE_mean"]])
    # (1) eager, RNG drawn inside the step
    torch.manual_seed(1234)
    eager = []
    for _ in range(2):
        out = torch.zeros(8, device=DEV)
        tr.step(x, cd, out=out)
        eager.append(out[:5].tolist())
    w_eager = tr.gEG.flat.clone()
    for got, ref in zip(eager, want):
        for a, b in zip(got, ref):
            assert abs(a - b) <= tol * max(abs(b), 0.1), (eager, want)
    # (2) captured graph from the same state
    _restore(tr, snap)
    tr.capture(x, cd, warmup=1)
    _restore(tr, snap)
    torch.manual_seed(1234)
    replay = []
    for _ in range(2):
        tr.static["out"].zero_()
        out = tr.replay(x, cd)
        torch.cuda.synchronize()
        replay.append(out[:5].tolist())
    print(dtype, "oracle", want, "eager", eager, "replay", replay)
    rtol = 1e-4 if dtype == "fp32" else tol          # bf16: float-atomic summation order differs run to run
    for got, ref in zip(replay, eager):
        for a, b in zip(got, ref):
            assert abs(a - b) <= rtol * max(abs(b), 0.1), (replay, eager)
    assert rel_err(tr.gEG.flat, w_eager) < (1e-5 if dtype == "fp32" else 2e-3)


def test_bench_batch_train_step_vs_oracle():
    """B = 4096, bf16: the configuration BASELINE.json quotes.  One fused step (masks and z injected) against the fp32 oracle
    on the whole batch: the five outputs within 2e-2, and E(x), G(z) after the update checked over ALL images (norm-wise and
    per image), so a wrong tile / tail of a persistent kernel at this batch size cannot hide."""
    from icf_b200.trainer import BiGANTrainer
    fam, n, seed, std = "mnist", 4096, 21, 0.05
    images, c, z, _ = golden_inputs(fam, n, seed)
    torch.manual_seed(5)
    masks6 = [R.draw_masks(fam, n) for _ in range(6)]
    o = R.BiGANOracle(fam, *(R.synth_state_dict(fam, k, seed, std) for k in "EGD"))
    r = o.train_step(images, c, z, masks6)
    want = [r["loss_EG"], r["loss_D_valid"], r["loss_D_fake"], r["DG_mean"], r["DE_mean"]]
    nets = build(fam, seed, std, "bf16")
    tr = BiGANTrainer(nets["E"], nets["G"], nets["D"], lr=1e-4)
    x, cd, zz = images.to(DEV), to_dev(c), z.to(DEV)
    out = tr.step(x, cd, zz, [[m.to(DEV) for m in ms] for ms in masks6])
    got = out[:5].tolist()
    print("B=4096 bf16 step", got, "oracle", want)
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-2 * max(abs(b), 0.1), (got, want)
    with torch.no_grad():
        ex = nets["E"](x, cd).cpu().reshape(n, -1)
        gz = nets["G"](zz, cd).cpu().reshape(n, -1)
        rex = R.encoder_fwd(fam, o.E, images, c).reshape(n, -1)
        rgz = R.generator_fwd(fam, o.G, z, c).reshape(n, -1)
    for name, a, b in (("E", ex, rex), ("G", gz, rgz)):
        assert rel_err(a, b) < 2e-2, name
        per = (a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-20)
        assert float(per.max()) < 6e-2, (name, int(per.argmax()), float(per.max()))     # every single image


def test_bench_batch_counterfactual_vs_oracle():
    """B = 65536 counterfactual batch (BASELINE.json configs[2]) in bf16: a strided sample of 64 images against the fp32
    oracle (2e-2), and batch independence as the size-independent property: the same images pushed through the pipeline in a
    batch of 4096 give bit-identical rows (E and G have no cross-sample operation), which pins every tile of the big batch."""
    from icf_b200 import synth
    from icf_b200.trainer import counterfactual
    fam, n, seed, std = "mnist", 65536, 22, 0.05
    x, a, _ = synth.mnist_batch(n, seed)
    stats = synth.mnist_attr_stats()
    images, c = synth.mnist_scale(x, a, stats)
    _, c_cf = synth.mnist_scale(x, synth.intervene_mnist(a), stats)
    nets = build(fam, seed, std, "bf16")
    E, G = nets["E"], nets["G"]
    xd, cd, ccf = images.to(DEV), to_dev(c), to_dev(c_cf)
    out = counterfactual(E, G, xd, cd, ccf)
    idx = torch.arange(0, n, n // 64)[:64]
    sds = {k: R.synth_state_dict(fam, k, seed, std) for k in "EG"}
    ref = R.counterfactual(fam, sds["E"], sds["G"], images[idx], {k: v[idx] for k, v in c.items()},
                           {k: v[idx] for k, v in c_cf.items()})
    got = out[idx.to(DEV)].cpu()
    assert rel_err(got, ref) < 2e-2
    per = (got - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
    assert float(per.max()) < 6e-2
    for lo in (0, 4096 * 7 + 128, n - 4096):
        sl = slice(lo, lo + 4096)
        part = counterfactual(E, G, xd[sl], {k: v[sl] for k, v in cd.items()}, {k: v[sl] for k, v in ccf.items()})
        assert torch.equal(part, out[sl]), lo
    assert torch.isfinite(out).all()


def test_mnist_train_entry_point():
    """image_scms.mnist.train (mnist.py:157-299): signature, return tuple, optimiser export, d_updates_per_g_update."""
    from icf_b200 import synth
    m = family_module("mnist")
    x, a, _ = synth.mnist_batch(160, 3)
    torch.manual_seed(0)
    E, G, D, optD, optE = m.train(x, a, None, None, 1, 1e-4, DEV, 0, '', 64, 3, dtype="bf16")
    assert isinstance(optD, torch.optim.Adam) and isinstance(optE, torch.optim.Adam)
    # 3 batches (64, 64, 32): phase A ran on batch 0 only, D stepped twice per batch
    stE = next(iter(optE.state.values()))
    stD = next(iter(optD.state.values()))
    assert int(stE["step"]) == 1 and int(stD["step"]) == 6
    assert int(D.state_dict()["dx.4.num_batches_tracked"]) == 2 + 3 * 4      # 2 D forwards in phase A, 4 in B/C/D per batch
    assert all(torch.isfinite(p).all() for net in (E, G, D) for p in net.parameters())


@pytest.mark.parametrize("fam", ["audio_mnist", "whalecalls", "esrf_acoustic"])
def test_spectrogram_train_entry_points(fam, tmp_path):
    """train() of the spectrogram families with the reference's positional signatures (audio_mnist.py:321-327,
    whalecalls.py:390-399, esrf_acoustic.py:263-272) and a synthetic reader; ESRF additionally resumes from a whole-module
    checkpoint (esrf_acoustic.py:276-284): the networks of the checkpoint are the ones trained and returned."""
    from icf_b200.synth import SpectrogramStream
    m = family_module(fam)
    data = SpectrogramStream(fam, 2, seed=5)
    torch.manual_seed(0)
    if fam == "audio_mnist":
        res = m.train("unused.zip", 1, 1e-4, DEV, 2, 2, '', data=data, dtype="bf16")
    elif fam == "whalecalls":
        res = m.train("nocall", "gunshot", "upcall", 1, 1e-4, DEV, 2, 2, '', None, data=data, dtype="bf16")
    else:
        ck = {k: cls().to(DEV) for k, cls in (("E", m.Encoder), ("G", m.Generator), ("D", m.Discriminator))}
        for net in ck.values():
            net.apply(lambda l: m.init_weights(l, 0.02))
        path = str(tmp_path / "esrf.tar")
        torch.save(ck, path)
        before = {k: {n_: p.detach().clone() for n_, p in net.named_parameters()} for k, net in ck.items()}
        del ck
        res = m.train("wavs", "labels", 1, 1e-4, DEV, 2, 2, '', 0.2, path, data=data, dtype="bf16")
        # an Adam step moves a weight by at most lr * (1-b1)/sqrt(1-b2) = 1.6 lr (E, G: one step; D: two); a re-initialised
        # network (std 0.001 instead of the checkpoint's 0.02) would sit ~0.02 away
        for k, net in zip("EGD", res[:3]):
            for n_, p in net.named_parameters():
                assert float((p.detach() - before[k][n_]).abs().max()) <= 4e-4, (k, n_)
    E, G, D, optD, optE = res
    assert type(E).__module__ == f"image_scms.{fam}"
    assert int(next(iter(optD.state.values()))["step"]) == 2 and int(next(iter(optE.state.values()))["step"]) == 1
    assert all(torch.isfinite(p).all() for net in (E, G, D) for p in net.parameters())


def _reference_like_mnist():
    """Objects shaped like what unpickling a checkpoint WRITTEN BY THE REFERENCE yields (train_mnist_image_scm.py:61-67 saves
    whole modules): instances of image_scms.mnist.{Encoder,Generator,Discriminator} whose __init__ never ran here and whose
    children are the reference's own torch.nn containers (mnist.py:24-40, 62-72, 92-136)."""
    m = family_module("mnist")

    def bare(cls):
        obj = cls.__new__(cls)
        nn.Module.__init__(obj)
        return obj

    def emb():
        return nn.Sequential(nn.Embedding(10, 256), nn.Unflatten(1, (1, 16, 16)), nn.Upsample(size=(28, 28)), nn.Tanh())
    E = bare(m.Encoder)
    E.digit_embedding = emb()
    E.layers = nn.Sequential(nn.Conv2d(5, 64, 3, 2, 1), nn.LeakyReLU(0.2), nn.Conv2d(64, 128, 4, 2, 1), nn.LeakyReLU(0.2),
                             nn.Conv2d(128, 256, 4, 2, 1), nn.LeakyReLU(0.2), nn.Conv2d(256, 512, 4, 2, 1), nn.LeakyReLU(0.2),
                             nn.Conv2d(512, 512, 1, 2))
    G = bare(m.Generator)
    G.digit_embedding = nn.Embedding(10, 256)
    G.layers = nn.Sequential(nn.ConvTranspose2d(771, 512, 3, 1), nn.LeakyReLU(0.2), nn.ConvTranspose2d(512, 256, 3, 2),
                             nn.LeakyReLU(0.2), nn.ConvTranspose2d(256, 128, 3, 2, 1), nn.LeakyReLU(0.2),
                             nn.ConvTranspose2d(128, 64, 3, 2, 1), nn.LeakyReLU(0.2), nn.ConvTranspose2d(64, 1, 4, 1), nn.Tanh())
    return E, G


def test_reference_written_pickle_runs_on_the_engine():
    """Whole-module checkpoints written by the reference keep working: they unpickle into our classes (same import path) with
    torch.nn children and without the attributes our __init__ sets; forward must run on the engine and match the oracle."""
    fam, n, seed = "mnist", 4, 31
    E, G = _reference_like_mnist()
    sdE, sdG = R.synth_state_dict(fam, "E", seed, 0.05), R.synth_state_dict(fam, "G", seed, 0.05)
    E.load_state_dict(sdE)
    G.load_state_dict(sdG)
    buf = io.BytesIO()
    torch.save({"E": E, "G": G}, buf)
    buf.seek(0)
    obj = torch.load(buf, map_location=DEV, weights_only=False)
    E2, G2 = obj["E"], obj["G"]
    assert "compute_dtype" not in E2.__dict__ and isinstance(E2.layers, nn.Sequential)
    images, c, z, c_cf = golden_inputs(fam, n, seed)
    with torch.no_grad():
        cf = G2(E2(images.to(DEV), to_dev(c)), to_dev(c_cf))
    ref = R.counterfactual(fam, sdE, sdG, images, c, c_cf)
    assert rel_err(cf, ref) < 1e-3


def test_load_model_state_dict_checkpoint(tmp_path):
    """mnist.load_model (mnist.py:302-313): a checkpoint of reference-layout state dicts round-trips through the drop-in classes,
    and state dicts exported by them load into nothing else than the same keys / shapes (App. A.5 of the survey)."""
    m = family_module("mnist")
    fam, n, seed = "mnist", 4, 33
    sds = {k: R.synth_state_dict(fam, k, seed, 0.05) for k in "EGD"}
    path = str(tmp_path / "ck.tar")
    torch.save({"E_state_dict": sds["E"], "G_state_dict": sds["G"], "D_state_dict": sds["D"], "epoch": 3}, path)
    E, G, D, raw = m.load_model(path, device=DEV, return_raw=True)
    assert raw["epoch"] == 3 and isinstance(E, m.Encoder) and isinstance(D, m.Discriminator)
    for net, k in ((E, "E"), (G, "G"), (D, "D")):
        got = net.state_dict()
        assert list(got.keys()) == list(sds[k].keys())
        assert all(torch.equal(got[key].cpu(), sds[k][key]) for key in got)
    images, c, z, c_cf = golden_inputs(fam, n, seed)
    E.to(DEV), G.to(DEV)
    with torch.no_grad():
        cf = G(E(images.to(DEV), to_dev(c)), to_dev(c_cf))
    assert rel_err(cf, R.counterfactual(fam, sds["E"], sds["G"], images, c, c_cf)) < 1e-3
    planes = m.continuous_feature_map(torch.tensor([[0.5], [-1.0]]), (28, 28))
    assert planes.shape == (2, 1, 28, 28) and float(planes[1].max()) == -1.0
