"""The oracle (oracle/bigan_ref.py) against the golden fixtures produced by running the reference itself
(tests/golden/make_golden.py), and oracle/np_ops.py against the torch primitives."""
import glob
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bigan_ref as R
from oracle import np_ops
from helpers import golden_inputs, rel_err, digest_close

torch.set_num_threads(max(1, os.cpu_count() or 1))
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")) if not os.path.basename(p).startswith("explain_"))
# ESRF (723 M parameters): its forward is replayed in the default CPU suite; gradients and the train step of that family
# (~15 GB of host memory, about a minute) only with ICF_HEAVY=1 — loss, Adam and loop body are family-independent and
# pinned by the other three families
GOLD_LIGHT = [p for p in GOLD if "esrf" not in p or os.environ.get("ICF_HEAVY")]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_forward_matches_reference(path):
    g = torch.load(path, weights_only=False)
    fam, n, seed, std = g["family"], g["n"], g["seed"], g["std"]
    images, c, z, c_cf = golden_inputs(fam, n, seed)
    sd = {k: R.synth_state_dict(fam, k, seed, std) for k in "EGD"}
    with torch.no_grad():
        ex = R.encoder_fwd(fam, sd["E"], images, c)
        gz = R.generator_fwd(fam, sd["G"], z, c)
        assert rel_err(ex, g["E_out"]) < 1e-6
        assert rel_err(gz, g["G_out"]) < 1e-6
        d_real = R.discriminator_fwd(fam, sd["D"], images, ex, c, training=False)
        d_fake = R.discriminator_fwd(fam, sd["D"], gz, z, c, training=False)
        assert rel_err(d_real, g["D_eval_real"]) < 1e-5
        assert rel_err(d_fake, g["D_eval_fake"]) < 1e-5
        cf = R.counterfactual(fam, sd["E"], sd["G"], images, c, c_cf)
        assert rel_err(cf, g["CF_out"]) < 1e-6


@pytest.mark.parametrize("path", GOLD_LIGHT, ids=[os.path.basename(p) for p in GOLD_LIGHT])
def test_train_mode_and_grads_match_reference(path):
    g = torch.load(path, weights_only=False)
    fam, n, seed, std = g["family"], g["n"], g["seed"], g["std"]
    images, c, z, _ = golden_inputs(fam, n, seed)
    o = R.BiGANOracle(fam, *(R.synth_state_dict(fam, k, seed, std) for k in "EGD"), betas=tuple(g["betas"]))
    torch.manual_seed(seed + 1)                       # the seed make_golden.py set before the forward
    m1, m2 = R.draw_masks(fam, n), R.draw_masks(fam, n)
    dv = R.discriminator_fwd(fam, o.D, images, R.encoder_fwd(fam, o.E, images, c), c, m1)
    df = R.discriminator_fwd(fam, o.D, R.generator_fwd(fam, o.G, z, c), z, c, m2)
    assert rel_err(dv, g["train_logits_valid"]) < 1e-5
    assert rel_err(df, g["train_logits_fake"]) < 1e-5
    loss = (R.bce_with_logits(dv, torch.zeros(n, 1)) + R.bce_with_logits(df, torch.ones(n, 1))) / 2
    assert abs(float(loss) - g["loss_EG"]) < 1e-6 * max(1, abs(g["loss_EG"]))
    loss.backward()
    for net in "EGD":
        got = o.grads(net)
        for k, d in g["grads"][net].items():
            assert digest_close(R.digest(got[k]), d, 1e-4), (net, k)


@pytest.mark.parametrize("path", GOLD_LIGHT, ids=lambda p: os.path.basename(p))
def test_train_steps_match_reference(path):
    g = torch.load(path, weights_only=False)
    if not g["step_log"]:
        pytest.skip("no train-step record")
    fam, n, seed, std = g["family"], g["n"], g["seed"], g["std"]
    images, c, z, _ = golden_inputs(fam, n, seed)
    o = R.BiGANOracle(fam, *(R.synth_state_dict(fam, k, seed, std) for k in "EGD"), betas=tuple(g["betas"]))
    torch.manual_seed(seed + 2)
    for ref_row in g["step_log"]:
        masks6 = [R.draw_masks(fam, n) for _ in range(6)]
        out = o.train_step(images, c, z, masks6)
        row = [out["loss_EG"], out["loss_D_valid"], out["loss_D_fake"], out["DG_mean"], out["DE_mean"]]
        assert np.allclose(row, ref_row, rtol=2e-4, atol=1e-6), (row, ref_row)
    for net, sd in (("E", o.E), ("G", o.G), ("D", o.D)):
        for k, d in g["state_after"][net].items():
            assert digest_close(R.digest(sd[k].float()), d, 2e-4), (net, k)


# ---- numpy restatement of the primitives vs torch -------------------------------------------------------
def test_np_conv_and_convT():
    g = torch.Generator().manual_seed(0)
    for (C, K, H, k, s, p) in [(5, 7, 28, 3, 2, 1), (6, 4, 9, 4, 2, 1), (3, 5, 11, 5, 2, 1), (4, 4, 8, 4, 1, 0)]:
        x, w, b = torch.randn(2, C, H, H, generator=g), torch.randn(K, C, k, k, generator=g), torch.randn(K, generator=g)
        assert np.allclose(np_ops.conv2d(x, w, b, s, p), F.conv2d(x, w, b, s, p).double().numpy(), atol=1e-4)
    for (C, K, H, k, s, p, op) in [(6, 3, 3, 3, 2, 0, 0), (5, 4, 7, 3, 2, 1, 0), (4, 2, 4, 5, 2, 2, 1), (3, 1, 6, 4, 1, 0, 0)]:
        x, w, b = torch.randn(2, C, H, H, generator=g), torch.randn(C, K, k, k, generator=g), torch.randn(K, generator=g)
        ref = F.conv_transpose2d(x, w, b, s, p, op).double().numpy()
        assert np.allclose(np_ops.conv_transpose2d(x, w, b, s, p, op), ref, atol=1e-4)


def test_np_batchnorm_bce_adam_nearest():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 6, 5, 5, generator=g)
    gamma, beta = torch.randn(6, generator=g), torch.randn(6, generator=g)
    rm, rv = torch.zeros(6), torch.ones(6)
    ref = F.batch_norm(x, rm, rv, gamma, beta, True, 0.1, 1e-5)
    y, nrm, nrv = np_ops.batch_norm_train(x, gamma, beta, np.zeros(6), np.ones(6))
    assert np.allclose(y, ref.double().numpy(), atol=1e-5)
    assert np.allclose(nrm, rm.numpy(), atol=1e-6) and np.allclose(nrv, rv.numpy(), atol=1e-6)
    l, t = torch.randn(9, 1, generator=g) * 3, torch.randint(0, 2, (9, 1), generator=g).float()
    assert math.isclose(np_ops.bce_with_logits_mean(l, t), float(F.binary_cross_entropy_with_logits(l, t)), rel_tol=1e-6)
    p = torch.randn(10, generator=g, dtype=torch.float64).requires_grad_()
    opt = torch.optim.Adam([p], lr=1e-2, betas=(0.5, 0.9))
    pn, m, v = p.detach().numpy().copy(), np.zeros(10), np.zeros(10)
    for step in range(1, 4):
        gr = torch.randn(10, generator=g, dtype=torch.float64)
        p.grad = gr.clone()
        opt.step()
        pn, m, v = np_ops.adam_update(pn, gr.numpy(), m, v, step, 1e-2, 0.5, 0.9)
    assert np.allclose(pn, p.detach().numpy(), atol=1e-12)
    idx = F.interpolate(torch.arange(16.).reshape(1, 1, 1, 16), size=(1, 28), mode="nearest").long().flatten()
    assert (np_ops.nearest_index(28, 16) == idx.numpy()).all()
    assert (np_ops.nearest_index(128, 16) == np.arange(128) // 8).all()


# ---- attribute-SCM intervention (N2) pinned to the reference's data-generating SCM ------------------------------
def test_scm_closed_form_matches_the_dataset_scm():
    """create_train_dataset.py:42-46: intensity = 191*sigmoid(0.5*eps + 2t - 5) + 64.  Abduction recovers eps, regeneration
    with the factual thickness returns the observed intensity, with thickness + 2 it equals generate_i(t + 2, noise=eps)."""
    from oracle import scm_ref
    g = torch.Generator().manual_seed(3)
    t = 0.5 + torch.rand(500, 1, generator=g, dtype=torch.float64) * 4
    eps = torch.randn(500, 1, generator=g, dtype=torch.float64)
    inten = 191 * torch.sigmoid(0.5 * eps + 2 * t - 5) + 64
    closed = (-5.0, 2.0, math.log(0.5))
    v_same, e_hat = scm_ref.affine_sigmoid_cf(inten, t, t, 64.0, 191.0, closed)
    assert torch.allclose(e_hat, eps.reshape(-1), atol=1e-6) and torch.allclose(v_same, inten.reshape(-1), atol=1e-8)
    v_cf, _ = scm_ref.affine_sigmoid_cf(inten, t, t + 2, 64.0, 191.0, closed)
    assert torch.allclose(v_cf, (191 * torch.sigmoid(0.5 * eps + 2 * (t + 2) - 5) + 64).reshape(-1), atol=1e-8)
    # the synthetic pipeline's own intervention (icf_b200.synth.intervene_mnist) is the same map
    from icf_b200 import synth
    a = {"thickness": t.float(), "intensity": inten.float()}
    assert torch.allclose(synth.intervene_mnist(a)["intensity"].double().reshape(-1), v_cf, rtol=1e-4)


def test_scm_hyper_network_form_is_invertible():
    from oracle import scm_ref
    g = torch.Generator().manual_seed(4)
    hyper = {"w1": 0.3 * torch.randn(10, generator=g), "b1": 0.3 * torch.randn(10, generator=g),
             "w2": 0.3 * torch.randn(2, 10, generator=g), "b2": 0.3 * torch.randn(2, generator=g)}
    t = 0.5 + torch.rand(200, generator=g) * 4
    eps = torch.randn(200, generator=g, dtype=torch.float64)
    loc, ls = scm_ref.hyper_net(t.double(), {k: v.double() for k, v in hyper.items()}, None)
    s = loc + torch.exp(ls) * eps
    inten = 64 + 191 * torch.sigmoid(s)
    v_same, e_hat = scm_ref.affine_sigmoid_cf(inten, t, t, 64.0, 191.0, hyper=hyper)
    ok = s.abs() < 10          # beyond that the sigmoid saturates and SigmoidTransform's clamp (1 - eps) discards the noise
    assert int(ok.sum()) > 150
    assert torch.allclose(e_hat[ok], eps[ok], atol=1e-4) and torch.allclose(v_same, inten, atol=1e-4)


def test_log_spectrogram_oracle_matches_torchaudio():
    """oracle/spectro_ref.py against the reference's own call, torchaudio.transforms.Spectrogram(n_fft=255, win_length=128, pad=96)
    (audio_mnist.py:59-61) + (. + 1e-6).log() (:116)."""
    torchaudio = pytest.importorskip("torchaudio")
    from oracle import spectro_ref
    g = torch.Generator().manual_seed(6)
    wave = torch.randn(3, 8000, generator=g)
    ref = (torchaudio.transforms.Spectrogram(n_fft=255, win_length=128, pad=96)(wave) + 1e-6).log()
    got = spectro_ref.log_spectrogram(wave)
    assert got.shape == (3, 128, 128) and torch.allclose(got.float(), ref, atol=2e-4, rtol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# N3: gradient-based explainers — oracle/explain_ref.py against outputs of the reference's own classes
# (tests/golden/make_golden_explain.py ran /root/reference/explain/cf_example.py)
# ---------------------------------------------------------------------------------------------------------------------
def _explain_fixture():
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "explain_mnist_s21.pt"), weights_only=False)
    sds = {k: R.synth_state_dict("mnist", k, g["seed"], g["std"]) for k in "EG"}

    def clf(img):
        return img.flatten(1) @ g["clf_w"].t() + g["clf_b"]
    return g, sds, clf


def test_hinge_explainer_oracle_matches_reference():
    from oracle import explain_ref as X
    g, sds, clf = _explain_fixture()
    for h in g["hinge"]:
        i = h["img"]
        x = g["x"][i:i + 1]
        attrs = {k: v[i:i + 1] for k, v in g["c"].items()}
        trace = []
        x_cf, _ = X.hinge_explain("mnist", sds["E"], sds["G"], clf, x, attrs, h["init"], target_class=h["target_class"],
                                  categorical=h["categorical"], ignore=h["ignore"], train_z=h["train_z"], steps=h["steps"],
                                  lr=h["lr"], trace=trace)
        assert x_cf.shape == h["x_cf"].shape
        assert rel_err(x_cf, h["x_cf"]) < 1e-5, (i, rel_err(x_cf, h["x_cf"]))
        # the optimisation moved the image: the fixture pins a trajectory, not a fixed point
        assert rel_err(trace[0]["x_cf"], h["x_cf"]) > 1e-3


def test_deep_explainer_oracle_matches_reference():
    from oracle import explain_ref as X
    g, sds, clf = _explain_fixture()
    for d in g["deep"]:
        i = d["img"]
        x = g["x"][i:i + 1]
        attrs = {k: v[i:i + 1] for k, v in g["c"].items()}
        samples, val = X.deep_explain("mnist", sds["E"], sds["G"], clf, x, attrs, "digit", d["target_class"],
                                      sample_points=d["sample_points"], metric=d["metric"])
        assert samples.shape == d["samples"].shape and val.shape == d["val"].shape
        assert rel_err(samples, d["samples"]) < 1e-5 and rel_err(val, d["val"]) < 1e-5
