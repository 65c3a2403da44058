"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF (CPU, fp32).

Run in the build container only (needs /root/reference; the GPU box never reads it):

    python tests/golden/make_golden.py

It imports /root/reference/image_scms with two empty stub modules for imports the hot path never touches
(``pytorch_msssim`` used by rec_loss only, ``librosa`` used by the zip reader only), loads
seeded synthetic weights (oracle.bigan_ref.synth_state_dict) into the reference's own
Encoder / Generator / Discriminator, and records
  * forward outputs of E, G, D (eval and train mode with the seeded torch dropout stream),
  * the counterfactual G(E(x,c),c_cf),
  * autograd gradients of the BCE losses,
  * two iterations of the reference loop body (image_scms/mnist.py:220-248, transcribed call for call with the
    reference modules, nn.BCEWithLogitsLoss and torch.optim.Adam) — losses, scores and digests of every
    parameter / buffer afterwards.
The reference has no golden vectors of its own (no test-suite); these files are the pin for the oracle.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "imagecfgen-pytorch_b200"))

from oracle.bigan_ref import digest, synth_state_dict  # noqa: E402
from icf_b200 import synth  # noqa: E402

REF = os.environ.get("ICF_REFERENCE", "/root/reference")


def import_reference():
    for name, attrs in (("pytorch_msssim", {"ssim": None}), ("librosa", {}), ("librosa.core", {})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import importlib
    mods = {}
    for fam in ("mnist", "audio_mnist", "whalecalls", "esrf_acoustic"):
        mods[fam] = importlib.import_module(f"image_scms.{fam}")
    sys.path.remove(REF)
    return mods


def build(ref_mod, family, seed, std):
    nets = {}
    for net, cls in (("E", ref_mod.Encoder), ("G", ref_mod.Generator), ("D", ref_mod.Discriminator)):
        m = cls()
        m.load_state_dict(synth_state_dict(family, net, seed, std))
        nets[net] = m
    return nets


def inputs(family, n, seed):
    if family == "mnist":
        x, a, z = synth.mnist_batch(n, seed)
        images, c = synth.mnist_scale(x, a, synth.mnist_attr_stats())
        a_cf = synth.intervene_mnist(a)
        _, c_cf = synth.mnist_scale(x, a_cf, synth.mnist_attr_stats())
        return images, c, z, c_cf
    x, c, z = synth.spectro_batch(family, n, seed)
    c_cf = dict(c)
    k = sorted(synth.ATTR_DIMS[family])[0]
    c_cf[k] = torch.roll(c[k], 1, dims=1)            # swap one categorical attribute
    return x, c, z, c_cf


def grads_of(module):
    return {k: digest(p.grad) for k, p in module.named_parameters() if p.grad is not None}


def state_digest(module):
    return {k: digest(v.float()) for k, v in module.state_dict().items()}


def make(family, ref_mod, n, seed, std, steps, betas):
    torch.manual_seed(seed)
    nets = build(ref_mod, family, seed, std)
    E, G, D = nets["E"], nets["G"], nets["D"]
    images, c, z, c_cf = inputs(family, n, seed)
    out = {"family": family, "n": n, "seed": seed, "std": std, "betas": betas}
    with torch.no_grad():
        E.eval(), G.eval(), D.eval()
        ex = E(images, c)
        gz = G(z, c)
        out["E_out"] = ex.clone()
        out["G_out"] = gz.clone()
        out["D_eval_real"] = D(images, ex, c).clone()
        out["D_eval_fake"] = D(gz, z, c).clone()
        out["CF_out"] = G(E(images, c), c_cf).clone()
    # train-mode forward + gradients with the seeded dropout stream
    E.train(), G.train(), D.train()
    torch.manual_seed(seed + 1)
    bce = torch.nn.BCEWithLogitsLoss()
    valid, fake = torch.ones(n, 1), torch.zeros(n, 1)
    D_valid = D(images, E(images, c), c)
    D_fake = D(G(z, c), z, c)
    loss = (bce(D_valid, fake) + bce(D_fake, valid)) / 2
    loss.backward()
    out["train_logits_valid"] = D_valid.detach().clone()
    out["train_logits_fake"] = D_fake.detach().clone()
    out["loss_EG"] = float(loss)
    out["grads"] = {"E": grads_of(E), "G": grads_of(G), "D": grads_of(D)}
    # loop body, `steps` iterations (fresh modules so BN buffers start clean)
    nets = build(ref_mod, family, seed, std)
    E, G, D = nets["E"], nets["G"], nets["D"]
    E.train(), G.train(), D.train()
    opt_E = torch.optim.Adam(list(E.parameters()) + list(G.parameters()), lr=1e-4, betas=betas)
    opt_D = torch.optim.Adam(D.parameters(), lr=1e-4, betas=betas)
    torch.manual_seed(seed + 2)
    log = []
    for _ in range(steps):
        opt_E.zero_grad()
        D_valid = D(images, E(images, c), c)
        D_fake = D(G(z, c), z, c)
        loss_EG = (bce(D_valid, fake) + bce(D_fake, valid)) / 2
        loss_EG.backward()
        opt_E.step()
        opt_D.zero_grad()
        D_valid = D(images, E(images, c), c)
        loss_Dv = bce(D_valid, valid)
        loss_Dv.backward()
        opt_D.step()
        opt_D.zero_grad()
        D_fake = D(G(z, c), z, c)
        loss_Df = bce(D_fake, fake)
        loss_Df.backward()
        opt_D.step()
        Gz = G(z, c).detach()
        EX = E(images, c).detach()
        DG = D(Gz, z, c).sigmoid()
        DE = D(images, EX, c).sigmoid()
        log.append([float(loss_EG), float(loss_Dv), float(loss_Df), float(DG.mean()), float(DE.mean())])
    out["step_log"] = log
    out["state_after"] = {"E": state_digest(E), "G": state_digest(G), "D": state_digest(D)}
    return out


def main():
    mods = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    cases = [("mnist", 6, 11, 0.05, 2, (0.5, 0.999)),
             ("mnist", 5, 12, 0.01, 1, (0.5, 0.999)),          # as-shipped init scale (ill-conditioned smoke)
             ("mnist", 64, 16, 0.05, 2, (0.5, 0.999)),         # batch large enough for 2e-2 bf16 bounds on batch means
             ("audio_mnist", 2, 13, 0.02, 1, (0.5, 0.9)),
             ("whalecalls", 1, 14, 0.02, 1, (0.5, 0.9)),
             ("esrf_acoustic", 1, 15, 0.01, 1, (0.5, 0.9))]    # 723 M parameters: needs ~20 GB of host memory
    only = os.environ.get("ICF_GOLDEN_ONLY")
    if only:
        cases = [c for c in cases if f"{c[0]}_n{c[1]}_s{c[2]}" in only.split(",")]
    for family, n, seed, std, steps, betas in cases:
        g = make(family, mods[family], n, seed, std, steps, betas)
        path = os.path.join(HERE, f"{family}_n{n}_s{seed}.pt")
        torch.save(g, path)
        print(path, os.path.getsize(path) // 1024, "KiB", "loss_EG", g["loss_EG"], g["step_log"])


if __name__ == "__main__":
    main()
