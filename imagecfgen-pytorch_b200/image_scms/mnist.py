"""Mirror of image_scms/mnist.py: MorphoMNIST conditional BiGAN (reference :12-313) on the B200 engine."""
from typing import Dict

import numpy as np
import torch

from icf_b200.modules import DiscriminatorBase, EncoderBase, GeneratorBase
from icf_b200.trainer import BiGANTrainer, counterfactual  # noqa: F401
from .training_utils import AdversariallyLearnedInference  # noqa: F401
from .training_utils import batchify, batchify_dict, init_weights

LATENT_DIM = 512
N_CONTINUOUS = 3
AttributeDict = Dict[str, torch.Tensor]


def continuous_feature_map(c: torch.Tensor, size: tuple = (28, 28)):
    """(N,) / (N,1) -> (N,1,H,W) constant planes (mnist.py:17-18); the engine builds these planes inside its
    feature kernel, this helper is kept for callers."""
    return c.reshape((c.size(0), 1, 1, 1)).expand(-1, 1, *size).contiguous()


class Encoder(EncoderBase):
    """E(X (N,1,28,28), c) -> (N,512,1,1)  (mnist.py:21-56)."""
    FAMILY = "mnist"


class Generator(GeneratorBase):
    """G(z (N,512,1,1), c) -> (N,1,28,28) in (-1,1)  (mnist.py:59-86)."""
    FAMILY = "mnist"


class Discriminator(DiscriminatorBase):
    """D(X, z, c) -> logits (N,1)  (mnist.py:89-154)."""
    FAMILY = "mnist"


def train(x_train: torch.Tensor,
          a_train: AttributeDict,
          x_test=None,
          a_test=None,
          n_epochs=200,
          l_rate=1e-4,
          device='cpu',
          save_images_every=2,
          image_output_path='',
          batch_size=64,
          d_updates_per_g_update=1,
          dtype=None,
          process_group=None):
    """Same contract as mnist.train (mnist.py:157-299): returns (E, G, D, optimizer_D, optimizer_E).

    The inner loop is ``BiGANTrainer.step``; per-epoch scores are accumulated on the device and read once
    per epoch (the reference syncs twice per step, mnist.py:247-248).  ``dtype`` ('fp32'|'bf16') and
    ``process_group`` are opt-in additions."""
    E = Encoder().to(device)
    G = Generator().to(device)
    D = Discriminator().to(device)
    E.apply(init_weights)
    G.apply(init_weights)
    D.apply(init_weights)
    trainer = BiGANTrainer(E, G, D, lr=l_rate, betas=(0.5, 0.999), dtype=dtype, process_group=process_group)
    for epoch in range(n_epochs):
        D.train()
        E.train()
        G.train()
        scores = torch.zeros(8, dtype=torch.float32, device=device)
        num_batches = 0
        perm = np.random.permutation(len(x_train))
        img_batches = batchify(x_train[perm], batch_size=batch_size)
        attr_batches = batchify_dict({k: v[perm] for k, v in a_train.items()}, batch_size=batch_size)
        attr_stats = {k: (v.min(dim=0).values, v.max(dim=0).values) for k, v in a_train.items() if k != "digit"}
        for i, ((images,), attrs) in enumerate(zip(img_batches, attr_batches)):
            num_batches += 1
            images = 2 * images.reshape((-1, 1, 28, 28)).float().to(device) / 255 - 1
            c = {k: (2 * (attrs[k] - attr_stats[k][0]) / (attr_stats[k][1] - attr_stats[k][0]) - 1).float().to(device)
                 for k in attr_stats}
            c["digit"] = attrs["digit"].to(device)
            trainer.step(images, c, phase_a=(i % d_updates_per_g_update == 0), out=scores)
        s = scores.tolist()
        print(s[3] / max(num_batches, 1), s[4] / max(num_batches, 1))
        if save_images_every and (epoch + 1) % save_images_every == 0 and x_test is not None:
            _save_panel(E, G, x_test, a_test, attr_stats, device, epoch, image_output_path)
    optimizer_D, optimizer_E = trainer.export_optimizers()
    return E, G, D, optimizer_D, optimizer_E


def _save_panel(E, G, x_test, a_test, attr_stats, device, epoch, image_output_path, n_show=10):
    """Generated / real / reconstructed panel (mnist.py:251-297); skipped when matplotlib is unavailable."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    with torch.no_grad():
        xdemo = x_test[:n_show]
        ademo = {k: v[:n_show] for k, v in a_test.items()}
        c = {k: (2 * (ademo[k] - attr_stats[k][0]) / (attr_stats[k][1] - attr_stats[k][0]) - 1).float().to(device)
             for k in attr_stats}
        c["digit"] = ademo["digit"].to(device)
        x = 2 * xdemo.reshape((-1, 1, 28, 28)).float().to(device) / 255 - 1
        z = torch.randn(len(x), 512, 1, 1, device=device)
        gener = G(z, c).reshape(n_show, 28, 28).cpu().numpy()
        recon = G(E(x, c), c).reshape(n_show, 28, 28).cpu().numpy()
        real = 2 * xdemo.cpu().numpy() / 255 - 1
    fig, ax = plt.subplots(3, n_show, figsize=(15, 5))
    fig.suptitle('Epoch {}'.format(epoch + 1))
    for i in range(n_show):
        for r, im in enumerate((gener[i], real[i].reshape(28, 28), recon[i])):
            ax[r, i].imshow(im, cmap='gray', vmin=-1, vmax=1)
            ax[r, i].axis('off')
    plt.savefig(f'{image_output_path}/epoch-{epoch + 1}.png')
    plt.close()


def load_model(tar_path, device='cpu', return_raw=False):
    """mnist.py:302-313: build E/G/D and load ``{E,G,D}_state_dict`` from a checkpoint file."""
    obj = torch.load(tar_path, map_location=device, weights_only=False)
    E, G, D = Encoder(), Generator(), Discriminator()
    E.load_state_dict(obj['E_state_dict'])
    G.load_state_dict(obj['G_state_dict'])
    D.load_state_dict(obj['D_state_dict'])
    if return_raw:
        return E, G, D, obj
    return E, G, D
