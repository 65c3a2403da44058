"""Mirror of image_scms/mnist.py: MorphoMNIST conditional BiGAN (reference :12-313) on the B200 engine."""
from typing import Dict

import torch

from icf_b200.dp import shard_permutation
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
          *, dtype=None, process_group=None, sync_bn=False):
    """Same contract as mnist.train (mnist.py:157-299): returns (E, G, D, optimizer_D, optimizer_E).

    The inner loop is ``BiGANTrainer.step``; per-epoch scores are accumulated on the device and read once
    per epoch (the reference syncs twice per step, mnist.py:247-248).  The matplotlib panel the reference saves
    every ``save_images_every`` epochs (mnist.py:251-297) is plotting, outside the hot path: the argument is accepted
    and ignored.  Keyword-only additions: ``dtype`` ('fp32'|'bf16'), ``process_group`` (data parallel: rank 0's
    initial weights are broadcast, each epoch's permutation is drawn by rank 0 and sharded so that the ranks see
    disjoint batches of ``batch_size``), ``sync_bn``."""
    E = Encoder().to(device)
    G = Generator().to(device)
    D = Discriminator().to(device)
    E.apply(init_weights)
    G.apply(init_weights)
    D.apply(init_weights)
    trainer = BiGANTrainer(E, G, D, lr=l_rate, betas=(0.5, 0.999), dtype=dtype, process_group=process_group,
                           sync_bn=sync_bn)
    for epoch in range(n_epochs):
        D.train()
        E.train()
        G.train()
        scores = torch.zeros(8, dtype=torch.float32, device=device)
        num_batches = 0
        attr_stats = {k: (v.min(dim=0).values, v.max(dim=0).values) for k, v in a_train.items() if k != "digit"}
        for i, idx in enumerate(shard_permutation(len(x_train), batch_size, trainer.group)):
            num_batches += 1
            attrs = {k: v[idx] for k, v in a_train.items()}
            images = 2 * x_train[idx].reshape((-1, 1, 28, 28)).float().to(device) / 255 - 1
            c = {k: (2 * (attrs[k] - attr_stats[k][0]) / (attr_stats[k][1] - attr_stats[k][0]) - 1).float().to(device)
                 for k in attr_stats}
            c["digit"] = attrs["digit"].to(device)
            trainer.step(images, c, phase_a=(i % d_updates_per_g_update == 0), out=scores)
        s = trainer.reduce_scores(scores).tolist()
        print(s[3] / max(num_batches, 1), s[4] / max(num_batches, 1))
    trainer.finish()
    optimizer_D, optimizer_E = trainer.export_optimizers()
    return E, G, D, optimizer_D, optimizer_E


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
