"""Mirror of image_scms/whalecalls.py (whalecalls.py:230-569): spectrogram conditional BiGAN on the B200 engine.
The dataset reader of the reference file is outside the hot path; ``train`` keeps the reference's positional
signature and takes the reader as the keyword ``data`` (an object with the reference's stream() protocol)."""
import torch

from icf_b200.modules import DiscriminatorBase, EncoderBase, GeneratorBase
from icf_b200.trainer import BiGANTrainer, counterfactual  # noqa: F401
from ._spectro import check_width, init_weights_std, need_reader, train_stream

LATENT_DIM = 512
IMAGE_SHAPE = (256, 256)
ATTRIBUTE_DIMS = {"call_type": 3}


def init_weights(layer, std=0.001):
    init_weights_std(layer, std)


class Encoder(EncoderBase):
    FAMILY = "whalecalls"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


class Generator(GeneratorBase):
    FAMILY = "whalecalls"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


class Discriminator(DiscriminatorBase):
    FAMILY = "whalecalls"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


def _fresh(device):
    E, G, D = Encoder().to(device), Generator().to(device), Discriminator().to(device)
    E.apply(init_weights)
    G.apply(init_weights)
    D.apply(init_weights)
    return E, G, D


def train(nocall_directory,
          gunshot_directory,
          upcall_directory,
          n_epochs=200,
          l_rate=1e-4,
          device='cpu',
          save_images_every=2,
          batch_size=32,
          image_output_path='',
          filter_length=None,
          *, data=None, dtype=None, process_group=None):
    """whalecalls.py:390-569 -> (E, G, D, optimizer_D, optimizer_E)."""
    data = need_reader(data, "nocall_directory / gunshot_directory / upcall_directory")
    E, G, D = _fresh(device)
    names = [k for k in ATTRIBUTE_DIMS]
    return train_stream(E, G, D, data, names, IMAGE_SHAPE, n_epochs, l_rate, device, batch_size, dtype=dtype,
                        process_group=process_group, stream_kw={},
                        cast=lambda t: t.int())            # whalecalls.py:455 feeds int32 one-hots
