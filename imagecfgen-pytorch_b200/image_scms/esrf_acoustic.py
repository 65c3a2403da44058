"""Mirror of image_scms/esrf_acoustic.py (esrf_acoustic.py:134-447): spectrogram conditional BiGAN on the B200 engine.
The dataset reader of the reference file is outside the hot path; ``train`` keeps the reference's positional
signature and takes the reader as the keyword ``data`` (an object with the reference's stream() protocol)."""
import torch

from icf_b200.modules import DiscriminatorBase, EncoderBase, GeneratorBase
from icf_b200.trainer import BiGANTrainer, counterfactual  # noqa: F401
from ._spectro import check_width, init_weights_std, need_reader, train_stream

LATENT_DIM = 512
IMAGE_SHAPE = (512, 512)
ATTRIBUTE_DIMS = {"closest_boat": 1, "has_boat": 2}


def init_weights(layer, std=0.001):
    init_weights_std(layer, std)


class Encoder(EncoderBase):
    FAMILY = "esrf_acoustic"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


class Generator(GeneratorBase):
    FAMILY = "esrf_acoustic"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


class Discriminator(DiscriminatorBase):
    FAMILY = "esrf_acoustic"

    def __init__(self, d=64):
        check_width(d)
        super().__init__()


def _fresh(device):
    E, G, D = Encoder().to(device), Generator().to(device), Discriminator().to(device)
    E.apply(init_weights)
    G.apply(init_weights)
    D.apply(init_weights)
    return E, G, D


def train(path_to_wavs: str,
          path_to_labels: str,
          n_epochs: int = 200,
          l_rate: float = 1e-4,
          device: str = 'cpu',
          save_images_every: int = 2,
          batch_size: int = 64,
          image_output_path: str = '',
          validation_split=0.2,
          start_model_path=None,
          *, data=None, dtype=None, process_group=None):
    """esrf_acoustic.py:263-447 -> (E, G, D, optimizer_D, optimizer_E).  ``start_model_path``: whole-module pickle
    {'E','G','D'} (train_esrf_bigan.py:31-35) whose networks REPLACE the freshly initialised ones (:276-284)."""
    data = need_reader(data, "path_to_wavs / path_to_labels")
    E, G, D = _fresh(device)
    if start_model_path is not None:
        model_dict = torch.load(start_model_path, map_location=device, weights_only=False)
        E, G, D = model_dict['E'], model_dict['G'], model_dict['D']
    names = [k for k in ATTRIBUTE_DIMS]
    return train_stream(E, G, D, data, names, IMAGE_SHAPE, n_epochs, l_rate, device, batch_size, dtype=dtype,
                        process_group=process_group, stream_kw={"mode": "train"})
