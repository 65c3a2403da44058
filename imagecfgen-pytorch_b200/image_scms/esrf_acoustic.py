"""Mirror of image_scms/esrf_acoustic.py (esrf_acoustic.py:134-447): spectrogram conditional BiGAN on the B200 engine.
The dataset reader of the reference file is outside the hot path; train takes a data object with the
reference's stream(batch_size=...) protocol instead of opening files."""
import torch

from icf_b200.modules import DiscriminatorBase, EncoderBase, GeneratorBase
from icf_b200.trainer import BiGANTrainer, counterfactual  # noqa: F401
from ._spectro import check_width, init_weights_std, train_stream

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


def train(station_dirs=None, n_epochs=200, l_rate=1e-4, device='cpu', save_images_every=2, image_output_path='', batch_size=64, start_model_path=None, data=None, dtype=None, process_group=None):
    """Returns (E, G, D, optimizer_D, optimizer_E) like the reference.  data must provide
    stream(batch_size=...); opening station_dirs itself is the out-of-scope dataset reader."""
    if data is None:
        raise NotImplementedError(
            "the dataset reader for station_dirs is outside the B200 hot path; pass data=<object with .stream()>")
    E, G, D = Encoder().to(device), Generator().to(device), Discriminator().to(device)
    if start_model_path is not None:
        obj = torch.load(start_model_path, map_location=device, weights_only=False)   # esrf_acoustic.py:280-284
        E.load_state_dict(obj["E"].state_dict())
        G.load_state_dict(obj["G"].state_dict())
        D.load_state_dict(obj["D"].state_dict())
    names = [k for k in ATTRIBUTE_DIMS]
    return train_stream(E, G, D, data, names, IMAGE_SHAPE, n_epochs, l_rate, device, batch_size, dtype=dtype,
                        process_group=process_group, stream_kw={})
