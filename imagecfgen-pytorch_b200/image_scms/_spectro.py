"""Shared train loop of the spectrogram families (audio_mnist.py:376-420, whalecalls.py:453-499,
esrf_acoustic.py:333-379): phases A-D with k = 1, Adam betas (0.5, 0.9), ``init_weights`` std 0.001.

The reference's dataset readers (zip-of-wav / wav+mat / station folders -> log-spectrograms) are outside the
hot path; ``data`` is any object with the reference's ``stream(batch_size=...)`` generator
protocol yielding dict batches with an "audio" entry plus the attribute one-hots."""
import torch

from icf_b200.trainer import BiGANTrainer


def init_weights_std(layer, std=0.001):
    name = layer.__class__.__name__
    if name.startswith('Conv'):
        torch.nn.init.normal_(layer.weight, mean=0, std=std)
        if getattr(layer, "bias", None) is not None:
            torch.nn.init.constant_(layer.bias, 0)


def spectrogram_stats(data, batch_size, **stream_kw):
    """Per-frequency-bin mean / std over the stream (audio_mnist.py:347-358)."""
    mean, ss, n = 0, 0, 0
    for batch in data.stream(batch_size=batch_size, **stream_kw):
        n += 1
        mean = mean + batch["audio"].mean(dim=(0, 1)).reshape((1, 1, -1))
        ss = ss + batch["audio"].square().mean(dim=(0, 1)).reshape((1, 1, -1))
    mean, ss = (mean / n).float(), (ss / n).float()
    return mean, torch.sqrt(ss - mean.square()), n


def need_reader(data, what):
    if data is None:
        raise NotImplementedError(
            f"the dataset reader behind {what} is outside the B200 hot path (SURVEY.md §2); pass "
            "data=<object with the reference's .stream(batch_size=...) protocol> (icf_b200.synth.SpectrogramStream "
            "is a synthetic one)")
    return data


def shard_stream(stream, rank, world):
    """Data parallelism: rank r keeps the batches r, r+world, ... of the (identically ordered) stream and every rank
    runs the same number of steps (a ragged tail is dropped so that the gradient all-reduces stay matched)."""
    if world <= 1:
        yield from stream
        return
    group = []
    for batch in stream:
        group.append(batch)
        if len(group) == world:
            yield group[rank]
            group = []


def train_stream(E, G, D, data, attribute_names, image_shape, n_epochs, l_rate, device, batch_size, dtype=None,
                 process_group=None, stream_kw=None, cast=None, stds_kept=3, sync_bn=False):
    """The epoch loop; E/G/D arrive initialised (the family's train() applies init_weights and, for ESRF, the warm
    start of esrf_acoustic.py:276-284 BEFORE calling this, as the reference does)."""
    stream_kw = stream_kw or {}
    cast = cast or (lambda t: t)
    trainer = BiGANTrainer(E, G, D, lr=l_rate, betas=(0.5, 0.9), dtype=dtype, process_group=process_group,
                           sync_bn=sync_bn)
    spect_mean, spect_std, n_batches = spectrogram_stats(data, batch_size, **stream_kw)
    spect_mean, spect_std = spect_mean.to(device), spect_std.to(device)

    def spect_to_img(s):   # audio_mnist.py:361-363
        return torch.clip((s - spect_mean) / (spect_std + 1e-6), -stds_kept, stds_kept) / float(stds_kept)

    for epoch in range(n_epochs):
        D.train()
        E.train()
        G.train()
        scores = torch.zeros(8, dtype=torch.float32, device=device)
        steps = 0
        for batch in shard_stream(data.stream(batch_size=batch_size, **stream_kw), trainer.rank, trainer.world):
            images = spect_to_img(batch["audio"].reshape((-1, 1, *image_shape)).float().to(device))
            c = {k: cast(torch.clone(batch[k])).to(device) for k in attribute_names}
            trainer.step(images, c, out=scores)
            steps += 1
        s = trainer.reduce_scores(scores).tolist()
        print(s[3] / max(steps, 1), s[4] / max(steps, 1))
    trainer.finish()
    optimizer_D, optimizer_E = trainer.export_optimizers()
    return E, G, D, optimizer_D, optimizer_E


def check_width(d):
    if d != 64:
        raise ValueError("the B200 engine ships the reference's published width d=64 only")
