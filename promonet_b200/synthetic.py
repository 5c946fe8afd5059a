"""Seeded synthetic inputs of the benchmark workloads, generated on
the CPU so that every path -- this library, the CPU oracle, the reference --
sees identical bits.  Shapes follow the reference's data contract
(promonet/data/collate.py:43-60, promonet/synthesize/core.py:18-59)."""
import torch


def synthesis(batch, frames, seed=1234, loudness_rows=8):
    """loudness U(-80, 0) dB, pitch log-uniform 50-550 Hz, periodicity U(0, 1),
    ppg softmax(2 N(0, 1)), speakers U{0..108}, ratios 2^U(-1, 1)"""
    generator = torch.Generator().manual_seed(seed)
    rand = lambda *shape: torch.rand(*shape, generator=generator)
    loudness = rand(batch, loudness_rows, frames) * 80. - 80.
    pitch = 50. * 11. ** rand(batch, frames)
    periodicity = rand(batch, frames)
    ppg = torch.softmax(
        2. * torch.randn(batch, 40, frames, generator=generator), dim=-2)
    speakers = torch.randint(0, 109, (batch,), generator=generator)
    sbr = 2. ** (rand(batch) * 2. - 1.)
    lr = 2. ** (rand(batch) * 2. - 1.)
    return loudness, pitch, periodicity, ppg, speakers, sbr, lr


def audio(batch, samples, seed=1234):
    """0.1 N(0, 1) noise plus a 50-550 Hz sinusoid of amplitude 0.3, clipped"""
    generator = torch.Generator().manual_seed(seed)
    noise = 0.1 * torch.randn(batch, samples, generator=generator)
    frequency = 50. * 11. ** torch.rand(batch, 1, generator=generator)
    time = torch.arange(samples)[None] / 22050.
    tone = 0.3 * torch.sin(2 * torch.pi * frequency * time)
    return torch.clip(noise + tone, -1., 1.)


def training(batch, frames, seed=1234):
    """One training batch without its target spectrograms: the seven generator
    inputs with 513-row loudness (the dataset layout) and audio (B, 1, 256 F).
    The spectrograms are a function of the audio
    (promonet/data/dataset.py:91-117): the caller computes them with whichever
    implementation it is measuring."""
    inputs = synthesis(batch, frames, seed=seed, loudness_rows=513)
    return (*inputs, audio(batch, frames * 256, seed=seed + 1)[:, None])
