"""Feature extraction API: drop-in for promonet.preprocess.from_audio
(promonet/preprocess/core.py:17-126) for the loudness / pitch / periodicity
branches, plus the spectrogram features the dataset driver adds
(promonet/data/preprocess/core.py:36-46) and a batched entry point."""
from typing import Optional, Tuple

import torch

from promonet_b200 import config
from promonet_b200.preprocess import loudness, penn, spectrogram

__all__ = ['from_audio', 'from_audio_batch']

SUPPORTED = ('loudness', 'pitch', 'periodicity', 'spectrogram', 'mels')


def from_audio(
    audio: torch.Tensor,
    sample_rate: int = config.SAMPLE_RATE,
    gpu: Optional[int] = None,
    features: list = ['loudness', 'pitch', 'periodicity'],
    loudness_bands: Optional[int] = config.LOUDNESS_BANDS,
    max_harmonics=None
) -> Tuple:
    """Preprocess audio

    Arguments
        audio: Audio to preprocess, (1, T)
        sample_rate: Audio sample rate
        gpu: The GPU index (None = current CUDA device)
        features: Any of 'loudness', 'pitch', 'periodicity', 'spectrogram', 'mels'.
            The reference's 'ppg', 'text', 'harmonics' and 'speaker' branches
            wrap foreign pretrained models and are out of scope: asking for
            them raises.
        loudness_bands: The number of A-weighted loudness bands

    Returns (in this order, those requested)
        loudness (bands, F), pitch (1, F), periodicity (1, F),
        spectrogram (513, F), mels (80, F)
    """
    return tuple(
        value[0] if name in ('loudness', 'spectrogram', 'mels') else value
        for name, value in zip(
            [f for f in SUPPORTED if f in features],
            from_audio_batch(audio, sample_rate, gpu, features, loudness_bands)))


def from_audio_batch(
    audio: torch.Tensor,
    sample_rate: int = config.SAMPLE_RATE,
    gpu: Optional[int] = None,
    features: list = ['loudness', 'pitch', 'periodicity'],
    loudness_bands: Optional[int] = config.LOUDNESS_BANDS
) -> Tuple:
    """Batched from_audio over equal-length utterances audio (B, T):
    loudness (B, bands, F), pitch (B, F), periodicity (B, F), spectrogram
    (B, 513, F), mels (B, 80, F)"""
    unsupported = [f for f in features if f not in SUPPORTED]
    if unsupported:
        raise NotImplementedError(
            f'features {unsupported} are outside the B200 hot path (foreign pretrained '
            f'models in the reference); supported: {SUPPORTED}')
    if not torch.cuda.is_available():
        raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
    device = torch.device('cuda', torch.cuda.current_device() if gpu is None else gpu)
    audio = audio.to(device, torch.float32)
    result = []
    spectral = [f for f in ('loudness', 'spectrogram', 'mels') if f in features]
    if spectral and sample_rate != config.SAMPLE_RATE:
        raise ValueError(f'spectral features expect {config.SAMPLE_RATE} Hz audio')
    if 'loudness' in features:
        value = loudness.from_audio(audio, loudness_bands)
        result.append(value if value.ndim == 3 else value[None])
    if 'pitch' in features or 'periodicity' in features:
        pitch, periodicity = _pitch_model(device)(
            audio, sample_rate, config.HOPSIZE / config.SAMPLE_RATE,
            config.FMIN, config.FMAX, 2048)
        if 'pitch' in features:
            result.append(pitch)
        if 'periodicity' in features:
            result.append(periodicity)
    if 'spectrogram' in features:
        result.append(spectrogram.from_audio(audio[:, None]))
    if 'mels' in features:
        result.append(spectrogram.from_audio(audio[:, None], mels=True))
    return tuple(result)


def _pitch_model(device):
    if not hasattr(_pitch_model, 'models'):
        _pitch_model.models = {}
    if device not in _pitch_model.models:
        _pitch_model.models[device] = penn.Model(device=device)
    return _pitch_model.models[device]
