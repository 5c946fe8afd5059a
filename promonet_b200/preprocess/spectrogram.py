"""Spectrograms: drop-in for promonet/preprocess/spectrogram.py (from_audio,
linear_to_mel); the STFT, magnitude and mel projection run in spectral.cu"""
import torch

from promonet_b200 import _lib, config

__all__ = ['from_audio', 'linear_to_mel']


def _cuda(tensor):
    if not tensor.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
        tensor = tensor.cuda()
    return tensor.to(torch.float32).contiguous()


def from_audio(audio, mels=False, log_dynamic_range_compression_threshold=None):
    """Compute spectrogram from audio (preprocess/spectrogram.py:15-60)

    audio: (1, T) or (B, 1, T) at 22 050 Hz -> (513 | 80, F) or (B, 513 | 80, F)"""
    device_audio = _cuda(audio)
    batched = device_audio.ndim == 3
    flat = device_audio.reshape(-1, device_audio.shape[-1])
    batch, samples = flat.shape
    frames = samples // config.HOPSIZE
    rows = config.NUM_MELS if mels else config.NUM_FFT // 2 + 1
    out = torch.empty(batch, rows, frames, device=flat.device)
    floor = (
        -float('inf') if log_dynamic_range_compression_threshold is None
        else log_dynamic_range_compression_threshold)
    with torch.cuda.device(flat.device):
        _lib.check(_lib.library().pmn_spectral_features(
            flat.data_ptr(), batch, samples,
            None if mels else out.data_ptr(), out.data_ptr() if mels else None, floor,
            None, 0, None, 0, _lib.stream()))
    return out if batched else out.squeeze(0)


def linear_to_mel(spectrogram, log_dynamic_range_compression_threshold=None):
    """Log-mel of a magnitude spectrogram (preprocess/spectrogram.py:111-135)"""
    device_spec = _cuda(spectrogram)
    batched = device_spec.ndim == 3
    spec = device_spec if batched else device_spec[None]
    batch, _, frames = spec.shape
    out = torch.empty(batch, config.NUM_MELS, frames, device=spec.device)
    floor = (
        -float('inf') if log_dynamic_range_compression_threshold is None
        else log_dynamic_range_compression_threshold)
    with torch.cuda.device(spec.device):
        _lib.check(_lib.library().pmn_linear_to_mel(
            spec.data_ptr(), out.data_ptr(), floor, batch, frames, _lib.stream()))
    return out if batched else out[0]
