"""A-weighted loudness: drop-in for promonet/preprocess/loudness.py (from_audio,
band_average, normalize); the STFT, dB conversion, top_db clamp, A-weighting and
band means run in spectral.cu (the reference does this on the CPU in numpy)"""
import torch

from promonet_b200 import _lib, config

__all__ = ['from_audio', 'band_average', 'normalize']


def from_audio(audio, bands=1):
    """Compute A-weighted loudness (preprocess/loudness.py:17-55)

    audio: (1, T), or (B, T) for a batch of equal-length utterances.
    Returns (bands, F) [(513, F) when bands is None], batched input -> (B, ., F)."""
    if not audio.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
        audio = audio.cuda()
    audio = audio.to(torch.float32).contiguous()
    batch, samples = audio.shape
    frames = samples // config.HOPSIZE
    rows = config.NUM_FFT // 2 + 1 if bands is None else int(bands)
    out = torch.empty(batch, rows, frames, device=audio.device)
    lib = _lib.library()
    with torch.cuda.device(audio.device):
        size = lib.pmn_spectral_workspace_bytes(batch, samples)
        workspace = torch.empty(size, dtype=torch.uint8, device=audio.device)
        _lib.check(lib.pmn_spectral_features(
            audio.data_ptr(), batch, samples, None, None, 0., out.data_ptr(),
            0 if bands is None else int(bands), workspace.data_ptr(), size, _lib.stream()))
    return out[0] if batch == 1 else out


def band_average(loudness, bands=config.LOUDNESS_BANDS):
    """Average over frequency bands (preprocess/loudness.py:84-111); shape glue for
    callers that hold a 513-row tensor -- the generator's feature kernel does this
    on the device for the synthesis path"""
    if bands is None:
        return loudness
    step = loudness.shape[-2] / bands
    return torch.stack(
        [
            loudness[..., int(band * step):int((band + 1) * step), :].mean(dim=-2)
            for band in range(int(bands))
        ],
        dim=-2)


def normalize(loudness):
    """Normalize loudness to [-1., 1.] (preprocess/loudness.py:144-146)"""
    return (loudness - config.MIN_DB) / (config.REF_DB - config.MIN_DB)
