"""Synthesis API: drop-in for promonet.synthesize.from_features
(promonet/synthesize/core.py:18-59) plus the batched entry the reference lacks
(its generate() is hard-wired to batch 1, synthesize/core.py:256-268)."""
import os
from pathlib import Path
from typing import List, Optional, Union

import torch

import promonet_b200
from promonet_b200 import _lib

__all__ = [
    'from_features', 'from_features_batch', 'from_file', 'from_file_to_file',
    'from_files_to_files', 'generate']


def from_features(
    loudness: torch.Tensor,
    pitch: torch.Tensor,
    periodicity: torch.Tensor,
    ppg: torch.Tensor,
    speaker: Union[int, torch.Tensor] = 0,
    spectral_balance_ratio: float = 1.,
    loudness_ratio: float = 1.,
    checkpoint: Optional[Union[str, os.PathLike]] = None,
    gpu: Optional[int] = None
) -> torch.Tensor:
    """Perform speech synthesis

    Args:
        loudness: The loudness contour, (8, F), (513, F) or with leading 1
        pitch: The pitch contour (1, F)
        periodicity: The periodicity contour (1, F)
        ppg: The phonetic posteriorgram (1, 40, F)
        speaker: The speaker index
        spectral_balance_ratio: > 1 for Alvin and the Chipmunks; < 1 for Patrick Star
        loudness_ratio: > 1 for louder; < 1 for quieter
        checkpoint: The generator checkpoint; None = seeded random init
            (the reference downloads pretrained weights here; no network)
        gpu: The GPU index (None = current CUDA device; there is no CPU path)

    Returns
        generated: The generated speech (1, 256 * F), float32
    """
    device = _device(gpu)
    if loudness.ndim == 2:
        loudness = loudness[None]
    return generate(
        loudness.to(device),
        pitch.to(device),
        periodicity.to(device),
        ppg.to(device),
        speaker,
        spectral_balance_ratio,
        loudness_ratio,
        checkpoint
    ).to(torch.float32)


def from_features_batch(
    loudness: torch.Tensor,
    pitch: torch.Tensor,
    periodicity: torch.Tensor,
    ppg: torch.Tensor,
    speakers: torch.Tensor,
    spectral_balance_ratios: Optional[torch.Tensor] = None,
    loudness_ratios: Optional[torch.Tensor] = None,
    checkpoint: Optional[Union[str, os.PathLike]] = None,
    gpu: Optional[int] = None
) -> torch.Tensor:
    """Batched synthesis: (B, 8|513, F), (B, F), (B, F), (B, 40, F), (B,)
    -> (B, 1, 256 F).  Host inputs go through the library's host entry (H2D,
    forward, D2H in one call) and come back on the host; device inputs stay
    on the device."""
    device = _device(gpu)
    batch = loudness.shape[0]
    if spectral_balance_ratios is None:
        spectral_balance_ratios = torch.ones(batch)
    if loudness_ratios is None:
        loudness_ratios = torch.ones(batch)
    model = _model(device, checkpoint)
    if loudness.is_cuda:
        return model(
            loudness, pitch, periodicity, ppg, speakers,
            spectral_balance_ratios, loudness_ratios)
    return model.forward_host(
        loudness, pitch, periodicity, ppg, speakers,
        spectral_balance_ratios.cpu(), loudness_ratios.cpu())


def from_file(
    loudness_file: Union[str, os.PathLike],
    pitch_file: Union[str, os.PathLike],
    periodicity_file: Union[str, os.PathLike],
    ppg_file: Union[str, os.PathLike],
    speaker: int = 0,
    spectral_balance_ratio: float = 1.,
    loudness_ratio: float = 1.,
    checkpoint: Optional[Union[str, os.PathLike]] = None,
    gpu: Optional[int] = None
) -> torch.Tensor:
    """Perform speech synthesis from features on disk (promonet/synthesize/core.py:62-111):
    `*-loudness.pt` (8 | 513, F), `*-pitch.pt` (1, F), `*-periodicity.pt` (1, F), `*-ppg.pt`
    (40, F'), the PPG resampled to the pitch's frame count (promonet/load.py:172-188)"""
    device = _device(gpu)
    loudness, pitch, periodicity, ppg = _load_features(
        loudness_file, pitch_file, periodicity_file, ppg_file, device)
    return from_features(
        loudness, pitch, periodicity, ppg[None], speaker, spectral_balance_ratio, loudness_ratio,
        checkpoint, gpu)


def from_file_to_file(
    loudness_file, pitch_file, periodicity_file, ppg_file, output_file, speaker=0,
    spectral_balance_ratio=1., loudness_ratio=1., checkpoint=None, gpu=None
) -> None:
    """Perform speech synthesis from features on disk and save (synthesize/core.py:114-155)"""
    generated = from_file(
        loudness_file, pitch_file, periodicity_file, ppg_file, speaker, spectral_balance_ratio,
        loudness_ratio, checkpoint, gpu)
    _save(output_file, generated.cpu())


def from_files_to_files(
    loudness_files: List[Union[str, os.PathLike]],
    pitch_files: List[Union[str, os.PathLike]],
    periodicity_files: List[Union[str, os.PathLike]],
    ppg_files: List[Union[str, os.PathLike]],
    output_files: List[Union[str, os.PathLike]],
    speakers: Optional[List[int]] = None,
    spectral_balance_ratio: float = 1.,
    loudness_ratio: float = 1.,
    checkpoint: Optional[Union[str, os.PathLike]] = None,
    gpu: Optional[int] = None,
    max_batch: int = 32
) -> None:
    """Perform batched speech synthesis from features on disk and save
    (synthesize/core.py:158-201, where it is a per-file loop at batch 1).  Utterances with the
    same number of frames are synthesized together, up to `max_batch` per launch; the vocoder
    is fully convolutional, so a batch gives each utterance exactly its batch-1 result."""
    device = _device(gpu)
    if speakers is None:
        speakers = [0] * len(pitch_files)
    features = [
        _load_features(*files, device) for files in zip(
            loudness_files, pitch_files, periodicity_files, ppg_files)]
    buckets = {}
    for index, (loudness, pitch, _, _) in enumerate(features):
        buckets.setdefault((pitch.shape[-1], loudness.shape[-2]), []).append(index)
    for members in buckets.values():
        for start in range(0, len(members), max_batch):
            chunk = members[start:start + max_batch]
            count = len(chunk)
            generated = from_features_batch(
                torch.stack([features[i][0] for i in chunk]),
                torch.cat([features[i][1] for i in chunk]),
                torch.cat([features[i][2] for i in chunk]),
                torch.stack([features[i][3] for i in chunk]),
                torch.tensor([int(speakers[i]) for i in chunk], device=device),
                torch.full((count,), spectral_balance_ratio, device=device),
                torch.full((count,), loudness_ratio, device=device),
                checkpoint, gpu).cpu()
            for i, audio in zip(chunk, generated):
                _save(output_files[i], audio)


def _load_features(loudness_file, pitch_file, periodicity_file, ppg_file, device):
    loudness = torch.load(loudness_file, map_location='cpu').to(device, torch.float32)
    pitch = torch.load(pitch_file, map_location='cpu').to(device, torch.float32)
    periodicity = torch.load(periodicity_file, map_location='cpu').to(device, torch.float32)
    ppg = promonet_b200.load.ppg(ppg_file, resample_length=pitch.shape[-1], device=device)
    return loudness, pitch, periodicity, ppg.to(torch.float32)


def _save(output_file, audio):
    """torchaudio.save(output_file, generated, SAMPLE_RATE) (synthesize/core.py:155): 16-bit PCM
    wav written directly (torchaudio's save backends are optional dependencies)"""
    import wave
    output_file = Path(output_file)
    output_file.parent.mkdir(exist_ok=True, parents=True)
    samples = (audio.reshape(-1).clamp(-1., 1.) * 32767.).round().to(torch.int16)
    with wave.open(str(output_file), 'wb') as file:
        file.setnchannels(1)
        file.setsampwidth(2)
        file.setframerate(promonet_b200.SAMPLE_RATE)
        file.writeframes(samples.numpy().tobytes())


def generate(
    loudness,
    pitch,
    periodicity,
    ppg,
    speaker=0,
    spectral_balance_ratio: float = 1.,
    loudness_ratio: float = 1.,
    checkpoint=None
) -> torch.Tensor:
    """Generate speech from phoneme and prosody features
    (promonet/synthesize/core.py:209-281; batch 1)"""
    device = pitch.device
    model = _model(device, checkpoint)
    speakers = torch.full((1,), int(speaker), dtype=torch.long, device=device)
    sbr = torch.tensor([spectral_balance_ratio], dtype=torch.float, device=device)
    lr = torch.tensor([loudness_ratio], dtype=torch.float, device=device)
    return model(
        loudness, pitch, periodicity, ppg, speakers, sbr, lr,
        model.default_previous_samples)[0]


###############################################################################
# Utilities
###############################################################################


def _device(gpu):
    if not torch.cuda.is_available():
        raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device() if gpu is None else gpu)


def _model(device, checkpoint):
    """Model cache keyed like the reference's (synthesize/core.py:225-248)"""
    if type(checkpoint) is str:
        checkpoint = Path(checkpoint)
    if (
        not hasattr(generate, 'model') or
        generate.checkpoint != checkpoint or
        generate.device != device
    ):
        if checkpoint is None:
            state = promonet_b200.model.init.hifigan_state(
                promonet_b200.RANDOM_SEED)
        else:
            if checkpoint.is_dir():
                checkpoint = latest_path(checkpoint, 'generator-*.pt')
            state = torch.load(checkpoint, map_location='cpu')
            # torchutil.checkpoint.save stores {'model': state_dict, ...}
            state = state.get('model', state)
        generate.model = promonet_b200.model.Generator(device=device, state=state)
        generate.checkpoint = checkpoint
        generate.device = device
    return generate.model


def latest_path(directory, pattern='generator-*.pt'):
    """torchutil.checkpoint.latest_path: highest step number in the file name"""
    files = list(Path(directory).glob(pattern))
    if not files:
        raise FileNotFoundError(f'no checkpoint matching {pattern} in {directory}')
    return max(files, key=lambda f: int(''.join(c for c in f.stem if c.isdigit()) or 0))
