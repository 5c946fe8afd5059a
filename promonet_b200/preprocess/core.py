"""Feature extraction API: drop-in for promonet.preprocess.from_audio
(promonet/preprocess/core.py:17-126) for the loudness / pitch / periodicity
branches, plus the spectrogram features the dataset driver adds
(promonet/data/preprocess/core.py:36-46) and a batched entry point."""
from typing import Optional, Tuple

import torch

from promonet_b200 import config
from promonet_b200.preprocess import loudness, penn, spectrogram

__all__ = ['from_audio', 'from_audio_batch', 'from_file', 'from_file_to_file', 'from_files_to_files']

SUPPORTED = ('loudness', 'pitch', 'periodicity', 'spectrogram', 'mels')

# Frames per pass of the pitch network.  The reference hands penn batch_size=2048
# (preprocess/core.py:77), a memory bound of its frame-by-frame network; no result depends on
# it.  Here a pass costs 0.6 MB of workspace per frame and larger passes fill the 148 SMs better
# (the last two blocks have 4 and 35 output positions per frame to spread over them).
FRAME_BATCH = 8192


def from_audio(
    audio: torch.Tensor,
    sample_rate: int = config.SAMPLE_RATE,
    gpu: Optional[int] = None,
    features: list = ['loudness', 'pitch', 'periodicity'],
    loudness_bands: Optional[int] = config.LOUDNESS_BANDS,
    max_harmonics=None,
    pitch_checkpoint=None
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
        pitch_checkpoint: FCNF0++ checkpoint for pitch / periodicity (penn's own or
            ours; default $PROMONET_B200_PITCH_CHECKPOINT; none = random init + warning)

    Returns (in this order, those requested)
        loudness (bands, F), pitch (1, F), periodicity (1, F),
        spectrogram (513, F), mels (80, F)
    """
    return tuple(
        value[0] if name in ('loudness', 'spectrogram', 'mels') else value
        for name, value in zip(
            [f for f in SUPPORTED if f in features],
            from_audio_batch(audio, sample_rate, gpu, features, loudness_bands, pitch_checkpoint)))


def from_audio_batch(
    audio: torch.Tensor,
    sample_rate: int = config.SAMPLE_RATE,
    gpu: Optional[int] = None,
    features: list = ['loudness', 'pitch', 'periodicity'],
    loudness_bands: Optional[int] = config.LOUDNESS_BANDS,
    pitch_checkpoint=None
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
        pitch, periodicity = _pitch_model(device, pitch_checkpoint)(
            audio, sample_rate, config.HOPSIZE / config.SAMPLE_RATE,
            config.FMIN, config.FMAX, FRAME_BATCH)
        if 'pitch' in features:
            result.append(pitch)
        if 'periodicity' in features:
            result.append(periodicity)
    if 'spectrogram' in features:
        result.append(spectrogram.from_audio(audio[:, None]))
    if 'mels' in features:
        result.append(spectrogram.from_audio(audio[:, None], mels=True))
    return tuple(result)


def from_file(file, gpu=None, features=['loudness', 'pitch', 'periodicity'],
              loudness_bands=config.LOUDNESS_BANDS, max_harmonics=None,
              pitch_checkpoint=None) -> Tuple:
    """Preprocess audio on disk (promonet/preprocess/core.py:129-166); 22.05 kHz PCM wav"""
    return from_audio(
        load_audio(file), gpu=gpu, features=features, loudness_bands=loudness_bands,
        pitch_checkpoint=pitch_checkpoint)


def from_file_to_file(file, output_prefix=None, gpu=None,
                      features=['loudness', 'pitch', 'periodicity'],
                      loudness_bands=config.LOUDNESS_BANDS, max_harmonics=None,
                      pitch_checkpoint=None) -> None:
    """Preprocess audio on disk and save (preprocess/core.py:169-224)"""
    from_files_to_files(
        [file], None if output_prefix is None else [output_prefix], gpu, features, loudness_bands,
        pitch_checkpoint=pitch_checkpoint)


def from_files_to_files(files, output_prefixes=None, gpu=None,
                        features=['loudness', 'pitch', 'periodicity'],
                        loudness_bands=config.LOUDNESS_BANDS, max_harmonics=None,
                        max_batch=32, pitch_checkpoint=None) -> None:
    """Preprocess multiple audio files on disk and save (preprocess/core.py:227-319): writes
    `{prefix}-loudness.pt`, `{prefix}-viterbi-pitch.pt`, `{prefix}-viterbi-periodicity.pt` (the
    names of the reference under VITERBI_DECODE_PITCH, :262-268), `{prefix}-spectrogram.pt`,
    `{prefix}-mels.pt`.  Files of equal length are processed as one batch (every feature is
    computed per utterance, so batching does not change a result)."""
    from pathlib import Path
    files = [Path(file) for file in files]
    if output_prefixes is None:
        output_prefixes = [file.parent / file.stem for file in files]
    suffixes = {
        'loudness': '-loudness.pt', 'pitch': '-viterbi-pitch.pt',
        'periodicity': '-viterbi-periodicity.pt', 'spectrogram': '-spectrogram.pt',
        'mels': '-mels.pt'}
    audio = [load_audio(file) for file in files]
    buckets = {}
    for index, item in enumerate(audio):
        buckets.setdefault(item.shape[-1], []).append(index)
    order = [f for f in SUPPORTED if f in features]
    for members in buckets.values():
        for start in range(0, len(members), max_batch):
            chunk = members[start:start + max_batch]
            results = from_audio_batch(
                torch.cat([audio[i] for i in chunk]), config.SAMPLE_RATE, gpu, features, loudness_bands,
                pitch_checkpoint)
            for name, values in zip(order, results):
                for i, value in zip(chunk, values.cpu()):
                    # pitch and periodicity are stored (1, F) like penn's outputs
                    value = value[None] if name in ('pitch', 'periodicity') else value
                    torch.save(value.clone(), f'{output_prefixes[i]}{suffixes[name]}')


def load_audio(file):
    """promonet.load.audio (promonet/load.py:16-36) for 16-bit PCM wav at the model's rate -> (1, T)"""
    import wave
    with wave.open(str(file), 'rb') as handle:
        if handle.getframerate() != config.SAMPLE_RATE or handle.getsampwidth() != 2:
            raise ValueError(
                f'{file}: expected 16-bit PCM at {config.SAMPLE_RATE} Hz (resampling of other '
                'rates is outside the accelerated path)')
        channels = handle.getnchannels()
        data = torch.frombuffer(
            bytearray(handle.readframes(handle.getnframes())), dtype=torch.int16)
    return (data.float() / 32768.).reshape(-1, channels).mean(1)[None]


PITCH_CHECKPOINT_VARIABLE = 'PROMONET_B200_PITCH_CHECKPOINT'


def _pitch_model(device, checkpoint=None):
    """The FCNF0++ network of `device`, loaded from `checkpoint` (or the file named by
    $PROMONET_B200_PITCH_CHECKPOINT).  penn's pretrained weights are downloaded by the
    reference and are not available offline: without a checkpoint the network is a seeded
    random initialisation, whose pitch is meaningless outside tests and benchmarks -- say so."""
    import os
    import warnings
    checkpoint = checkpoint or os.environ.get(PITCH_CHECKPOINT_VARIABLE) or None
    if not hasattr(_pitch_model, 'models'):
        _pitch_model.models = {}
    key = (device, None if checkpoint is None else str(checkpoint))
    if key not in _pitch_model.models:
        if checkpoint is None:
            warnings.warn(
                'promonet_b200.preprocess: no FCNF0++ checkpoint given (pitch_checkpoint= or '
                f'${PITCH_CHECKPOINT_VARIABLE}): pitch and periodicity come from a RANDOMLY '
                'INITIALISED network (seed 1234) and are only good for tests and benchmarks',
                RuntimeWarning, stacklevel=3)
            state = None
        else:
            state = penn.load_checkpoint(checkpoint)
        _pitch_model.models[key] = penn.Model(device=device, state=state)
    return _pitch_model.models[key]
