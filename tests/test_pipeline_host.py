"""Host logic of the file pipelines with the GPU call replaced by a test double (no GPU):
length bucketing, batch limits, output naming and shapes of
promonet_b200.preprocess.from_files_to_files (promonet/preprocess/core.py:227-319)"""
import wave

import torch

from promonet_b200 import preprocess
from promonet_b200.preprocess import core


def write_wav(file, samples, value):
    with wave.open(str(file), 'wb') as handle:
        handle.setnchannels(1)
        handle.setsampwidth(2)
        handle.setframerate(22050)
        handle.writeframes(torch.full((samples,), value, dtype=torch.int16).numpy().tobytes())


def test_preprocess_files_are_bucketed_by_length_and_saved_under_the_reference_names(tmp_path, monkeypatch):
    calls = []

    def double(audio, sample_rate, gpu, features, loudness_bands, pitch_checkpoint=None):
        """Stands in for from_audio_batch: every feature is the utterance's first sample"""
        calls.append(tuple(audio.shape))
        batch, frames = audio.shape[0], audio.shape[1] // 256
        first = audio[:, :1]
        shapes = {'loudness': (loudness_bands,), 'pitch': (), 'periodicity': (), 'mels': (80,)}
        return tuple(
            first.reshape(batch, *([1] * (len(shapes[name]) + 1))).expand(batch, *shapes[name], frames)
            for name in core.SUPPORTED if name in features)

    monkeypatch.setattr(core, 'from_audio_batch', double)
    lengths = (5120, 2560, 5120, 5120, 2560, 5120, 5120)
    files = []
    for index, samples in enumerate(lengths):
        files.append(tmp_path / f'utt{index}.wav')
        write_wav(files[-1], samples, 1000 * (index + 1))
    prefixes = [tmp_path / 'out' / f'item{index}' for index in range(len(files))]
    (tmp_path / 'out').mkdir()
    preprocess.from_files_to_files(
        files, prefixes, features=['mels', 'pitch', 'loudness'], max_batch=3)
    # five utterances of 5120 samples in batches of at most three, two of 2560 together
    assert sorted(calls) == [(2, 2560), (2, 5120), (3, 5120)]
    for index, (prefix, samples) in enumerate(zip(prefixes, lengths)):
        value = 1000 * (index + 1) / 32768.
        loudness = torch.load(f'{prefix}-loudness.pt')
        pitch = torch.load(f'{prefix}-viterbi-pitch.pt')            # preprocess/core.py:262-268
        mels = torch.load(f'{prefix}-mels.pt')
        assert loudness.shape == (8, samples // 256) and mels.shape == (80, samples // 256)
        assert pitch.shape == (1, samples // 256)                    # stored like penn's outputs
        for tensor in (loudness, pitch, mels):
            assert torch.allclose(tensor, torch.full_like(tensor, value))   # each file got its own result
        assert not (prefix.parent / f'{prefix.name}-viterbi-periodicity.pt').exists()
    # default prefixes: next to the audio (preprocess/core.py:251-254)
    calls.clear()
    preprocess.from_file_to_file(files[1], features=['periodicity'])
    assert calls == [(1, 2560)]
    assert torch.load(tmp_path / 'utt1-viterbi-periodicity.pt').shape == (1, 10)


def test_audio_loading_rejects_other_formats(tmp_path):
    import pytest
    file = tmp_path / 'other.wav'
    with wave.open(str(file), 'wb') as handle:
        handle.setnchannels(2)
        handle.setsampwidth(2)
        handle.setframerate(16000)
        handle.writeframes(bytes(64))
    with pytest.raises(ValueError, match='22050'):
        core.load_audio(file)
    stereo = tmp_path / 'stereo.wav'
    with wave.open(str(stereo), 'wb') as handle:
        handle.setnchannels(2)
        handle.setsampwidth(2)
        handle.setframerate(22050)
        handle.writeframes(torch.tensor([16384, -16384] * 8, dtype=torch.int16).numpy().tobytes())
    audio = core.load_audio(stereo)                                   # channels are averaged
    assert audio.shape == (1, 8) and float(audio.abs().max()) == 0.
