"""Parity of the CUDA feature-extraction path (spectral.cu, viterbi.cu, pitch.cu)
through the C ABI.  Floating point: max|a - b| / max|b| <= 1e-4 (north_star);
Viterbi indices: bit-exact on identical log inputs."""
import math

import numpy as np
import pytest
import torch

from conftest import relative_error
from oracle import dsp, inputs
from oracle import penn as oracle_penn
from oracle import viterbi as oracle_viterbi

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-4


@pytest.fixture(scope='module')
def pb():
    import promonet_b200
    return promonet_b200


###############################################################################
# STFT features
###############################################################################


def test_spectrogram_matches_reference_golden(pb, golden):
    """promonet.preprocess.spectrogram.from_audio, outputs of the reference itself"""
    g = golden('spectrogram')
    linear = pb.preprocess.spectrogram.from_audio(g['audio'])
    mels = pb.preprocess.spectrogram.from_audio(g['audio'], mels=True)
    assert linear.shape == g['linear'].shape and mels.shape == g['mels'].shape
    assert relative_error(linear, g['linear']) < TOLERANCE
    assert relative_error(mels, g['mels']) < TOLERANCE
    assert relative_error(pb.preprocess.spectrogram.linear_to_mel(g['linear']), g['mels']) < TOLERANCE


@pytest.mark.parametrize('batch,samples', [(1, 512), (3, 22050), (2, 220500), (1, 5000)])
def test_spectrogram_matches_oracle(pb, batch, samples):
    audio = inputs.audio(batch, samples, seed=samples)
    expected = dsp.magnitude(audio.double())
    actual = pb.preprocess.spectrogram.from_audio(audio[:, None])
    assert actual.shape == (batch, 513, samples // 256)
    assert relative_error(actual, expected) < TOLERANCE
    mels = pb.preprocess.spectrogram.from_audio(audio[:, None], mels=True)
    assert relative_error(mels, dsp.linear_to_mel(expected)) < TOLERANCE
    clamped = pb.preprocess.spectrogram.from_audio(
        audio[:, None], mels=True, log_dynamic_range_compression_threshold=math.log(1e-2))
    assert float(clamped.min()) >= math.log(1e-2) - 1e-6


@pytest.mark.parametrize('bands', [8, 1, None])
@pytest.mark.parametrize('samples', [22050, 4096, 110250])
def test_loudness_matches_oracle(pb, bands, samples):
    """promonet.preprocess.loudness.from_audio (librosa arithmetic, restated)"""
    audio = inputs.audio(2, samples, seed=bands or 0)
    audio[1] *= 1e-3  # a quiet utterance exercises the 1e-10 floor and top_db clamp
    actual = pb.preprocess.loudness.from_audio(audio, bands)
    for i in range(2):
        expected = dsp.loudness(audio[i:i + 1], bands)
        assert actual[i].shape == expected.shape
        assert relative_error(actual[i], expected) < TOLERANCE


def test_loudness_silence_hits_the_floor(pb):
    loudness = pb.preprocess.loudness.from_audio(torch.zeros(1, 2048), None)
    expected = dsp.loudness(torch.zeros(1, 2048), None)
    assert torch.equal(loudness.cpu(), expected)


###############################################################################
# Viterbi
###############################################################################


def viterbi_inputs(batch, frames, states, band=None, seed=0):
    rng = np.random.default_rng(seed)
    observation = rng.random((batch, frames, states)).astype(np.float32) + 1e-3
    observation /= observation.sum(-1, keepdims=True)
    transition = rng.random((states, states)).astype(np.float32)
    if band is not None:
        index = np.arange(states)
        transition[np.abs(index[:, None] - index[None]) > band] = 0.
    transition /= transition.sum(1, keepdims=True)
    initial = rng.random(states).astype(np.float32)
    initial /= initial.sum()
    return observation, transition, initial


@pytest.mark.parametrize('batch,frames,states,band', [
    (1, 1, 7, None), (3, 50, 40, None), (2, 200, 333, 12), (4, 64, 1440, 90), (1, 30, 1025, 0),
    # more than 16 utterances: two per cluster, interleaved (odd batch: the last cluster holds one);
    # band 100 is wider than the register-resident band and takes the shared-memory scan;
    # 40 utterances: three per cluster
    (19, 40, 1440, 90), (18, 33, 1440, 100), (40, 12, 1440, 90)])
def test_viterbi_bit_exact_on_log_inputs(pb, batch, frames, states, band):
    observation, transition, initial = viterbi_inputs(batch, frames, states, band, seed=frames)
    with np.errstate(divide='ignore'):
        logs = [np.log(x) for x in (observation, transition, initial)]
    lengths = np.array([max(1, frames - 3 * i) for i in range(batch)], dtype=np.int32)
    expected = oracle_viterbi.decode(logs[0], lengths, logs[1], logs[2], log_probs=True)
    actual = pb.preprocess.viterbi.from_probabilities(
        torch.from_numpy(logs[0]), torch.from_numpy(lengths), torch.from_numpy(logs[1]),
        torch.from_numpy(logs[2]), log_probs=True)
    assert actual.dtype == torch.int32
    assert np.array_equal(actual.cpu().numpy(), expected)


def test_viterbi_full_size_bit_exact(pb):
    """config 3 shape: 861 frames x 1440 pitch bins, penn's banded transition"""
    observation, _, _ = viterbi_inputs(2, 861, 1440, seed=3)
    transition = oracle_penn.transition_matrix(256 / 22050).numpy()
    initial = oracle_penn.initial_distribution().numpy()
    with np.errstate(divide='ignore'):
        logs = [np.log(x) for x in (observation, transition, initial)]
    expected = oracle_viterbi.decode(logs[0], None, logs[1], logs[2], log_probs=True)
    actual = pb.preprocess.viterbi.from_probabilities(
        torch.from_numpy(logs[0]), None, torch.from_numpy(logs[1]), torch.from_numpy(logs[2]),
        log_probs=True)
    assert np.array_equal(actual.cpu().numpy(), expected)


def path_score(path, observation, transition, initial):
    with np.errstate(divide='ignore'):
        o, a, p = (np.log(x.astype(np.float64)) for x in (observation, transition, initial))
    score = p[path[0]] + o[0, path[0]]
    for t in range(1, len(path)):
        score += a[path[t - 1], path[t]] + o[t, path[t]]
    return score


def test_viterbi_probability_inputs_decode_an_optimal_path(pb):
    """log_probs=False takes logs on the device (logf): the path must score as well
    as the CPU optimum even where a last-bit difference in a log flips a tie"""
    observation, transition, initial = viterbi_inputs(2, 120, 200, 15, seed=9)
    expected = oracle_viterbi.decode(observation, None, transition, initial)
    actual = pb.preprocess.viterbi.from_probabilities(
        torch.from_numpy(observation), None, torch.from_numpy(transition),
        torch.from_numpy(initial)).cpu().numpy()
    for b in range(2):
        assert path_score(actual[b], observation[b], transition, initial) == pytest.approx(
            path_score(expected[b], observation[b], transition, initial), rel=1e-6)
    assert (actual == expected).mean() > 0.99


def test_viterbi_rejects_bad_shapes(pb):
    with pytest.raises(ValueError):
        pb.preprocess.viterbi.from_probabilities(torch.rand(4, 5))
    with pytest.raises(ValueError):
        pb.preprocess.viterbi.from_probabilities(torch.rand(1, 4, 5), transition=torch.rand(4, 5))


###############################################################################
# Pitch / periodicity network
###############################################################################


@pytest.fixture(scope='module')
def pitch_state(pb):
    return pb.preprocess.penn.init_state(1234)


@pytest.fixture(scope='module', params=['fp32', 'bf16x3'])
def pitch_model(pb, pitch_state, request):
    """fp32 FMA convolutions, or blocks 1-5 on the tensor cores: same parity bar"""
    math = pb._lib.MATH_FP32_SIMT if request.param == 'fp32' else pb._lib.MATH_BF16X3_TC
    return pb.preprocess.penn.Model(state=pitch_state, math=math)


def test_seeded_pitch_state_equals_oracle_state(pitch_state):
    other = oracle_penn.init_state(1234)
    assert all(torch.equal(pitch_state[k], other[k]) for k in other)


@pytest.mark.parametrize('batch,samples,frame_batch', [(1, 22050, 2048), (2, 33000, 64), (1, 3000, 7), (1, 44100, 50)])
def test_pitch_pipeline_stages_match_oracle(pitch_model, pitch_state, batch, samples, frame_batch):
    audio = inputs.audio(batch, samples, seed=samples)
    pitch, periodicity, logits, bins = pitch_model(
        audio, batch_size=frame_batch, return_intermediates=True)
    transition = oracle_penn.transition_matrix(256 / 22050).numpy()
    initial = oracle_penn.initial_distribution().numpy()
    agreement = []
    for b in range(batch):
        expected_pitch, expected_periodicity, aux = oracle_penn.from_audio(pitch_state, audio[b:b + 1])
        frames = aux['logits'].shape[0]
        assert pitch.shape == (batch, frames)
        # network: resample + framing + FCNF0++ (masked logits, finite entries)
        expected_logits, _, _ = oracle_penn.postprocess(aux['logits'])
        live = torch.isfinite(expected_logits)
        assert torch.equal(torch.isfinite(logits[b].cpu()), live)
        assert relative_error(logits[b].cpu()[live], expected_logits[live]) < TOLERANCE
        # periodicity = 1 - normalised entropy
        assert (periodicity[b].cpu() - expected_periodicity[0]).abs().max() < TOLERANCE
        # decoding: the device path is an optimal path of the oracle's own posterior
        path = bins[b].cpu().numpy()
        distribution = aux['distribution'].numpy()
        assert path_score(path, distribution, transition, initial) == pytest.approx(
            path_score(aux['bins'].numpy(), distribution, transition, initial), rel=1e-5)
        agreement.append(float((path == aux['bins'].numpy()).mean()))
        # local expected value, evaluated on the device's own bins and logits
        expected_hz = oracle_penn.local_expected_value(bins[b].cpu(), logits[b].cpu())
        assert relative_error(pitch[b], expected_hz) < 1e-5
        assert float(pitch[b].min()) >= 31. and float(pitch[b].max()) <= 31. * 2 ** 6
    assert min(agreement) > 0.95, agreement


@pytest.mark.parametrize('f8', ['1', '0'])
def test_pitch_block1_with_and_without_fp8_corrections_matches_oracle(pb, pitch_state, monkeypatch, f8):
    """PMN_PITCH_F8: block 1 as fp16 x fp16 + two e4m3 correction products (conv1d_tc.cuh; the default)
    or as bf16 x 3; same bar"""
    monkeypatch.setenv('PMN_PITCH_F8', f8)
    model = pb.preprocess.penn.Model(state=pitch_state, math=pb._lib.MATH_BF16X3_TC)
    audio = inputs.audio(2, 33000, seed=5)
    _, periodicity, logits, _ = model(audio, batch_size=64, return_intermediates=True)
    for b in range(2):
        _, expected_periodicity, aux = oracle_penn.from_audio(pitch_state, audio[b:b + 1])
        expected_logits, _, _ = oracle_penn.postprocess(aux['logits'])
        live = torch.isfinite(expected_logits)
        assert relative_error(logits[b].cpu()[live], expected_logits[live]) < TOLERANCE
        assert (periodicity[b].cpu() - expected_periodicity[0]).abs().max() < TOLERANCE


def test_preprocess_from_audio_signature(pb):
    """promonet.preprocess.from_audio (preprocess/core.py:17-126)"""
    audio = inputs.audio(1, 22050, seed=1)
    loudness, pitch, periodicity = pb.preprocess.from_audio(audio, gpu=0)
    assert loudness.shape == (8, 86) and pitch.shape == (1, 86) and periodicity.shape == (1, 86)
    assert relative_error(loudness, dsp.loudness(audio, 8)) < TOLERANCE
    only = pb.preprocess.from_audio(audio, features=['periodicity', 'mels'])
    assert only[0].shape == (1, 86) and only[1].shape == (80, 86)
    with pytest.raises(NotImplementedError):
        pb.preprocess.from_audio(audio, features=['ppg'])


def test_preprocess_batch_items_are_independent(pb):
    audio = inputs.audio(3, 30000, seed=2)
    full = pb.preprocess.from_audio_batch(audio, features=['loudness', 'pitch', 'periodicity', 'mels'])
    single = pb.preprocess.from_audio_batch(audio[1:2], features=['loudness', 'pitch', 'periodicity', 'mels'])
    for a, b in zip(full, single):
        assert torch.equal(a[1:2], b)


@pytest.mark.parametrize('method', ['linear', 'nearest'])
def test_grid_sample_matches_reference_golden(golden, method):
    """promonet.edit.grid.sample (edit/grid.py:12-43): golden frozen from the reference"""
    import promonet_b200
    g = golden('grid')
    out = promonet_b200.edit.grid.sample(g['sequence'].cuda(), g['grid'].cuda(), method)
    assert out.shape == g[method].shape
    assert relative_error(out, g[method]) < 1e-6


def test_ppg_resampling_matches_oracle(tmp_path):
    """preprocess/core.py:97-103 (resample + softmax(log(p + 1e-8))) and load.ppg (load.py:172-188)"""
    import promonet_b200
    from oracle import features as oracle_features
    torch.manual_seed(4)
    ppg = torch.softmax(2. * torch.randn(40, 130), dim=-2)
    for length in (87, 130, 301):
        grid = promonet_b200.edit.grid.of_length(ppg, length)
        out = promonet_b200.edit.grid.sample(ppg.cuda(), grid, 'linear', renormalize=True)
        assert relative_error(out, oracle_features.resample_ppg(ppg, length)) < 1e-6
    torch.save(ppg, tmp_path / 'x-ppg.pt')
    loaded = promonet_b200.load.ppg(tmp_path / 'x-ppg.pt', resample_length=87)
    expected = oracle_features.grid_sample(ppg, torch.linspace(0., 129., 87))
    assert relative_error(loaded, expected) < 1e-6
    assert promonet_b200.load.ppg(tmp_path / 'x-ppg.pt', resample_length=130).shape == (40, 130)


def test_from_files_to_files(tmp_path):
    """preprocess/core.py:227-319: wav files in, feature files out, equal lengths batched"""
    import wave
    import promonet_b200
    from promonet_b200 import preprocess
    files = []
    for index, samples in enumerate((8192, 6400, 8192)):
        audio = inputs.audio(1, samples, seed=70 + index)
        file = tmp_path / f'utt{index}.wav'
        with wave.open(str(file), 'wb') as handle:
            handle.setnchannels(1)
            handle.setsampwidth(2)
            handle.setframerate(22050)
            handle.writeframes((audio[0] * 32767.).round().to(torch.int16).numpy().tobytes())
        files.append(file)
    features = ['loudness', 'pitch', 'periodicity', 'mels']
    preprocess.from_files_to_files(files, features=features)
    for file in files:
        expected = preprocess.from_file(file, features=features)
        prefix = file.parent / file.stem
        loudness = torch.load(f'{prefix}-loudness.pt')
        pitch = torch.load(f'{prefix}-viterbi-pitch.pt')
        periodicity = torch.load(f'{prefix}-viterbi-periodicity.pt')
        mels = torch.load(f'{prefix}-mels.pt')
        frames = preprocess.core.load_audio(file).shape[-1] // 256
        assert loudness.shape == (8, frames) and pitch.shape == (1, frames) and mels.shape == (80, frames)
        for saved, value in zip((loudness, pitch, periodicity, mels), expected):
            assert relative_error(saved, value.reshape(saved.shape)) < 1e-5
