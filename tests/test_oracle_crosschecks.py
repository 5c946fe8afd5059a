"""Independent cross-checks of the parity-UNPINNED restatements (librosa, torbi, ppgs are
absent from /root/reference and from this image, so no golden vector of theirs can exist):
each is held to an implementation that IS in the image and was written by someone else
(torchaudio, torch.quantile), or to an exhaustive search.  They do not pin the oracle to the
upstream packages; they rule out errors of our own in restating the published algorithms."""
import itertools

import numpy as np
import pytest
import torch

from conftest import relative_error
from oracle import dsp, features, inputs
from oracle import penn as oracle_penn
from oracle import viterbi


def test_amplitude_to_db_matches_torchaudio():
    """librosa.amplitude_to_db(S, ref=1, amin=1e-5, top_db=80) as restated in oracle.dsp.loudness
    (loudness.py:38-46) against torchaudio.functional.amplitude_to_DB on the same spectrogram"""
    import torchaudio
    audio = inputs.audio(1, 22050, seed=3)
    audio[:, 8000:12000] *= 1e-4          # a quiet stretch reaches the top_db clamp
    frames = dsp.frames(audio[None].float()).squeeze(0).squeeze(0) * dsp.hann(torch.float32)
    magnitude = torch.fft.rfft(frames.double(), dim=-1).abs().T.float()
    expected = torchaudio.functional.amplitude_to_DB(
        magnitude.square()[None], multiplier=10., amin=1e-10, db_multiplier=0., top_db=80.)[0]
    ours = dsp.loudness(audio, bands=None)
    # undo the A-weighting and the -100 dB floor, which torchaudio does not have
    weights = torch.from_numpy(dsp.perceptual_weights()).float()
    floor = (expected + weights) < dsp.MIN_DB
    assert floor.float().mean() < 0.5
    assert torch.allclose((ours - weights)[~floor], expected[~floor], atol=1e-4)
    assert float(expected.max() - expected.min()) == pytest.approx(80., abs=1e-3)


def test_a_weighting_known_values():
    """IEC 61672 A-weighting table (what librosa.A_weighting approximates): 0 dB at 1 kHz,
    -19.1 dB at 100 Hz, +1.2 dB at 2.5 kHz, -2.5 dB at 10 kHz, floor -80 at 0 Hz"""
    values = dsp.a_weighting(np.array([0., 100., 1000., 2500., 10000.]))
    assert values[0] == -80.
    np.testing.assert_allclose(values[1:], [-19.1, 0., 1.3, -2.5], atol=0.1)


def test_sparsify_matches_torch_quantile():
    """ppgs.sparsify(ppg, 'percentile', 0.85) (generator.py:140-147): threshold = the 0.85
    quantile over the 40 classes with linear interpolation, torch.quantile's default"""
    torch.manual_seed(0)
    ppg = torch.softmax(2. * torch.randn(2, 40, 9), dim=-2)
    threshold = torch.quantile(ppg, 0.85, dim=-2, keepdim=True)
    kept = torch.where(ppg > threshold, ppg, torch.zeros(()))
    expected = torch.softmax(torch.log(kept + 1e-8), dim=-2)
    assert torch.allclose(features.sparsify(ppg, 'percentile', torch.tensor(0.85)), expected, atol=1e-7)


@pytest.mark.parametrize('states,frames,band', [(4, 6, None), (5, 5, 1), (3, 8, None)])
def test_viterbi_oracle_is_optimal_by_exhaustive_search(states, frames, band):
    """torbi's recurrence (harmonics.py:270-276) decodes the arg max over ALL state
    sequences of log pi + sum log A + sum log o: enumerate them"""
    rng = np.random.default_rng(states * 100 + frames)
    observation = rng.random((2, frames, states)).astype(np.float32) + 0.05
    observation /= observation.sum(-1, keepdims=True)
    transition = rng.random((states, states)).astype(np.float32) + 0.05
    if band is not None:
        index = np.arange(states)
        transition[np.abs(index[:, None] - index[None]) > band] = 0.
    transition /= transition.sum(1, keepdims=True)
    initial = rng.random(states).astype(np.float32) + 0.05
    initial /= initial.sum()
    decoded = viterbi.decode(observation, None, transition, initial)
    with np.errstate(divide='ignore'):
        log_o, log_a, log_p = (np.log(x.astype(np.float64)) for x in (observation, transition, initial))
    for b in range(2):
        def score(path):
            total = log_p[path[0]] + log_o[b, 0, path[0]]
            for t in range(1, frames):
                total += log_a[path[t - 1], path[t]] + log_o[b, t, path[t]]
            return total
        best = max(itertools.product(range(states), repeat=frames), key=score)
        assert score(tuple(decoded[b])) == pytest.approx(score(best), rel=1e-6, abs=1e-6)
        assert tuple(decoded[b]) == best   # random inputs: the optimum is unique


def test_viterbi_ties_resolve_to_the_lowest_index():
    """uniform inputs: every path ties; argmax by strict '>' keeps the first state"""
    observation = np.full((1, 5, 6), 1. / 6, np.float32)
    assert (viterbi.decode(observation) == 0).all()
    assert (viterbi.decode_numpy(observation) == 0).all()


def test_penn_resampling_is_torchaudio():
    """penn resamples with torchaudio.functional.resample (sinc, Hann-windowed, width 6):
    22 050 -> 8 000 Hz keeps a 440 Hz tone's frequency and amplitude"""
    time = torch.arange(22050) / 22050.
    tone = torch.sin(2 * torch.pi * 440. * time)[None]
    out = oracle_penn.resample(tone, 22050)
    assert out.shape == (1, 8000)
    expected = torch.sin(2 * torch.pi * 440. * torch.arange(8000) / 8000.)
    assert float((out[0, 100:-100] - expected[100:-100]).abs().max()) < 5e-3
