"""Pin the CPU oracle: against the golden vectors frozen from the reference,
and against the reference itself when /root/reference is present"""
import pytest
import torch

from conftest import relative_error
from oracle import dsp, features, hifigan, inputs, ref_shim
from promonet_b200.model import init


@pytest.fixture(scope='module')
def state():
    return init.hifigan_state(1234)


def test_seeded_init_matches_reference_checksums(golden, state):
    g = golden('generator')
    for name, value in zip(g['checksum_names'], g['checksum_values']):
        assert float(state[str(name)].double().abs().sum()) == pytest.approx(
            float(value), rel=1e-12), name


@pytest.mark.parametrize('tag', ['r8', 'r513'])
def test_generator_oracle_matches_reference_output(golden, state, tag):
    g = golden('generator')
    args = [g[f'{tag}_{k}'] for k in (
        'loudness', 'pitch', 'periodicity', 'ppg', 'speakers', 'sbr', 'lr')]
    with torch.no_grad():
        feats = features.prepare_features(state, *args[:4])
        audio = hifigan.generator(state, *args)
    assert torch.equal(feats, g[f'{tag}_features'])
    assert relative_error(audio, g[f'{tag}_audio']) < 1e-5


@pytest.mark.parametrize('tag,kernel', [('c32k3', 3), ('c64k11', 11)])
def test_block_oracle_matches_reference(golden, tag, kernel):
    g = golden('block')
    state = {
        k[len(tag) + 1:]: v for k, v in g.items()
        if k.startswith(tag) and k not in (f'{tag}_x', f'{tag}_y')}
    state = {f'b.{k}': v for k, v in state.items()}
    with torch.no_grad():
        y = hifigan.block(state, 'b', g[f'{tag}_x'], kernel)
    assert relative_error(y, g[f'{tag}_y']) < 1e-6


def test_spectrogram_oracle_matches_reference(golden):
    g = golden('spectrogram')
    linear = dsp.magnitude(g['audio'])
    assert relative_error(linear, g['linear']) < 1e-6
    assert relative_error(dsp.linear_to_mel(linear), g['mels']) < 1e-5


def test_mel_basis_cross_check():
    import torchaudio
    other = torchaudio.functional.melscale_fbanks(
        513, 0., 11025., 80, 22050, norm='slaney', mel_scale='slaney').T
    assert (other - torch.from_numpy(dsp.mel_basis())).abs().max() < 1e-6


def test_band_average_matches_reference(golden):
    g = golden('loudness_bands')
    assert torch.equal(features.band_average(g['loudness']), g['averaged'])
    assert torch.equal(features.normalize(g['loudness']), g['normalized'])


def test_sparsify_keeps_top_six():
    ppg = torch.softmax(torch.randn(3, 40, 7), dim=-2)
    sparse = features.sparsify(ppg)
    assert torch.allclose(sparse.sum(-2), torch.ones(3, 7), atol=1e-6)
    assert ((sparse > 1e-6).sum(-2) == 6).all()


def test_loudness_oracle_properties():
    audio = inputs.audio(1, 22050)
    loud = dsp.loudness(audio, bands=None)
    assert loud.shape == (513, 22050 // 256)
    assert loud.min() >= -100. and loud.max() <= 40.
    # top_db: nothing more than 80 dB below the loudest bin before weighting
    assert dsp.loudness(audio).shape == (8, 86)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree absent')
def test_oracle_against_live_reference(state):
    promonet = ref_shim.load()
    torch.manual_seed(promonet.RANDOM_SEED)
    model = promonet.model.Generator().eval()
    args = inputs.synthesis(1, 20, seed=7)
    with torch.no_grad():
        expected = model(*args, model.default_previous_samples)
        actual = hifigan.generator(model.state_dict(), *args)
    assert relative_error(actual, expected) < 1e-5
    assert all(torch.equal(v, state[k]) for k, v in model.state_dict().items())


def test_viterbi_c_oracle_matches_numpy():
    import numpy as np
    from oracle import viterbi
    rng = np.random.default_rng(0)
    observation = rng.random((3, 40, 37)).astype(np.float32)
    observation /= observation.sum(-1, keepdims=True)
    transition = rng.random((37, 37)).astype(np.float32)
    index = np.arange(37)
    transition[np.abs(index[:, None] - index[None]) > 5] = 0
    transition /= transition.sum(1, keepdims=True)
    initial = np.full(37, 1 / 37, np.float32)
    lengths = [40, 17, 1]
    a = viterbi.decode(observation, lengths, transition, initial)
    b = viterbi.decode_numpy(observation, lengths, transition, initial)
    assert np.array_equal(a, b)
    # a path that can only stay within the band
    assert (np.abs(np.diff(a[0])) <= 5).all()


def test_penn_oracle_shapes():
    from oracle import penn
    state = penn.init_state(1234)
    assert sum(v.numel() for v in state.values()) == 8934624
    audio = inputs.audio(1, 8000)
    pitch, periodicity, aux = penn.from_audio(state, audio)
    assert pitch.shape == periodicity.shape == (1, 31)
    assert aux['frames'].shape == (31, 1, 1024)
    assert float(periodicity.min()) >= 0. and float(periodicity.max()) <= 1.
    transition = penn.transition_matrix(256 / 22050)
    assert torch.allclose(transition.sum(1), torch.ones(1440))


def test_fargan_oracle_matches_reference_output(golden):
    from oracle import fargan
    g = golden('fargan')
    state = init.fargan_state(1234)
    for name, value in zip(g['checksum_names'], g['checksum_values']):
        assert float(state[str(name)].double().abs().sum()) == pytest.approx(
            float(value), rel=1e-12), name
    args = [g[k] for k in ('loudness', 'pitch', 'periodicity', 'ppg', 'speakers', 'sbr', 'lr')]
    with torch.no_grad():
        assert relative_error(fargan.generator(state, *args), g['audio']) < 1e-5
        assert relative_error(
            fargan.generator(state, *args, g['previous']), g['audio_previous']) < 1e-5


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
@pytest.mark.parametrize('method', ['linear', 'nearest'])
def test_grid_sample_oracle_matches_live_reference(method):
    promonet = ref_shim.load()
    torch.manual_seed(3)
    sequence = torch.rand(2, 40, 57)
    grid = torch.linspace(0., 56., 91)
    assert torch.equal(
        features.grid_sample(sequence, grid, method), promonet.edit.grid.sample(sequence, grid, method))


def test_grid_sample_oracle_matches_golden(golden):
    g = golden('grid')
    for method in ('linear', 'nearest'):
        assert torch.equal(features.grid_sample(g['sequence'], g['grid'], method), g[method])
