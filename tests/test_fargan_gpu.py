"""FARGAN generator (config/fargan.py) on the GPU against the reference's own
output (tests/golden/fargan.npz) and the CPU oracle; tolerance 1e-4 relative."""
import pytest
import torch

from conftest import relative_error
from oracle import fargan as oracle_fargan
from oracle import inputs

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-4


@pytest.fixture(scope='module')
def state():
    from promonet_b200.model import init
    return init.fargan_state(1234)


@pytest.fixture(scope='module')
def model(state):
    import promonet_b200
    return promonet_b200.model.FarganGenerator(state=state)


def test_fargan_matches_reference_golden(model, golden):
    g = golden('fargan')
    args = [g[k].cuda() for k in ('loudness', 'pitch', 'periodicity', 'ppg', 'speakers', 'sbr', 'lr')]
    audio = model(*args)
    assert audio.shape == g['audio'].shape
    assert relative_error(audio, g['audio']) < TOLERANCE
    audio = model(*args, g['previous'].cuda())
    assert relative_error(audio, g['audio_previous']) < TOLERANCE


@pytest.mark.parametrize('batch,frames', [(1, 5), (17, 3), (33, 2), (3, 40)])
def test_fargan_matches_oracle(model, state, batch, frames):
    """1, 2 and 3 CTA groups (16 utterances each); 33 needs a second cooperative launch"""
    args = inputs.synthesis(batch, frames, seed=batch + frames)
    with torch.no_grad():
        expected = oracle_fargan.generator(state, *args)
    audio = model(*[a.cuda() for a in args])
    assert audio.shape == (batch, 1, 256 * frames)
    for i in range(batch):
        assert relative_error(audio[i], expected[i]) < TOLERANCE


def test_fargan_batch_items_are_independent(model):
    args = inputs.synthesis(5, 6, seed=1)
    full = model(*[a.cuda() for a in args])
    part = model(*[a[3:].cuda() for a in args])
    assert torch.equal(full[3:], part)
