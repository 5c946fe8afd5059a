"""Pin oracle/train.py (the CPU restatement of the discriminators, losses and training step)
against tests/golden/train.npz, which oracle/make_golden.py --train froze from the UNMODIFIED
reference modules; and against the live reference when /root/reference is present."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, relative_error
from oracle import make_golden, ref_shim
from oracle import train as oracle_train
from promonet_b200.model import init


@pytest.fixture(scope='module')
def golden_train():
    return np.load(GOLDEN / 'train.npz')


@pytest.fixture(scope='module')
def oracle_steps():
    """Two iterations of the oracle step with AdamW (one per golden step)"""
    torch.manual_seed(0)
    g = oracle_train.leaf_state(init.hifigan_state(1234))
    d = oracle_train.leaf_state(init.discriminator_state(1234))
    optimizers = oracle_train.make_optimizers(g, d)
    batch = oracle_train.batch(make_golden.TRAIN_BATCH, make_golden.TRAIN_FRAMES, make_golden.TRAIN_SEED)
    results = [oracle_train.step(g, d, batch, optimizers) for _ in range(make_golden.TRAIN_STEPS)]
    return batch, results, g, d


def test_batch_is_the_golden_batch(golden_train, oracle_steps):
    batch = oracle_steps[0]
    checksums = np.array([float(t.double().abs().sum()) for t in batch])
    np.testing.assert_allclose(checksums, golden_train['input_checksums'], rtol=1e-6)
    assert torch.equal(batch[8], torch.from_numpy(golden_train['audio']))


@pytest.mark.parametrize('step', range(make_golden.TRAIN_STEPS))
def test_oracle_step_matches_reference_modules(golden_train, oracle_steps, step):
    losses, g_grads, d_grads, generated = oracle_steps[1][step]
    expected = golden_train[f'losses_{step}']
    actual = [float(losses[k]) for k in (
        'discriminator', 'mel', 'feature_matching', 'adversarial', 'generator')]
    np.testing.assert_allclose(actual, expected, rtol=1e-4 if step == 0 else 1e-3)
    assert relative_error(generated, torch.from_numpy(golden_train[f'generated_{step}'])) < (
        1e-5 if step == 0 else 2e-3)
    for kind, grads in (('generator', g_grads), ('discriminator', d_grads)):
        names = [str(n) for n in golden_train[f'{kind}_names']]
        assert sorted(grads) == names
        norms = np.array([float(grads[n].double().norm()) for n in names])
        np.testing.assert_allclose(
            norms, golden_train[f'{kind}_grad_norms_{step}'], rtol=1e-3 if step == 0 else 2e-2)
        if step == 0:
            for key in golden_train.files:
                if key.startswith(f'grad:{kind}:'):
                    assert relative_error(
                        grads[key.split(':', 2)[2]], torch.from_numpy(golden_train[key])) < 1e-3, key


def test_parameters_after_two_steps(golden_train, oracle_steps):
    _, _, g, d = oracle_steps
    for kind, state in (('generator', g), ('discriminator', d)):
        names = [str(n) for n in golden_train[f'{kind}_names']]
        sums = np.array([float(state[n].detach().double().abs().sum()) for n in names])
        np.testing.assert_allclose(sums, golden_train[f'{kind}_param_checksums'], rtol=1e-5)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
def test_discriminator_init_and_forward_match_live_reference():
    promonet = ref_shim.load()
    torch.manual_seed(1234)
    reference = promonet.model.Discriminator()
    state = init.discriminator_state(1234)
    expected = reference.state_dict()
    assert list(state) == list(expected)
    assert all(torch.equal(state[k], expected[k]) for k in state)
    torch.manual_seed(5)
    y, y_hat = .3 * torch.randn(2, 1, 4096), .3 * torch.randn(2, 1, 4096)
    with torch.no_grad():
        theirs = reference(y, y_hat)
        ours = oracle_train.discriminator(state, y, y_hat)
    for a, b in zip(theirs[0] + theirs[1], ours[0] + ours[1]):
        assert relative_error(b, a) < 1e-5
    for maps_a, maps_b in zip(theirs[2] + theirs[3], ours[2] + ours[3]):
        assert len(maps_a) == len(maps_b)
        for a, b in zip(maps_a, maps_b):
            assert relative_error(b, a) < 1e-5


def test_discriminator_state_layout():
    state = init.discriminator_state(1234)
    assert len(state) == 168
    assert sum(v.numel() for v in state.values()) == 41572780
    assert state['discriminators.5.band_convs.4.1.0.weight_v'].shape == (32, 32, 3, 9)


def test_oracle_with_multi_scale_and_spectral_convergence_matches_reference():
    """tests/golden/train_flags.npz: MULTI_SCALE_DISCRIMINATOR + SPECTRAL_CONVERGENCE_LOSS on"""
    golden = np.load(GOLDEN / 'train_flags.npz')
    d_state = init.discriminator_state(1234, multi_scale=True)
    sums = np.array([
        float(d_state[k].double().abs().sum()) for k in sorted(d_state)
        if k.startswith('discriminators.5.')])
    np.testing.assert_allclose(sums, golden['msd_checksums'], rtol=1e-12)
    g = oracle_train.leaf_state(init.hifigan_state(1234))
    d = oracle_train.leaf_state(d_state)
    batch = oracle_train.batch(make_golden.TRAIN_BATCH, make_golden.TRAIN_FRAMES, make_golden.TRAIN_SEED)
    losses, g_grads, d_grads, _ = oracle_train.step(g, d, batch, spectral_convergence=True)
    actual = [float(losses[k]) for k in (
        'discriminator', 'mel', 'feature_matching', 'adversarial', 'generator', 'spectral_convergence')]
    np.testing.assert_allclose(actual, golden['losses'], rtol=1e-4)
    for kind, grads in (('generator', g_grads), ('discriminator', d_grads)):
        names = [str(n) for n in golden[f'{kind}_names']]
        norms = np.array([float(grads[n].double().norm()) for n in names])
        np.testing.assert_allclose(norms, golden[f'{kind}_grad_norms'], rtol=1e-3)


def test_multi_resolution_oracle_matches_reference_golden():
    """tests/golden/train_resolution.npz: MULTI_RESOLUTION_DISCRIMINATOR on (the
    oracle and the seeded state only — the CUDA path does not build DiscriminatorR yet)"""
    golden = np.load(GOLDEN / 'train_resolution.npz')
    plain = init.discriminator_state(1234, multi_resolution=True)
    assert oracle_train.kinds(plain) == ['p'] * 5 + ['r'] * 3 + ['cmb']
    np.testing.assert_allclose(
        [float(plain[k].double().abs().sum()) for k in sorted(plain)], golden['checksums'], rtol=1e-12)
    state = oracle_train.leaf_state(plain)
    audio = torch.from_numpy(golden['audio'])
    generated = torch.from_numpy(golden['generated']).requires_grad_()
    real_logits, fake_logits, real_maps, fake_maps = oracle_train.discriminator(state, audio, generated)
    for index in (5, 6, 7):
        assert relative_error(real_logits[index], torch.from_numpy(golden[f'logits_real_{index}'])) < 1e-5
        assert relative_error(fake_logits[index], torch.from_numpy(golden[f'logits_fake_{index}'])) < 1e-5
        np.testing.assert_allclose(
            [float(m.detach().double().abs().sum()) for m in fake_maps[index]],
            golden[f'fmap_checksums_{index}'], rtol=1e-5)
    discriminator_loss = oracle_train.discriminator_loss(real_logits, fake_logits)
    generator_loss = oracle_train.feature_matching_loss(real_maps, fake_maps) + \
        oracle_train.generator_loss(fake_logits)
    np.testing.assert_allclose(
        [float(discriminator_loss), float(generator_loss)], golden['losses'], rtol=1e-5)
    names = [str(n) for n in golden['names']]
    grads = torch.autograd.grad(discriminator_loss, [state[n] for n in names], retain_graph=True)
    np.testing.assert_allclose(
        [float(g.double().norm()) for g in grads], golden['grad_norms'], rtol=1e-3)
    gradient, = torch.autograd.grad(generator_loss, generated)
    assert relative_error(gradient, torch.from_numpy(golden['generated_grad'])) < 1e-3


def test_unbuilt_discriminator_configurations_are_refused():
    """The CUDA Discriminator refuses a state dict whose sub-discriminators it does not build
    (NotImplementedError from the key inspection, before any device work)"""
    from promonet_b200.train.discriminator import sub_discriminators
    assert sub_discriminators(init.discriminator_state(1234)) == ['p'] * 5 + ['cmb']
    assert sub_discriminators(init.discriminator_state(1234, multi_scale=True)) == \
        ['p'] * 5 + ['s', 'cmb']
    state = init.discriminator_state(1234, multi_scale=True, multi_resolution=True)
    assert sub_discriminators(state) == ['p'] * 5 + ['s'] + ['r'] * 3 + ['cmb']
    without_cmb = {k: v for k, v in state.items() if not k.startswith('discriminators.9.')}
    with pytest.raises(NotImplementedError):
        sub_discriminators(without_cmb)
    reordered = {k.replace('discriminators.0.', 'discriminators.10.'): v for k, v in state.items()}
    with pytest.raises((NotImplementedError, KeyError)):
        sub_discriminators(reordered)
