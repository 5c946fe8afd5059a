"""MULTI_RESOLUTION_DISCRIMINATOR (promonet/model/discriminator.py:96-141) on the GPU against the
oracle and tests/golden/train_resolution.npz (the reference module's own outputs and gradients):
the DFT basis and magnitude kernels, the three DiscriminatorR forward (fp32 and tf32), their weight
gradients and the gradient with respect to the audio, and a whole training step with the flag on."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, relative_error
from oracle import train as oracle_train

pytestmark = pytest.mark.gpu

FORWARD_TOLERANCE = 1e-4
GRADIENT_TOLERANCE = 2e-3


@pytest.fixture(scope='module')
def state():
    from promonet_b200.model import init
    return init.discriminator_state(1234, multi_resolution=True)


@pytest.mark.parametrize('n_fft,win', [(512, 240), (1024, 600), (2048, 1200), (64, 64)])
def test_rectangular_dft_basis_and_magnitude(n_fft, win):
    from promonet_b200.train import ops
    basis = ops.dft_basis_rect(n_fft, win, 'cuda').cpu().double()
    bins = n_fft // 2 + 1
    window = torch.zeros(n_fft, dtype=torch.float64)
    left = (n_fft - win) // 2
    window[left:left + win] = 1.
    n = torch.arange(n_fft, dtype=torch.float64)
    phase = 2 * torch.pi * torch.arange(bins, dtype=torch.float64)[:, None] * n[None] / n_fft
    expected = torch.cat([window * torch.cos(phase), -window * torch.sin(phase)])
    assert (basis - expected).abs().max() < 1e-6
    torch.manual_seed(n_fft)
    spec = torch.randn(3, 2 * bins, 7)
    spec[0, :, 0] = 0.                                     # |X| = 0: zero subgradient
    magnitude = ops.complex_magnitude(spec.cuda())
    leaf = spec.clone().requires_grad_()
    reference = torch.sqrt(leaf[:, :bins] ** 2 + leaf[:, bins:] ** 2)
    assert relative_error(magnitude[:, 0], reference) < 1e-6
    g = torch.randn_like(reference)
    torch.norm(torch.stack([leaf[:, :bins], leaf[:, bins:]], -1), p=2, dim=-1).backward(g)
    gspec = ops.complex_magnitude_backward(g[:, None].contiguous().cuda(), spec.cuda())
    assert relative_error(gspec, leaf.grad) < 1e-6


@pytest.mark.parametrize('math,tolerance', [('fp32', FORWARD_TOLERANCE), ('tf32', 2e-2)])
def test_multi_resolution_forward_matches_reference_golden(state, math, tolerance):
    from promonet_b200.train.discriminator import Discriminator
    golden = np.load(GOLDEN / 'train_resolution.npz')
    D = Discriminator(state, math=math)
    assert [type(m).__name__ for m in D.modules[5:]] == ['Resolution'] * 3
    D.refresh()
    both = torch.cat([torch.from_numpy(golden['audio']), torch.from_numpy(golden['generated'])])
    records = D.forward(both.cuda())
    logits, maps = D.logits(records), D.feature_maps(records)
    assert len(logits) == 9 and [len(m) for m in maps] == [6] * 8 + [26]
    for index in (5, 6, 7):
        expected = torch.cat([
            torch.from_numpy(golden[f'logits_real_{index}']),
            torch.from_numpy(golden[f'logits_fake_{index}'])])
        assert logits[index].shape == expected.shape
        assert relative_error(logits[index], expected) < tolerance, index
        sums = [float(m[2:].double().abs().sum()) for m in maps[index]]
        np.testing.assert_allclose(sums, golden[f'fmap_checksums_{index}'], rtol=10 * tolerance)


def test_multi_resolution_gradients_match_reference_golden(state):
    """Discriminator step: weight gradients of the three DiscriminatorR; generator step: the
    gradient of feature matching + adversarial loss with respect to the generated audio"""
    from promonet_b200 import config
    from promonet_b200.train import ops
    from promonet_b200.train.discriminator import Discriminator
    golden = np.load(GOLDEN / 'train_resolution.npz')
    D = Discriminator(state, math='fp32')
    D.refresh()
    count = golden['audio'].shape[0]
    both = torch.cat([torch.from_numpy(golden['audio']), torch.from_numpy(golden['generated'])]).cuda()
    records = D.forward(both)
    loss = torch.zeros(2, device='cuda')
    # discriminator loss (train/loss.py:29-40) and its weight gradients
    gmaps = []
    for logits, maps in zip(D.logits(records), D.feature_maps(records)):
        glogits = torch.empty_like(logits)
        ops.mse_to_target(logits[:count], 1., 1., loss[0:1], glogits[:count])
        ops.mse_to_target(logits[count:], 0., 1., loss[0:1], glogits[count:])
        gmaps.append([None] * (len(maps) - 1) + [glogits.view(maps[-1].shape)])
    D.layers.zero_grad()
    D.backward(records, gmaps, 0, 2 * count, weights=True)
    gradients = D.params.gradients()
    names = [str(n) for n in golden['names']]
    norms = np.array([float(gradients[n].double().norm()) for n in names])
    np.testing.assert_allclose(norms, golden['grad_norms'], rtol=GRADIENT_TOLERANCE)
    # generator-side losses (:11-26, :43-53) and their gradient with respect to the audio
    ggenerated = torch.zeros(count, 1, both.shape[-1], device='cuda')
    gmaps = []
    for logits, maps in zip(D.logits(records), D.feature_maps(records)):
        per_map = []
        for fmap in maps:
            g = torch.empty_like(fmap[count:])
            ops.l1_mean(fmap[count:], fmap[:count], config.FEATURE_MATCHING_LOSS_WEIGHT, loss[1:2], g)
            per_map.append(g)
        gadversarial = torch.empty_like(logits[count:])
        ops.mse_to_target(logits[count:], 1., config.ADVERSARIAL_LOSS_WEIGHT, loss[1:2], gadversarial)
        ops.axpby(1., gadversarial.view(-1), 1., per_map[-1].view(-1))
        gmaps.append(per_map)
    D.backward(records, gmaps, count, 2 * count, weights=False, gaudio=ggenerated)
    np.testing.assert_allclose(loss.cpu().numpy(), golden['losses'], rtol=1e-4)
    assert relative_error(ggenerated, torch.from_numpy(golden['generated_grad'])) < GRADIENT_TOLERANCE


def test_training_step_with_the_flag_matches_autograd():
    from promonet_b200.model import init
    from promonet_b200.train.core import Trainer
    states = init.hifigan_state(1234), init.discriminator_state(1234, multi_resolution=True)
    batch = oracle_train.batch(2, 8, seed=23)
    g_state = oracle_train.leaf_state(states[0], torch.float64)
    d_state = oracle_train.leaf_state(states[1], torch.float64)
    losses, g_grads, d_grads, _ = oracle_train.step(
        g_state, d_state, [t.double() if t.is_floating_point() else t for t in batch])
    trainer = Trainer(*states, math='fp32')
    ours = trainer.step(*[t.cuda().contiguous() for t in batch], update=False).cpu()
    for i, name in enumerate(('discriminator', 'mel', 'feature_matching', 'adversarial', 'generator')):
        assert abs(float(ours[i]) - float(losses[name])) < 1e-4 * abs(float(losses[name])), name
    for module, expected in ((trainer.discriminators, d_grads), (trainer.generator, g_grads)):
        gradients = module.params.gradients()
        errors = sorted(relative_error(gradients[n], g) for n, g in expected.items())
        assert errors[len(errors) // 2] < GRADIENT_TOLERANCE / 5 and errors[-1] < 5e-2
