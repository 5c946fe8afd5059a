"""Host logic of the training step on the CPU: promonet_b200/train/{core,generator,discriminator}.py
run over tests/emulated_ops.py (a plain-torch double of the kernel launches, see its header)
against oracle/train.py's autograd.

1. The double is held to the oracle through the modules that ARE verified on the GPU (5 x
   multi-period + complex multi-band): if this passes, the double has the operators' semantics.
2. The same double then checks the sequencing of DiscriminatorR (forward, weight gradients,
   gradient with respect to the audio): written when no GPU time was left, checked here first,
   and green on the B200 at its first run afterwards.
3. One whole Trainer.step (generator forward, discriminator step, generator step: every loss and
   every parameter gradient of both modules) as a CPU regression test of the step's sequencing."""
import pytest
import torch

import emulated_ops
from conftest import relative_error
from oracle import train as oracle_train
from promonet_b200 import config
from promonet_b200.model import init


def run(monkeypatch, state, samples=2048, count=1, seed=0):
    """One discriminator step and one generator-side backward, emulated and by autograd"""
    from promonet_b200.train import ops
    from promonet_b200.train.discriminator import Discriminator
    emulated_ops.install(monkeypatch)
    torch.manual_seed(seed)
    audio, generated = .3 * torch.randn(count, 1, samples), .3 * torch.randn(count, 1, samples)
    # ---- oracle ----
    leaves = oracle_train.leaf_state(state, torch.float64)
    fake = generated.double().requires_grad_()
    real_logits, fake_logits, real_maps, fake_maps = oracle_train.discriminator(
        leaves, audio.double(), fake)
    d_loss = oracle_train.discriminator_loss(real_logits, fake_logits)
    names = [k for k, v in leaves.items() if v.requires_grad]
    d_grads = dict(zip(names, torch.autograd.grad(
        d_loss, [leaves[k] for k in names], retain_graph=True)))
    g_loss = config.FEATURE_MATCHING_LOSS_WEIGHT * oracle_train.feature_matching_loss(
        real_maps, fake_maps) + config.ADVERSARIAL_LOSS_WEIGHT * oracle_train.generator_loss(fake_logits)
    g_audio, = torch.autograd.grad(g_loss, fake)
    # ---- the module, over the double ----
    D = Discriminator(state, 'cpu', 'fp32')
    D.refresh()
    both = torch.cat([audio, generated])
    records = D.forward(both)
    logits, maps = D.logits(records), D.feature_maps(records)
    losses = torch.zeros(2)
    gmaps = []
    for logit, fmaps in zip(logits, maps):
        glogits = torch.empty_like(logit)
        ops.mse_to_target(logit[:count], 1., 1., losses[0:1], glogits[:count])
        ops.mse_to_target(logit[count:], 0., 1., losses[0:1], glogits[count:])
        gmaps.append([None] * (len(fmaps) - 1) + [glogits.view(fmaps[-1].shape)])
    D.layers.zero_grad()
    D.backward(records, gmaps, 0, 2 * count, weights=True)
    gradients = {k: v.clone() for k, v in D.params.gradients().items()}
    ggenerated = torch.zeros(count, 1, samples)
    gmaps = []
    for logit, fmaps in zip(logits, maps):
        per_map = []
        for fmap in fmaps:
            g = torch.empty_like(fmap[count:])
            ops.l1_mean(fmap[count:], fmap[:count], config.FEATURE_MATCHING_LOSS_WEIGHT, losses[1:2], g)
            per_map.append(g)
        gadversarial = torch.empty_like(logit[count:])
        ops.mse_to_target(logit[count:], 1., config.ADVERSARIAL_LOSS_WEIGHT, losses[1:2], gadversarial)
        ops.axpby(1., gadversarial.view(-1), 1., per_map[-1].view(-1))
        gmaps.append(per_map)
    D.backward(records, gmaps, count, 2 * count, weights=False, gaudio=ggenerated)
    # ---- comparison ----
    for i, (logit, fmaps) in enumerate(zip(logits, maps)):
        assert relative_error(logit, torch.cat([real_logits[i], fake_logits[i]])) < 1e-4, i
        assert len(fmaps) == len(real_maps[i])
        for j, fmap in enumerate(fmaps):
            assert relative_error(fmap, torch.cat([real_maps[i][j], fake_maps[i][j]])) < 1e-4, (i, j)
    assert float(losses[0]) == pytest.approx(float(d_loss.detach()), rel=1e-4)
    assert float(losses[1]) == pytest.approx(float(g_loss.detach()), rel=1e-4)
    errors = {k: relative_error(gradients[k], d_grads[k]) for k in names}
    worst = max(errors, key=errors.get)
    assert errors[worst] < 2e-3, (worst, errors[worst])
    assert relative_error(ggenerated, g_audio) < 2e-3
    return D


def test_double_agrees_with_the_oracle_on_the_gpu_verified_discriminators(monkeypatch):
    D = run(monkeypatch, init.discriminator_state(1234))
    assert [type(m).__name__ for m in D.modules] == ['Period'] * 5


def test_multi_resolution_sequencing_matches_autograd(monkeypatch):
    """DiscriminatorR (train/discriminator.py:Resolution): its launches, views and flags over
    the double (how it was developed before it first ran on a GPU)"""
    D = run(monkeypatch, init.discriminator_state(1234, multi_resolution=True), seed=1)
    assert [type(m).__name__ for m in D.modules[5:]] == ['Resolution'] * 3


@pytest.mark.parametrize('flags', [
    {}, {'multi_scale': True, 'multi_resolution': True, 'spectral_convergence': True}],
    ids=['config/promonet.py', 'every flag'])
def test_whole_training_step_sequencing_matches_autograd(monkeypatch, flags):
    """Trainer.step(update=False) over the double against oracle/train.py's step in fp64: the same
    comparison tests/test_train_gpu.py makes on the GPU, here for the host logic alone; once for
    the default configuration, once with MULTI_SCALE_DISCRIMINATOR, MULTI_RESOLUTION_DISCRIMINATOR
    and SPECTRAL_CONVERGENCE_LOSS on"""
    from promonet_b200.train import Trainer
    emulated_ops.install(monkeypatch)
    spectral = flags.get('spectral_convergence', False)
    states = init.hifigan_state(1234), init.discriminator_state(
        1234, flags.get('multi_scale', False), flags.get('multi_resolution', False))
    batch = oracle_train.batch(2 if not flags else 1, 8, seed=21)
    g_state = oracle_train.leaf_state(states[0], torch.float64)
    d_state = oracle_train.leaf_state(states[1], torch.float64)
    losses, g_grads, d_grads, generated = oracle_train.step(
        g_state, d_state, [t.double() if t.is_floating_point() else t for t in batch],
        spectral_convergence=spectral)
    trainer = Trainer(*states, device='cpu', math='fp32', spectral_convergence_loss=spectral)
    assert len(trainer.discriminators.modules) == 5 + (4 if flags else 0)
    ours = trainer.step(*[t.contiguous() for t in batch], update=False)
    assert relative_error(trainer.generated, generated) < 1e-4
    names = ('discriminator', 'mel', 'feature_matching', 'adversarial', 'generator') + (
        ('spectral_convergence',) if spectral else ())
    for i, name in enumerate(names):
        assert float(ours[i]) == pytest.approx(float(losses[name]), rel=1e-4), name
    for module, expected in ((trainer.discriminators, d_grads), (trainer.generator, g_grads)):
        gradients = module.params.gradients()
        assert sorted(gradients) == sorted(expected)
        errors = sorted((relative_error(gradients[n], g), n) for n, g in expected.items())
        # sign / mask decisions at near-zero values may move a few tensors (see test_train_gpu.py)
        assert errors[len(errors) // 2][0] < 4e-4 and errors[-1][0] < 5e-2, errors[-4:]
        assert sum(e >= 2e-3 for e, _ in errors) <= 4, errors[-6:]
