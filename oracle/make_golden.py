"""Freeze golden vectors from the UNMODIFIED reference (container only).

    python -m oracle.make_golden

Imports /root/reference through oracle.ref_shim, constructs the reference
modules under seed 1234 and stores small input/output pairs in tests/golden/.
The weights themselves are not stored (57 MB): promonet_b200.model.init draws
bit-identical tensors from the same seed, and the per-tensor checksums stored
here prove it at test time.
"""
from pathlib import Path

import numpy as np
import torch

from oracle import inputs, ref_shim

GOLDEN = Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def checksums(state):
    return {
        k: float(v.double().abs().sum()) for k, v in state.items()
        if v.is_floating_point()}


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    promonet = ref_shim.load()
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator().eval()
    state = generator.state_dict()

    # Generator.forward (generator.py:116-135), 8-band and 513-row loudness
    result = {}
    for rows, frames, batch in ((8, 24, 2), (513, 16, 1)):
        loud, pitch, per, ppg, spk, sbr, lr = inputs.synthesis(
            batch, frames, seed=promonet.RANDOM_SEED + rows, loudness_rows=rows)
        with torch.no_grad():
            features = generator.prepare_features(loud, pitch, per, ppg)
            audio = generator(
                loud, pitch, per, ppg, spk, sbr, lr,
                generator.default_previous_samples)
        tag = f'r{rows}'
        result.update({
            f'{tag}_loudness': loud, f'{tag}_pitch': pitch,
            f'{tag}_periodicity': per, f'{tag}_ppg': ppg,
            f'{tag}_speakers': spk, f'{tag}_sbr': sbr, f'{tag}_lr': lr,
            f'{tag}_features': features, f'{tag}_audio': audio})
    names = sorted(checksums(state))
    result['checksum_names'] = np.array(names)
    result['checksum_values'] = np.array([checksums(state)[n] for n in names])
    np.savez_compressed(
        GOLDEN / 'generator.npz',
        **{k: v.numpy() if torch.is_tensor(v) else v for k, v in result.items()})

    # Block.forward (hifigan.py:198-210) at C=32, k=3 and C=64, k=11
    blocks = {}
    for channels, kernel, length in ((32, 3, 300), (64, 11, 200)):
        torch.manual_seed(promonet.RANDOM_SEED + channels)
        block = promonet.model.hifigan.Block(channels, kernel, (1, 3, 5)).eval()
        x = torch.randn(2, channels, length)
        with torch.no_grad():
            y = block(x)
        tag = f'c{channels}k{kernel}'
        blocks[f'{tag}_x'] = x.numpy()
        blocks[f'{tag}_y'] = y.numpy()
        for key, value in block.state_dict().items():
            blocks[f'{tag}_{key}'] = value.numpy()
    np.savez_compressed(GOLDEN / 'block.npz', **blocks)

    # spectrogram.from_audio (preprocess/spectrogram.py:15-60,111-135)
    audio = inputs.audio(2, 4096, seed=promonet.RANDOM_SEED)[:, None]
    with torch.no_grad():
        linear = promonet.preprocess.spectrogram.from_audio(audio)
        mels = promonet.preprocess.spectrogram.from_audio(audio, mels=True)
    np.savez_compressed(
        GOLDEN / 'spectrogram.npz',
        audio=audio.numpy(), linear=linear.numpy(), mels=mels.numpy())

    # loudness.band_average / normalize (preprocess/loudness.py:84-146)
    loudness = torch.rand(513, 12) * 100. - 100.
    np.savez_compressed(
        GOLDEN / 'loudness_bands.npz',
        loudness=loudness.numpy(),
        averaged=promonet.preprocess.loudness.band_average(loudness, 8).numpy(),
        normalized=promonet.preprocess.loudness.normalize(loudness).numpy())
    # edit.grid.sample (promonet/edit/grid.py:12-43): time-stretch of a PPG-shaped sequence
    torch.manual_seed(99)
    sequence = torch.softmax(2. * torch.randn(2, 40, 57), dim=-2)
    grid = torch.linspace(0., 56., 91)
    np.savez_compressed(
        GOLDEN / 'grid.npz', sequence=sequence.numpy(), grid=grid.numpy(),
        linear=promonet.edit.grid.sample(sequence, grid, 'linear').numpy(),
        nearest=promonet.edit.grid.sample(sequence, grid, 'nearest').numpy())
    print('wrote', sorted(p.name for p in GOLDEN.iterdir()))


def fargan():
    """config/fargan.py needs its own process (derived constants freeze at import)"""
    GOLDEN.mkdir(parents=True, exist_ok=True)
    promonet = ref_shim.load(ref_shim.REFERENCE_ROOT + '/config/fargan.py')
    assert promonet.MODEL == 'fargan'
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator().eval()
    state = generator.state_dict()
    loud, pitch, per, ppg, spk, sbr, lr = inputs.synthesis(2, 12, seed=99)
    previous = 0.1 * torch.randn(2, 1, promonet.NUM_PREVIOUS_SAMPLES)
    with torch.no_grad():
        audio = generator(loud, pitch, per, ppg, spk, sbr, lr, torch.zeros_like(previous))
        audio_previous = generator(loud, pitch, per, ppg, spk, sbr, lr, previous)
    names = sorted(checksums(state))
    np.savez_compressed(
        GOLDEN / 'fargan.npz',
        loudness=loud.numpy(), pitch=pitch.numpy(), periodicity=per.numpy(), ppg=ppg.numpy(),
        speakers=spk.numpy(), sbr=sbr.numpy(), lr=lr.numpy(), previous=previous.numpy(),
        audio=audio.numpy(), audio_previous=audio_previous.numpy(),
        checksum_names=np.array(names),
        checksum_values=np.array([checksums(state)[n] for n in names]))
    print('wrote fargan.npz')


TRAIN_BATCH, TRAIN_FRAMES, TRAIN_SEED, TRAIN_STEPS = 2, 32, 77, 2
# small tensors whose full gradients are stored (the rest are pinned by their norms)
TRAIN_FULL_GRADIENTS = (
    'generator:model.model.5.weight',
    'generator:model.input_speaker_conv.bias',
    'generator:model.model.3.model.2.model.1.convs1.2.weight_g',
    'generator:model.model.0.model.1.weight_g',
    'discriminator:discriminators.0.conv_post.weight_v',
    'discriminator:discriminators.3.convs.0.weight_v',
    'discriminator:discriminators.5.band_convs.2.1.0.weight_g',
    'discriminator:discriminators.5.conv_post.bias')


def train():
    """Two iterations of the reference training step (promonet/train/core.py:183-369)
    composed from the UNMODIFIED reference modules: promonet.model.Generator,
    promonet.model.Discriminator, promonet.loss.*, promonet.preprocess.spectrogram.*,
    promonet.OPTIMIZER; fp32, no autocast (GradScaler is the identity on finite fp32)."""
    from oracle import train as oracle_train
    promonet = ref_shim.load()
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator()
    torch.manual_seed(promonet.RANDOM_SEED)
    discriminators = promonet.model.Discriminator()
    generator.train()
    discriminators.train()
    discriminator_optimizer = promonet.OPTIMIZER(discriminators.parameters())
    generator_optimizer = promonet.OPTIMIZER(generator.parameters())
    batch = oracle_train.batch(TRAIN_BATCH, TRAIN_FRAMES, TRAIN_SEED)
    (loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio) = batch
    result = {
        'speakers': speakers.numpy(), 'sbr': sbr.numpy(), 'lr': lr.numpy(),
        'pitch': pitch.numpy(), 'periodicity': periodicity.numpy(), 'audio': audio.numpy(),
        'input_checksums': np.array([float(t.double().abs().sum()) for t in batch])}
    previous_samples = torch.zeros(promonet.HOPSIZE)
    for step in range(TRAIN_STEPS):
        # train/core.py:223-256
        generated = generator(
            loudness, pitch, periodicity, ppg, speakers, sbr, lr, previous_samples)
        real_logits, fake_logits, _, _ = discriminators(audio, generated.detach())
        discriminator_losses, _, _ = promonet.loss.discriminator(
            [logit.float() for logit in real_logits], [logit.float() for logit in fake_logits])
        discriminator_optimizer.zero_grad()
        discriminator_losses.backward()
        d_grads = {k: v.grad.clone() for k, v in discriminators.named_parameters()}
        discriminator_optimizer.step()
        # :262-369
        _, fake_logits, real_feature_maps, fake_feature_maps = discriminators(audio, generated)
        mels = promonet.preprocess.spectrogram.linear_to_mel(spectrograms, None)
        generated_mels = promonet.preprocess.spectrogram.from_audio(generated, True, None)
        mel_loss = torch.nn.functional.l1_loss(mels, generated_mels)
        feature_matching_loss = promonet.loss.feature_matching(real_feature_maps, fake_feature_maps)
        adversarial_loss, _ = promonet.loss.generator([logit.float() for logit in fake_logits])
        generator_losses = (
            promonet.MEL_LOSS_WEIGHT * mel_loss +
            promonet.FEATURE_MATCHING_LOSS_WEIGHT * feature_matching_loss +
            promonet.ADVERSARIAL_LOSS_WEIGHT * adversarial_loss)
        generator_optimizer.zero_grad()
        generator_losses.backward()
        g_grads = {k: v.grad.clone() for k, v in generator.named_parameters()}
        generator_optimizer.step()
        result[f'losses_{step}'] = np.array([
            float(discriminator_losses), float(mel_loss), float(feature_matching_loss),
            float(adversarial_loss), float(generator_losses)])
        result[f'generated_{step}'] = generated.detach().numpy()
        for kind, grads in (('generator', g_grads), ('discriminator', d_grads)):
            names = sorted(grads)
            result[f'{kind}_names'] = np.array(names)
            result[f'{kind}_grad_norms_{step}'] = np.array(
                [float(grads[n].double().norm()) for n in names])
            if step == 0:
                for full in TRAIN_FULL_GRADIENTS:
                    owner, name = full.split(':')
                    if owner == kind:
                        result[f'grad:{full}'] = grads[name].numpy()
    for kind, module in (('generator', generator), ('discriminator', discriminators)):
        params = dict(module.named_parameters())
        result[f'{kind}_param_checksums'] = np.array(
            [float(params[n].detach().double().abs().sum()) for n in sorted(params)])
    np.savez_compressed(GOLDEN / 'train.npz', **result)
    print('wrote train.npz', {k: result[k] for k in ('losses_0', 'losses_1')})


def train_flags():
    """One iteration with MULTI_SCALE_DISCRIMINATOR and SPECTRAL_CONVERGENCE_LOSS on (the
    configuration BASELINE.json words for the training step), from the unmodified reference"""
    from oracle import train as oracle_train
    promonet = ref_shim.load()
    promonet.MULTI_SCALE_DISCRIMINATOR = True
    promonet.SPECTRAL_CONVERGENCE_LOSS = True
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator()
    torch.manual_seed(promonet.RANDOM_SEED)
    discriminators = promonet.model.Discriminator()
    assert len(discriminators.discriminators) == 7
    spectral_convergence = promonet.loss.MultiResolutionSpectralConvergence('cpu')
    batch = oracle_train.batch(TRAIN_BATCH, TRAIN_FRAMES, TRAIN_SEED)
    (loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio) = batch
    generated = generator(
        loudness, pitch, periodicity, ppg, speakers, sbr, lr, torch.zeros(promonet.HOPSIZE))
    real_logits, fake_logits, _, _ = discriminators(audio, generated.detach())
    discriminator_losses, _, _ = promonet.loss.discriminator(real_logits, fake_logits)
    discriminator_losses.backward()
    d_grads = {k: v.grad.clone() for k, v in discriminators.named_parameters()}
    _, fake_logits, real_feature_maps, fake_feature_maps = discriminators(audio, generated)
    mels = promonet.preprocess.spectrogram.linear_to_mel(spectrograms, None)
    generated_mels = promonet.preprocess.spectrogram.from_audio(generated, True, None)
    mel_loss = torch.nn.functional.l1_loss(mels, generated_mels)
    spectral_loss = spectral_convergence(generated, audio)
    feature_matching_loss = promonet.loss.feature_matching(real_feature_maps, fake_feature_maps)
    adversarial_loss, _ = promonet.loss.generator(fake_logits)
    generator_losses = (
        promonet.MEL_LOSS_WEIGHT * mel_loss + spectral_loss +
        promonet.FEATURE_MATCHING_LOSS_WEIGHT * feature_matching_loss +
        promonet.ADVERSARIAL_LOSS_WEIGHT * adversarial_loss)
    generator_losses.backward()
    g_grads = {k: v.grad.clone() for k, v in generator.named_parameters()}
    result = {
        'losses': np.array([
            float(discriminator_losses), float(mel_loss), float(feature_matching_loss),
            float(adversarial_loss), float(generator_losses), float(spectral_loss)]),
        'logit_sizes': np.array([logit.shape[1] for logit in fake_logits]),
        'feature_maps': np.array([len(maps) for maps in fake_feature_maps])}
    for kind, grads in (('generator', g_grads), ('discriminator', d_grads)):
        names = sorted(grads)
        result[f'{kind}_names'] = np.array(names)
        result[f'{kind}_grad_norms'] = np.array([float(grads[n].double().norm()) for n in names])
    state = discriminators.state_dict()
    result['msd_checksums'] = np.array([
        float(state[k].double().abs().sum()) for k in sorted(state) if k.startswith('discriminators.5.')])
    np.savez_compressed(GOLDEN / 'train_flags.npz', **result)
    promonet.MULTI_SCALE_DISCRIMINATOR = False
    promonet.SPECTRAL_CONVERGENCE_LOSS = False
    print('wrote train_flags.npz', result['losses'], result['logit_sizes'])


def train_resolution():
    """MULTI_RESOLUTION_DISCRIMINATOR (config/defaults.py:177) on: the three DiscriminatorR of the
    unmodified reference (model/discriminator.py:96-141) — logits, feature-map checksums, the
    gradient norms of the discriminator loss and of the generator-side losses"""
    promonet = ref_shim.load()
    promonet.MULTI_RESOLUTION_DISCRIMINATOR = True
    torch.manual_seed(promonet.RANDOM_SEED)
    discriminators = promonet.model.Discriminator()
    promonet.MULTI_RESOLUTION_DISCRIMINATOR = False
    assert len(discriminators.discriminators) == 9
    generator = torch.Generator().manual_seed(TRAIN_SEED)
    audio = .3 * torch.randn(2, 1, 8192, generator=generator)
    generated = (.3 * torch.randn(2, 1, 8192, generator=generator)).requires_grad_()
    real_logits, fake_logits, real_maps, fake_maps = discriminators(audio, generated)
    discriminator_loss, _, _ = promonet.loss.discriminator(
        [logit.float() for logit in real_logits], [logit.float() for logit in fake_logits])
    names = sorted(k for k, _ in discriminators.named_parameters() if k.split('.')[1] in '567')
    parameters = dict(discriminators.named_parameters())
    d_grads = torch.autograd.grad(
        discriminator_loss, [parameters[k] for k in names], retain_graph=True)
    generator_loss = promonet.loss.feature_matching(real_maps, fake_maps) + \
        promonet.loss.generator(fake_logits)[0]
    g_grad, = torch.autograd.grad(generator_loss, generated)
    state = discriminators.state_dict()
    result = {
        'audio': audio.numpy(), 'generated': generated.detach().numpy(),
        'losses': np.array([float(discriminator_loss), float(generator_loss)]),
        'names': np.array(names),
        'grad_norms': np.array([float(g.double().norm()) for g in d_grads]),
        'generated_grad': g_grad.numpy(),
        'checksums': np.array([float(state[k].double().abs().sum()) for k in sorted(state)])}
    for index in (5, 6, 7):
        result[f'logits_real_{index}'] = real_logits[index].detach().numpy()
        result[f'logits_fake_{index}'] = fake_logits[index].detach().numpy()
        result[f'fmap_checksums_{index}'] = np.array(
            [float(m.double().abs().sum()) for m in fake_maps[index]])
    np.savez_compressed(GOLDEN / 'train_resolution.npz', **result)
    print('wrote train_resolution.npz', result['losses'], [real_logits[i].shape for i in (5, 6, 7)])


def metric_inputs(seed, frames, rows):
    """Seeded (loudness, pitch, periodicity, ppg) straddling the loudness and voicing thresholds"""
    generator = torch.Generator().manual_seed(seed)
    rand = lambda *shape: torch.rand(*shape, generator=generator)
    return (
        rand(rows, frames) * 70. - 100.,                        # dB around the -60 dB threshold
        50. * 11. ** rand(1, frames),                           # 50-550 Hz
        rand(1, frames) * .4,                                   # around VOICING_THRESHOLD .1625
        torch.softmax(2. * torch.randn(1, 40, frames, generator=generator), 1))


METRIC_CASES = ((37, 8, 8), (200, 8, 513), (1, 513, 8), (64, 8, 8))   # frames, rows, rows


def metrics():
    """promonet.evaluate.Metrics (evaluate/metrics.py:17-312) and promonet.edit.from_features
    (edit/core.py:17-132) of the UNMODIFIED reference, on top of the restated third-party
    primitives the shim provides (torchutil.metrics, penn.voicing.threshold, ppgs.distance,
    ppgs.edit.grid)"""
    GOLDEN.mkdir(parents=True, exist_ok=True)
    promonet = ref_shim.load()
    result = {'cases': np.array(METRIC_CASES)}
    accumulated = promonet.evaluate.Metrics()
    for index, (frames, rows, target_rows) in enumerate(METRIC_CASES):
        predicted = metric_inputs(100 + index, frames, rows)
        target = metric_inputs(200 + index, frames, target_rows)
        single = promonet.evaluate.Metrics()
        single.update(*predicted, *target)
        if index < 3:
            accumulated.update(*predicted, *target)
        names = sorted(single())
        result['names'] = np.array(names)
        result[f'single_{index}'] = np.array([single()[name] for name in names])
    result['accumulated'] = np.array([accumulated()[name] for name in names])
    # edits: the two evaluation ratios (config/defaults.py:204) and a pitch shift + loudness scale
    loudness, pitch, periodicity, ppg = metric_inputs(300, 57, 8)
    ppg = ppg[0]
    edits = (
        {'time_stretch_ratio': .717}, {'time_stretch_ratio': 1.414},
        {'pitch_shift_cents': 600., 'loudness_scale_db': 5.},
        {'pitch_shift_cents': -900., 'time_stretch_ratio': 2.})
    for index, kwargs in enumerate(edits):
        outputs = promonet.edit.from_features(
            loudness.clone(), pitch.clone(), periodicity.clone(), ppg.clone(), **kwargs)
        for name, value in zip(('loudness', 'pitch', 'periodicity', 'ppg'), outputs):
            result[f'edit_{index}_{name}'] = value.numpy()
    result['edit_arguments'] = np.array([
        [kwargs.get(key, np.nan) for key in
         ('pitch_shift_cents', 'time_stretch_ratio', 'loudness_scale_db')] for kwargs in edits])
    np.savez_compressed(GOLDEN / 'metrics.npz', **result)
    print('wrote metrics.npz', dict(zip(names, result['accumulated'])))


if __name__ == '__main__':
    import sys
    if '--metrics' in sys.argv:
        metrics()
    elif '--train-resolution' in sys.argv:
        train_resolution()
    elif '--train-flags' in sys.argv:
        train_flags()
    elif '--fargan' in sys.argv:
        fargan()
    elif '--train' in sys.argv:
        train()
    else:
        main()
