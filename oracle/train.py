"""Oracle: discriminators, losses and one GAN training step (test infrastructure,
see oracle/__init__).

Plain-torch functional restatement, differentiated by torch.autograd, of
  promonet/model/discriminator.py:13-208   Discriminator = 5 x DiscriminatorP + DiscriminatorCMB
  promonet/train/loss.py:11-53             feature matching, LSGAN discriminator / generator
  promonet/train/core.py:183-369           discriminator step then generator step
over reference state dicts (weight_g / weight_v keys).  Pinned against the
unmodified reference modules by oracle/make_golden.py --train (tests/golden/train.npz)
and, in the build container, live in tests/test_oracle.py.  fp32, no autocast
(SURVEY 8a note 2); the reference's GradScaler is a no-op on finite fp32 gradients.
"""
import torch
import torch.nn.functional as F

from oracle import dsp, hifigan

LRELU_SLOPE = 0.1                 # promonet/config/defaults.py:216
PERIODS = (2, 3, 5, 7, 11)        # discriminator.py:19
CMB_BANDS = ((0, 51), (51, 128), (128, 256), (256, 384), (384, 513))  # :150,161-163
MEL_LOSS_WEIGHT = 45.             # defaults.py:340
FEATURE_MATCHING_LOSS_WEIGHT = 1.  # :328
ADVERSARIAL_LOSS_WEIGHT = 1.      # :322
LEARNING_RATE, BETAS, EPS = 2e-4, (.8, .99), 1e-9   # :390-394


def weight(state, prefix):
    return hifigan.weight(state, prefix)


def discriminator_p(state, prefix, x, period):
    """DiscriminatorP.forward discriminator.py:74-93"""
    fmaps = []
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), 'reflect')
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for i in range(5):
        stride = (3, 1) if i < 4 else 1
        x = F.conv2d(
            x, weight(state, f'{prefix}.convs.{i}'), state[f'{prefix}.convs.{i}.bias'],
            stride, (2, 0))
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmaps.append(x)
    x = F.conv2d(
        x, weight(state, f'{prefix}.conv_post'), state[f'{prefix}.conv_post.bias'], 1, (1, 0))
    fmaps.append(x)
    return torch.flatten(x, 1, -1), fmaps


def cmb_spectrogram(x):
    """DiscriminatorCMB.spectrogram discriminator.py:175-195 (no window: rectangular)"""
    x = F.pad(x, (384, 384), mode='reflect')
    x = torch.stft(
        x.squeeze(1), n_fft=1024, hop_length=256, win_length=1024,
        window=torch.ones(1024, dtype=x.dtype), center=False, return_complex=True)
    x = torch.norm(torch.view_as_real(x), p=2, dim=-1).unsqueeze(1)
    x = torch.permute(x, (0, 1, 3, 2))
    return [x[..., lo:hi] for lo, hi in CMB_BANDS]


def discriminator_cmb(state, prefix, x):
    """DiscriminatorCMB.forward discriminator.py:197-208"""
    outputs, fmaps = [], []
    for b, band in enumerate(cmb_spectrogram(x)):
        for i in range(5):
            name = f'{prefix}.band_convs.{b}.{i}.0'
            kernel_w = 9 if i < 4 else 3
            stride = (1, 2) if 1 <= i <= 3 else (1, 1)
            band = F.conv2d(
                band, weight(state, name), state[f'{name}.bias'], stride, (1, kernel_w // 2))
            band = F.leaky_relu(band, 0.1)
            fmaps.append(band)
        outputs.append(band)
    x = torch.cat(outputs, dim=-1)
    x = F.conv2d(
        x, weight(state, f'{prefix}.conv_post'), state[f'{prefix}.conv_post.bias'], 1, (1, 1))
    fmaps.append(x)
    return torch.flatten(x, 1, -1), fmaps


# DiscriminatorS discriminator.py:218-225: (kernel, stride, groups, padding)
MULTI_SCALE_CONVS = ((15, 1, 1, 7), (41, 4, 4, 20), (41, 4, 16, 20), (41, 4, 64, 20), (41, 4, 256, 20), (5, 1, 1, 2))


def discriminator_s(state, prefix, x):
    """DiscriminatorS.forward discriminator.py:227-239"""
    fmaps = []
    for i, (_, stride, groups, padding) in enumerate(MULTI_SCALE_CONVS):
        x = F.conv1d(
            x, weight(state, f'{prefix}.convs.{i}'), state[f'{prefix}.convs.{i}.bias'],
            stride, padding, 1, groups)
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmaps.append(x)
    x = F.conv1d(x, weight(state, f'{prefix}.conv_post'), state[f'{prefix}.conv_post.bias'], 1, 1)
    fmaps.append(x)
    return torch.flatten(x, 1, -1), fmaps


MULTI_RESOLUTIONS = ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))   # discriminator.py:23


def resolution_spectrogram(x, resolution):
    """DiscriminatorR.spectrogram discriminator.py:127-141: reflect pad (n_fft - hop) / 2, STFT
    without a window (rectangular over win_length, centred in n_fft by torch.stft), magnitude"""
    n_fft, hop_length, win_length = resolution
    pad = int((n_fft - hop_length) / 2)
    x = F.pad(x, (pad, pad), mode='reflect')
    x = torch.stft(
        x.squeeze(1), n_fft=n_fft, hop_length=hop_length, win_length=win_length,
        window=torch.ones(win_length, dtype=x.dtype), center=False, return_complex=True)
    return torch.norm(torch.view_as_real(x), p=2, dim=-1).unsqueeze(1)


def discriminator_r(state, prefix, x, resolution):
    """DiscriminatorR.forward discriminator.py:112-125 (MULTI_RESOLUTION_DISCRIMINATOR): note the
    LeakyReLU slope 0.2 here (:121), not LRELU_SLOPE"""
    x = resolution_spectrogram(x, resolution)
    fmaps = []
    for i in range(5):
        kernel_w = 9 if i < 4 else 3
        stride = (1, 2) if 1 <= i <= 3 else (1, 1)
        x = F.conv2d(
            x, weight(state, f'{prefix}.convs.{i}'), state[f'{prefix}.convs.{i}.bias'], stride,
            (1, kernel_w // 2))
        x = F.leaky_relu(x, 0.2)
        fmaps.append(x)
    x = F.conv2d(
        x, weight(state, f'{prefix}.conv_post'), state[f'{prefix}.conv_post.bias'], 1, (1, 1))
    fmaps.append(x)
    return torch.flatten(x, 1, -1), fmaps


def kinds(state):
    """The sub-discriminators of a state dict in order (discriminator.py:15-34): 'p' x 5, then 's'
    if MULTI_SCALE_DISCRIMINATOR, 'r' x 3 if MULTI_RESOLUTION_DISCRIMINATOR, then 'cmb'"""
    result = []
    count = len({k.split('.')[1] for k in state if k.startswith('discriminators.')})
    for i in range(count):
        prefix = f'discriminators.{i}'
        if f'{prefix}.band_convs.0.0.0.weight_v' in state:
            result.append('cmb')
        elif state[f'{prefix}.convs.0.weight_v'].ndim == 3:
            result.append('s')
        elif state[f'{prefix}.convs.0.weight_v'].shape[-1] == 9:
            result.append('r')
        else:
            result.append('p')
    return result


def discriminator(state, y, y_hat):
    """Discriminator.forward discriminator.py:36-49 (config/promonet.py: MPD x 5 + CMB; with
    MULTI_SCALE_DISCRIMINATOR the DiscriminatorS and with MULTI_RESOLUTION_DISCRIMINATOR the three
    DiscriminatorR sit between them, :18-28)"""
    logits_real, logits_fake, fmaps_real, fmaps_fake = [], [], [], []
    periods, resolutions = iter(PERIODS), iter(MULTI_RESOLUTIONS)
    for i, kind in enumerate(kinds(state)):
        prefix = f'discriminators.{i}'
        argument = next(periods) if kind == 'p' else next(resolutions) if kind == 'r' else None
        for x, logits, fmaps in ((y, logits_real, fmaps_real), (y_hat, logits_fake, fmaps_fake)):
            if kind == 'p':
                logit, fmap = discriminator_p(state, prefix, x, argument)
            elif kind == 's':
                logit, fmap = discriminator_s(state, prefix, x)
            elif kind == 'r':
                logit, fmap = discriminator_r(state, prefix, x, argument)
            else:
                logit, fmap = discriminator_cmb(state, prefix, x)
            logits.append(logit)
            fmaps.append(fmap)
    return logits_real, logits_fake, fmaps_real, fmaps_fake


def feature_matching_loss(real_fmaps, fake_fmaps):
    """loss.py:11-26"""
    loss = 0.
    for real_fmap, fake_fmap in zip(real_fmaps, fake_fmaps):
        for real, fake in zip(real_fmap, fake_fmap):
            loss = loss + torch.mean(torch.abs(real.float().detach() - fake.float()))
    return loss


def discriminator_loss(real_outputs, fake_outputs):
    """loss.py:29-40 (LSGAN)"""
    real = [torch.mean((1. - r) ** 2.) for r in real_outputs]
    fake = [torch.mean(f ** 2.) for f in fake_outputs]
    return sum(real) + sum(fake)


def generator_loss(outputs):
    """loss.py:43-53 (LSGAN)"""
    return sum(torch.mean((1. - o) ** 2.) for o in outputs)


def mel_loss(spectrograms, generated):
    """train/core.py:277-305 with SPARSE_MEL_LOSS False (no clamp)"""
    basis = torch.from_numpy(dsp.mel_basis()).to(generated.dtype)
    target = torch.log(basis @ spectrograms)
    predicted = torch.log(basis @ dsp.magnitude(generated))
    return F.l1_loss(target, predicted)


SPECTRAL_FFT_SIZES = (2560, 1280, 640, 320, 160, 80)   # loss.py:129-131


def spectral_convergence_loss(x, y):
    """MultiResolutionSpectralConvergence.forward loss.py:124-150: x predicted, y target (B, 1, T)"""
    total = 0.
    for n_fft in SPECTRAL_FFT_SIZES:
        window = torch.hann_window(n_fft, dtype=x.dtype)
        magnitudes = []
        for signal in (x, y):
            magnitude = torch.abs(torch.stft(
                signal.squeeze(1), n_fft, n_fft // 4, n_fft, window, return_complex=True))
            magnitudes.append(torch.sqrt(torch.clamp(magnitude, min=1e-7)))
        x_mag, y_mag = magnitudes
        total = total + torch.norm(y_mag - x_mag, p=1) / torch.norm(y_mag, p=1)
    return total / len(SPECTRAL_FFT_SIZES)


def parameters(state):
    """The entries torch registers as parameters (everything but the buffers)"""
    buffers = ('default_previous_samples', 'ppg_threshold', 'pitch_distribution')
    return {k: v for k, v in state.items() if k not in buffers}


def step(generator_state, discriminator_state, batch, optimizers=None, spectral_convergence=False):
    """One iteration of train/core.py:183-369.  States are dicts of leaf tensors
    (requires_grad on the parameters); returns losses, gradients and the audio.
    When `optimizers` = (discriminator AdamW, generator AdamW) is given they are stepped
    in the reference's order (the generator step sees the updated discriminator)."""
    (loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio) = batch
    g_params, d_params = parameters(generator_state), parameters(discriminator_state)
    for p in list(g_params.values()) + list(d_params.values()):
        p.grad = None

    # :223
    generated = hifigan.generator(
        generator_state, loudness, pitch, periodicity, ppg, speakers, sbr, lr)

    # :239-256 discriminator step
    real_logits, fake_logits, _, _ = discriminator(
        discriminator_state, audio, generated.detach())
    d_loss = discriminator_loss(real_logits, fake_logits)
    d_loss.backward()
    d_grads = {k: v.grad.clone() for k, v in d_params.items()}
    if optimizers is not None:
        optimizers[0].step()

    # :262-338 generator step
    _, fake_logits, real_fmaps, fake_fmaps = discriminator(
        discriminator_state, audio, generated)
    mel = mel_loss(spectrograms, generated)
    fm = feature_matching_loss(real_fmaps, fake_fmaps)
    adv = generator_loss(fake_logits)
    g_loss = MEL_LOSS_WEIGHT * mel + FEATURE_MATCHING_LOSS_WEIGHT * fm + ADVERSARIAL_LOSS_WEIGHT * adv
    if spectral_convergence:
        # train/core.py:308-310 (SPECTRAL_CONVERGENCE_LOSS)
        spectral = spectral_convergence_loss(generated, audio)
        g_loss = g_loss + spectral
    for p in g_params.values():
        p.grad = None
    g_loss.backward()
    g_grads = {k: v.grad.clone() for k, v in g_params.items()}
    if optimizers is not None:
        optimizers[1].step()
    losses = {
        'discriminator': d_loss.detach(), 'mel': mel.detach(), 'feature_matching': fm.detach(),
        'adversarial': adv.detach(), 'generator': g_loss.detach()}
    if spectral_convergence:
        losses['spectral_convergence'] = spectral.detach()
    return losses, g_grads, d_grads, generated.detach()


def make_optimizers(generator_state, discriminator_state):
    """promonet.OPTIMIZER, config/defaults.py:390-394; train/core.py:63-64"""
    make = lambda params: torch.optim.AdamW(params, lr=LEARNING_RATE, betas=BETAS, eps=EPS)
    return (make(list(parameters(discriminator_state).values())),
            make(list(parameters(generator_state).values())))


def leaf_state(state, dtype=torch.float32):
    """Detached copies with requires_grad on the parameters"""
    out = {}
    for k, v in state.items():
        v = v.detach().clone()
        if v.is_floating_point():
            v = v.to(dtype)
        out[k] = v
    for v in parameters(out).values():
        v.requires_grad_(True)
    return out


def batch(batch_size, frames, seed=1234):
    """Synthetic training batch with the shapes of data/collate.py:43-60"""
    from promonet_b200 import synthetic
    loudness, pitch, periodicity, ppg, speakers, sbr, lr, audio = synthetic.training(
        batch_size, frames, seed)
    with torch.no_grad():
        spectrograms = dsp.magnitude(audio)
    return loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio
