"""Oracle: HiFi-GAN generator forward (test infrastructure, see oracle/__init__).

Functional restatement of promonet/model/hifigan.py over a reference
state_dict (keys as produced by promonet.model.Generator().state_dict(), i.e.
`weight_g`/`weight_v` pairs for the weight-normed convs).  Runs in the dtype of
the state dict, so `to_double(state)` gives the fp64 yardstick.
"""
import torch

from oracle import features

LRELU_SLOPE = 0.1                       # promonet/config/defaults.py:216
RESBLOCK_KERNEL_SIZES = (3, 7, 11)      # :250
RESBLOCK_DILATIONS = (1, 3, 5)          # :253
UPSAMPLE_KERNEL_SIZES = (16, 16, 4, 4)  # :259
UPSAMPLE_RATES = (8, 8, 2, 2)           # :262


def fold_weight_norm(g, v):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v|| over dims != 0.

    Conv1d: dim 0 = C_out; ConvTranspose1d: dim 0 = C_in (hifigan.py:100-106).
    """
    norm = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.ndim - 1)))
    return g * v / norm


def weight(state, prefix):
    if f'{prefix}.weight' in state:
        return state[f'{prefix}.weight']
    return fold_weight_norm(state[f'{prefix}.weight_g'], state[f'{prefix}.weight_v'])


def lrelu(x):
    return torch.nn.functional.leaky_relu(x, LRELU_SLOPE)


def block(state, prefix, x, kernel_size):
    """hifigan.py:198-210"""
    for i, dilation in enumerate(RESBLOCK_DILATIONS):
        xt = torch.nn.functional.conv1d(
            lrelu(x),
            weight(state, f'{prefix}.convs1.{i}'),
            state[f'{prefix}.convs1.{i}.bias'],
            padding=dilation * (kernel_size - 1) // 2,
            dilation=dilation)
        xt = torch.nn.functional.conv1d(
            lrelu(xt),
            weight(state, f'{prefix}.convs2.{i}'),
            state[f'{prefix}.convs2.{i}.bias'],
            padding=(kernel_size - 1) // 2)
        x = xt + x
    return x


def vocoder(state, x, g, prefix='model.'):
    """HiFiGAN.forward hifigan.py:63-70 (+ Sequential :36-61)"""
    x = torch.nn.functional.conv1d(
        x,
        state[f'{prefix}input_feature_conv.weight'],
        state[f'{prefix}input_feature_conv.bias'],
        padding=3)
    x = x + torch.nn.functional.conv1d(
        g,
        state[f'{prefix}input_speaker_conv.weight'],
        state[f'{prefix}input_speaker_conv.bias'])
    for i, (k, s) in enumerate(zip(UPSAMPLE_KERNEL_SIZES, UPSAMPLE_RATES)):
        stage = f'{prefix}model.{i}.model'
        x = torch.nn.functional.conv_transpose1d(
            lrelu(x),
            weight(state, f'{stage}.1'),
            state[f'{stage}.1.bias'],
            stride=s,
            padding=(k - s) // 2)
        xs = None
        for j, kernel_size in enumerate(RESBLOCK_KERNEL_SIZES):
            y = block(state, f'{stage}.2.model.{j}', x, kernel_size)
            xs = y if xs is None else xs + y
        x = xs / len(RESBLOCK_KERNEL_SIZES)
    index = len(UPSAMPLE_RATES) + 1
    x = torch.nn.functional.conv1d(
        lrelu(x), state[f'{prefix}model.{index}.weight'], None, padding=3)
    return torch.tanh(x)


def generator(
    state,
    loudness,
    pitch,
    periodicity,
    ppg,
    speakers,
    spectral_balance_ratios,
    loudness_ratios
):
    """Generator.forward promonet/model/generator.py:116-135 (MODEL='hifigan')"""
    x = features.prepare_features(state, loudness, pitch, periodicity, ppg)
    g = features.prepare_global_features(
        state, speakers, spectral_balance_ratios, loudness_ratios)
    return vocoder(state, x, g)


def to_double(state):
    return {
        k: v.double() if v.is_floating_point() else v for k, v in state.items()}
