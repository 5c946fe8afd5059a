"""Random initialisation of a generator state dict

`promonet.model.Generator()` (promonet/model/generator.py:84-114) is what the
reference means by a random-init generator.  This builds the same state dict
(same keys, shapes, distributions) and draws from the torch RNG in the same
order as that constructor does, so that under the same seed the tensors are
bit-identical to the reference's -- which is what lets tests/golden pin the
CUDA path against outputs of the real reference without shipping 57 MB of
weights.  Order of draws: HiFiGAN (input convs, then per stage the transposed
conv and the 3 Blocks' convs1/convs2, each ModuleList followed by the
`init_weights` normal_ draws of hifigan.py:220-223, which land in `.weight`
and are discarded by weight norm but still advance the RNG), output conv,
speaker embedding, pitch embedding.
"""
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

from promonet_b200 import config

ASSETS = Path(__file__).resolve().parent.parent / 'assets'


def pitch_distribution():
    """Pitch bin edges (promonet.load.pitch_distribution, load.py:54-71;
    values of assets/stats/vctk-256-loudness-pitch-viterbi.pt)"""
    return torch.from_numpy(np.load(ASSETS / 'pitch_distribution.npy'))


def _conv(state, name, module):
    state[f'{name}.weight'] = module.weight.detach().clone()
    if module.bias is not None:
        state[f'{name}.bias'] = module.bias.detach().clone()


def _weight_norm_conv(state, name, module):
    """Keys in the order torch.nn.utils.weight_norm registers them"""
    v = module.weight.detach().clone()
    g = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.ndim - 1)))
    state[f'{name}.bias'] = module.bias.detach().clone()
    state[f'{name}.weight_g'] = g
    state[f'{name}.weight_v'] = v


def _discard_init_weights(module):
    # hifigan.py:220-223 draws normal_(0, 0.01) into `.weight`
    torch.empty_like(module.weight).normal_(0., 0.01)


def hifigan_state(seed=None):
    """State dict of a freshly constructed hifigan Generator"""
    if seed is not None:
        torch.manual_seed(seed)
    state = OrderedDict()
    state['default_previous_samples'] = torch.zeros(1, 1, 1)
    initial = config.HIFIGAN_UPSAMPLE_INITIAL_SIZE
    _conv(state, 'model.input_feature_conv',
          torch.nn.Conv1d(config.NUM_FEATURES, initial, 7, 1, padding=3))
    _conv(state, 'model.input_speaker_conv',
          torch.nn.Conv1d(config.GLOBAL_CHANNELS, initial, 1))
    channels = initial
    for i, (k, s) in enumerate(zip(
        config.HIFIGAN_UPSAMPLE_KERNEL_SIZES,
        config.HIFIGAN_UPSAMPLE_RATES
    )):
        stage = f'model.model.{i}.model'
        up = torch.nn.ConvTranspose1d(
            channels, channels // 2, k, s, padding=(k - s) // 2)
        channels //= 2
        blocks = []
        for j, kernel in enumerate(config.HIFIGAN_RESBLOCK_KERNEL_SIZES):
            for group in ('convs1', 'convs2'):
                convs = [
                    torch.nn.Conv1d(channels, channels, kernel)
                    for _ in config.HIFIGAN_RESBLOCK_DILATION_SIZES]
                for conv in convs:
                    _discard_init_weights(conv)
                blocks.append((f'{stage}.2.model.{j}.{group}', convs))
        # MultiReceptiveFieldFusion applies init_weights to the upsampler after
        # the ResidualBlock has been constructed (hifigan.py:95-112)
        _discard_init_weights(up)
        _weight_norm_conv(state, f'{stage}.1', up)
        for name, convs in blocks:
            for d, conv in enumerate(convs):
                _weight_norm_conv(state, f'{name}.{d}', conv)
    _conv(state, f'model.model.{len(config.HIFIGAN_UPSAMPLE_RATES) + 1}',
          torch.nn.Conv1d(channels, 1, 7, 1, 3, bias=False))
    state['speaker_embedding.weight'] = torch.nn.Embedding(
        config.NUM_SPEAKERS, config.SPEAKER_CHANNELS).weight.detach().clone()
    state['pitch_embedding.weight'] = torch.nn.Embedding(
        config.PITCH_BINS, config.PITCH_EMBEDDING_SIZE).weight.detach().clone()
    state['ppg_threshold'] = torch.tensor(
        config.SPARSE_PPG_THRESHOLD, dtype=torch.float)
    state['pitch_distribution'] = pitch_distribution()
    return state


###############################################################################
# FARGAN (config/fargan.py)
###############################################################################


def _discard_orthogonal(rows, cols):
    # fargan.py:418-424 init_weights: orthogonal_ lands in the `.weight` that
    # weight norm recomputes from (g, v), but it still draws rows * cols normals
    torch.empty(rows, cols).normal_(0, 1)


def _weight_norm_linear(state, name, linear):
    v = linear.weight.detach().clone()
    state[f'{name}.weight_g'] = v.norm(dim=1, keepdim=True)
    state[f'{name}.weight_v'] = v


def fargan_state(seed=None):
    """State dict of a freshly constructed fargan Generator
    (promonet/model/fargan.py:16-19,139-197,349-388; same RNG draw order)"""
    if seed is not None:
        torch.manual_seed(seed)
    hop = config.HOPSIZE
    state = OrderedDict()
    state['default_previous_samples'] = torch.zeros(1, 1, 2 * hop)
    channels = config.NUM_FEATURES + config.GLOBAL_CHANNELS
    for i, out in zip((0, 2, 4), (channels, channels, 2 * hop)):
        state[f'model.conditioning_network.{i}.weight'] = torch.nn.Linear(
            channels, out, bias=False).weight.detach().clone()
    net = 'model.subframe_network'
    # FramewiseConv: Linear(520 -> 256), then its GLU (constructed, then GLU.apply),
    # then FramewiseConv.apply over both
    fw = torch.nn.Linear(2 * (hop + 4), hop, bias=False)
    fw_gate = torch.nn.Linear(hop, hop, bias=False)
    _discard_orthogonal(hop, hop)
    _discard_orthogonal(hop, 2 * (hop + 4))
    _discard_orthogonal(hop, hop)
    _weight_norm_linear(state, f'{net}.framewise_convolution.model.0', fw)
    _weight_norm_linear(state, f'{net}.framewise_convolution.model.2.gate', fw_gate)
    for i in (1, 2, 3):
        cell = torch.nn.GRUCell(hop + hop // 2, hop, bias=False)
        state[f'{net}.gru{i}.weight_ih'] = cell.weight_ih.detach().clone()
        state[f'{net}.gru{i}.weight_hh'] = cell.weight_hh.detach().clone()
    gates = {}
    for name in ('gru1_glu', 'gru2_glu', 'gru3_glu', 'skip_glu'):
        gates[name] = torch.nn.Linear(hop, hop, bias=False)
        _discard_orthogonal(hop, hop)
    skip = torch.nn.Linear(4 * hop + hop // 2, hop, bias=False)
    output = torch.nn.Linear(hop, hop // 4, bias=False)
    # SubframeNetwork.apply(init_weights): children in registration order
    _discard_orthogonal(hop, 2 * (hop + 4))   # framewise Linear
    _discard_orthogonal(hop, hop)             # framewise GLU gate
    for name in gates:
        _discard_orthogonal(hop, hop)
    torch.nn.init.orthogonal_(skip.weight.data)
    torch.nn.init.orthogonal_(output.weight.data)
    for name, gate in gates.items():
        _weight_norm_linear(state, f'{net}.{name}.gate', gate)
    state[f'{net}.skip_dense.weight'] = skip.weight.detach().clone()
    state[f'{net}.output_layer.weight'] = output.weight.detach().clone()
    state['speaker_embedding.weight'] = torch.nn.Embedding(
        config.NUM_SPEAKERS, config.SPEAKER_CHANNELS).weight.detach().clone()
    state['pitch_embedding.weight'] = torch.nn.Embedding(
        config.PITCH_BINS, config.PITCH_EMBEDDING_SIZE).weight.detach().clone()
    state['ppg_threshold'] = torch.tensor(config.SPARSE_PPG_THRESHOLD, dtype=torch.float)
    state['pitch_distribution'] = pitch_distribution()
    return state


###############################################################################
# Discriminator (config/promonet.py: 5 x DiscriminatorP + DiscriminatorCMB)
###############################################################################


# DiscriminatorS, model/discriminator.py:218-225: (c_in, c_out, kernel, stride, groups, padding)
MULTI_SCALE_CONVS = (
    (1, 16, 15, 1, 1, 7), (16, 64, 41, 4, 4, 20), (64, 256, 41, 4, 16, 20),
    (256, 1024, 41, 4, 64, 20), (1024, 1024, 41, 4, 256, 20), (1024, 1024, 5, 1, 1, 2))


# DiscriminatorR, model/discriminator.py:23: (n_fft, hop_length, win_length)
MULTI_RESOLUTIONS = ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))


def discriminator_state(seed=None, multi_scale=False, multi_resolution=False):
    """State dict of a freshly constructed promonet.model.Discriminator()
    (promonet/model/discriminator.py:15-34,61-72,148-173): same keys and the same
    torch RNG draw order (Conv2d.reset_parameters per layer, in construction order).
    multi_scale = MULTI_SCALE_DISCRIMINATOR (config/defaults.py:180): DiscriminatorS is
    inserted after the period discriminators (:20-21); multi_resolution =
    MULTI_RESOLUTION_DISCRIMINATOR (:177): three DiscriminatorR after that (:22-25, :99-110)"""
    if seed is not None:
        torch.manual_seed(seed)
    state = OrderedDict()
    for index, _ in enumerate(config.DISCRIMINATOR_PERIODS):
        prefix = f'discriminators.{index}'
        channels = (1, 32, 128, 512, 1024, 1024)
        for layer in range(5):
            stride = (3, 1) if layer < 4 else 1
            _weight_norm_conv(state, f'{prefix}.convs.{layer}', torch.nn.Conv2d(
                channels[layer], channels[layer + 1], (5, 1), stride, (2, 0)))
        _weight_norm_conv(
            state, f'{prefix}.conv_post', torch.nn.Conv2d(1024, 1, (3, 1), 1, (1, 0)))
    index = len(config.DISCRIMINATOR_PERIODS)
    if multi_scale:
        prefix = f'discriminators.{index}'
        for layer, (c_in, c_out, kernel, stride, groups, padding) in enumerate(MULTI_SCALE_CONVS):
            _weight_norm_conv(state, f'{prefix}.convs.{layer}', torch.nn.Conv1d(
                c_in, c_out, kernel, stride, groups=groups, padding=padding))
        _weight_norm_conv(state, f'{prefix}.conv_post', torch.nn.Conv1d(1024, 1, 3, 1, padding=1))
        index += 1
    for _ in MULTI_RESOLUTIONS if multi_resolution else ():
        prefix = f'discriminators.{index}'
        for layer in range(5):
            kernel = (3, 9) if layer < 4 else (3, 3)
            stride = (1, 2) if 1 <= layer <= 3 else (1, 1)
            _weight_norm_conv(
                state, f'{prefix}.convs.{layer}',
                torch.nn.Conv2d(1 if layer == 0 else 32, 32, kernel, stride,
                                padding=(1, kernel[1] // 2)))
        _weight_norm_conv(
            state, f'{prefix}.conv_post', torch.nn.Conv2d(32, 1, (3, 3), padding=(1, 1)))
        index += 1
    prefix = f'discriminators.{index}'
    for band, _ in enumerate(config.CMB_BANDS):
        for layer in range(5):
            kernel = (3, 9) if layer < 4 else (3, 3)
            stride = (1, 2) if 1 <= layer <= 3 else (1, 1)
            _weight_norm_conv(
                state, f'{prefix}.band_convs.{band}.{layer}.0',
                torch.nn.Conv2d(1 if layer == 0 else 32, 32, kernel, stride,
                                padding=(1, kernel[1] // 2)))
    _weight_norm_conv(
        state, f'{prefix}.conv_post', torch.nn.Conv2d(32, 1, (3, 3), (1, 1), padding=(1, 1)))
    return state
