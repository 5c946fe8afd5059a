"""Oracle: penn-style pitch and periodicity (test infrastructure).

promonet/preprocess/core.py:64-85 calls `penn.from_audio(audio, sample_rate,
hopsize=256/22050 s, fmin=50, fmax=550, batch_size=2048, center='half-hop',
decoder='viterbi', interp_unvoiced_at=None, gpu)`.  penn is a third-party,
un-vendored dependency (dev branch, unpinned: setup.py:22, README.md:68-70) and
its pretrained fcnf0++ weights are not available offline, so this restates the
published FCNF0++ pipeline with random-init weights:
PARITY UNPINNED.  Every choice the upstream source would settle is written down
here once and the CUDA path is held to *this* definition:

* resample 22 050 -> 8 000 Hz with torchaudio's sinc_interp_hann kernel
  (lowpass_filter_width 6, rolloff 0.99);
* hop = int(hopsize_seconds * 8000) = 92 samples; reflect-pad (1024 - hop) // 2
  on both sides; frame i = padded[92 i : 92 i + 1024]; number of frames =
  max(1, int(T / (hopsize_seconds * sample_rate))) counted at the input rate;
* network: x[:, :, 16:-15] -> 6 x [Conv1d k32 -> ReLU -> (MaxPool 2) ->
  LayerNorm over (C, L)] -> Conv1d(512 -> 1440, k4): logits (frames, 1440);
* bins outside [floor(bin(fmin)), ceil(bin(fmax))) are masked to -inf;
  periodicity = 1 + sum p log(p + 1e-7) / ln 1440 with p = softmax(logits);
* Viterbi over p with a triangular-band transition (rows normalised) and a
  uniform initial distribution (both explicit inputs, torbi semantics);
* pitch = local expected value: softmax of the logits in a 19-bin window
  centred on the decoded bin, expected cents (5 cents per bin), Hz = 31 * 2^(c/1200).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from oracle import viterbi

SAMPLE_RATE = 8000
WINDOW_SIZE = 1024
PITCH_BINS = 1440
CENTS_PER_BIN = 5.
FMIN = 31.
OCTAVE = 1200.
MAX_OCTAVES_PER_SECOND = 32.
LOCAL_WINDOW = 19
LAYERS = (  # (c_in, c_out, pooled, length after the block)
    (1, 256, True, 481), (256, 32, True, 225), (32, 32, True, 97),
    (32, 128, False, 66), (128, 256, False, 35), (256, 512, False, 4))


def init_state(seed=1234):
    """Random FCNF0++ parameters (torch default Conv1d / LayerNorm init)"""
    generator_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    state = OrderedDict()
    for i, (c_in, c_out, _, length) in enumerate(LAYERS):
        conv = torch.nn.Conv1d(c_in, c_out, 32)
        state[f'layers.{i}.conv.weight'] = conv.weight.detach().clone()
        state[f'layers.{i}.conv.bias'] = conv.bias.detach().clone()
        # LayerNorm starts at weight 1, bias 0; perturb so that the affine is exercised
        state[f'layers.{i}.norm.weight'] = 1. + 0.1 * torch.randn(c_out, length)
        state[f'layers.{i}.norm.bias'] = 0.1 * torch.randn(c_out, length)
    conv = torch.nn.Conv1d(512, PITCH_BINS, 4)
    state['layers.6.weight'] = conv.weight.detach().clone()
    state['layers.6.bias'] = conv.bias.detach().clone()
    torch.random.set_rng_state(generator_state)
    return state


def resample(audio, sample_rate):
    import torchaudio
    if sample_rate == SAMPLE_RATE:
        return audio
    return torchaudio.functional.resample(audio, sample_rate, SAMPLE_RATE)


def hop_samples(hopsize_seconds):
    return int(hopsize_seconds * SAMPLE_RATE)


def expected_frames(samples, sample_rate, hopsize_seconds):
    return max(1, int(samples / (hopsize_seconds * sample_rate)))


def frames(audio, sample_rate, hopsize_seconds):
    """(1, T) at sample_rate -> (F, 1, 1024) at 8 kHz"""
    total = expected_frames(audio.shape[-1], sample_rate, hopsize_seconds)
    audio = resample(audio, sample_rate)
    hop = hop_samples(hopsize_seconds)
    padding = (WINDOW_SIZE - hop) // 2
    padded = torch.nn.functional.pad(audio[None], (padding, padding), mode='reflect')[0]
    needed = (total - 1) * hop + WINDOW_SIZE
    if padded.shape[-1] < needed:
        padded = torch.nn.functional.pad(padded, (0, needed - padded.shape[-1]))
    return padded.unfold(-1, WINDOW_SIZE, hop)[0, :total, None]


def infer(state, x):
    """(F, 1, 1024) -> logits (F, 1440)"""
    x = x[:, :, 16:-15]
    for i, (_, c_out, pooled, length) in enumerate(LAYERS):
        x = torch.nn.functional.conv1d(
            x, state[f'layers.{i}.conv.weight'], state[f'layers.{i}.conv.bias'])
        x = torch.relu(x)
        if pooled:
            x = torch.nn.functional.max_pool1d(x, 2, 2)
        x = torch.nn.functional.layer_norm(
            x, (c_out, length), state[f'layers.{i}.norm.weight'],
            state[f'layers.{i}.norm.bias'])
    x = torch.nn.functional.conv1d(x, state['layers.6.weight'], state['layers.6.bias'])
    return x[:, :, 0]


def frequency_to_bins(frequency, quantize=math.floor):
    cents = OCTAVE * math.log2(frequency / FMIN)
    return int(quantize(cents / CENTS_PER_BIN))


def postprocess(logits, fmin=50., fmax=550.):
    """Mask out-of-range bins; returns (masked logits, probabilities, periodicity)"""
    logits = logits.clone()
    logits[:, :frequency_to_bins(fmin)] = -float('inf')
    logits[:, frequency_to_bins(fmax, math.ceil):] = -float('inf')
    distribution = torch.softmax(logits, dim=1)
    periodicity = 1. + (
        distribution * torch.log(distribution + 1e-7)).sum(dim=1) / math.log(PITCH_BINS)
    return logits, distribution, periodicity


def transition_matrix(hopsize_seconds):
    """Triangular band: max(0, max_bins_per_frame - |i - j|), rows normalised"""
    bins_per_octave = OCTAVE / CENTS_PER_BIN
    max_bins = MAX_OCTAVES_PER_SECOND * hopsize_seconds * bins_per_octave + 1
    index = torch.arange(PITCH_BINS)
    transition = torch.clip(
        max_bins - (index[:, None] - index[None]).abs().float(), min=0.)
    return transition / transition.sum(dim=1, keepdim=True)


def initial_distribution():
    return torch.full((PITCH_BINS,), 1. / PITCH_BINS)


def local_expected_value(bins, logits):
    """bins (F,) int, masked logits (F, 1440) -> Hz (F,)"""
    half = LOCAL_WINDOW // 2
    padded = torch.nn.functional.pad(logits, (half, half), value=-float('inf'))
    offsets = torch.arange(LOCAL_WINDOW)
    indices = bins.long()[:, None] + offsets[None]          # into padded
    selected = padded.gather(1, indices)
    local = torch.softmax(selected, dim=1)
    cents = CENTS_PER_BIN * (indices - half).float()
    expected = (local * cents).sum(dim=1)
    return FMIN * 2. ** (expected / OCTAVE)


def from_audio(
    state,
    audio,
    sample_rate=22050,
    hopsize_seconds=256 / 22050,
    fmin=50.,
    fmax=550.,
    transition=None,
    initial=None
):
    """(1, T) -> pitch (1, F) Hz, periodicity (1, F), plus intermediates"""
    with torch.no_grad():
        x = frames(audio, sample_rate, hopsize_seconds)
        logits = infer(state, x)
        masked, distribution, periodicity = postprocess(logits, fmin, fmax)
        if transition is None:
            transition = transition_matrix(hopsize_seconds)
        if initial is None:
            initial = initial_distribution()
        bins = viterbi.decode(
            distribution[None].numpy(), None, transition.numpy(), initial.numpy(),
            log_probs=False)[0]
        bins = torch.from_numpy(bins)
        pitch = local_expected_value(bins, masked)
    return pitch[None], periodicity[None], {
        'frames': x, 'logits': logits, 'distribution': distribution, 'bins': bins}
