"""Oracle: FARGAN generator forward (test infrastructure, see oracle/__init__).

Functional restatement of promonet/model/fargan.py over a reference state dict
(config/fargan.py: MODEL='fargan').  Inference only: additive noise is a
training-time feature (fargan.py:396-403) and gain normalisation is off
(config/defaults.py:238).  Pinned against the reference module through
tests/golden/fargan.npz (oracle/make_golden.py --fargan).
"""
import torch

from oracle import features
from oracle.hifigan import fold_weight_norm

HOPSIZE = 256
SUBFRAMES = 4           # FARGAN_SUBFRAMES, config/defaults.py:244
SUBFRAME_SIZE = 64      # :247
PREVIOUS_SAMPLES = 512  # NUM_PREVIOUS_SAMPLES, config/static.py:69-70


def linear_weight(state, prefix):
    if f'{prefix}.weight' in state:
        return state[f'{prefix}.weight']
    return fold_weight_norm(state[f'{prefix}.weight_g'], state[f'{prefix}.weight_v'])


def glu(state, prefix, x):
    """GLU fargan.py:375-388"""
    return x * torch.sigmoid(x @ linear_weight(state, f'{prefix}.gate').T)


def gru_cell(state, prefix, x, h):
    """torch.nn.GRUCell(bias=False): gates r, z, n"""
    gi = x @ state[f'{prefix}.weight_ih'].T
    gh = h @ state[f'{prefix}.weight_hh'].T
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1. - z) * n + z * h


def conditioning(state, x, prefix='model.conditioning_network'):
    """ConditioningNetwork fargan.py:139-160"""
    for i in (0, 2, 4):
        x = torch.tanh(x @ state[f'{prefix}.{i}.weight'].T)
    return x


def subframe(state, feats, previous_samples, period, states, prefix='model.subframe_network'):
    """SubframeNetwork.forward fargan.py:199-335 (eval mode)"""
    total = previous_samples.shape[-1]
    index = total - period[:, None] + torch.arange(SUBFRAME_SIZE + 4)[None] - 2
    index = index - period[:, None] * (index >= total)
    lookback = torch.gather(previous_samples.squeeze(1), 1, index)
    previous = previous_samples[:, 0, -SUBFRAME_SIZE:]
    inputs = torch.cat((feats, previous, lookback), dim=1)
    fw = torch.tanh(
        torch.cat((inputs, states[3]), -1) @
        linear_weight(state, f'{prefix}.framewise_convolution.model.0').T)
    fw = glu(state, f'{prefix}.framewise_convolution.model.2', fw)
    lookback = lookback[:, 2:-2]
    outs, new_states, x = [], [], fw
    for i in (1, 2, 3):
        h = gru_cell(
            state, f'{prefix}.gru{i}', torch.cat([x, lookback, previous], dim=1), states[i - 1])
        new_states.append(h)
        x = glu(state, f'{prefix}.gru{i}_glu', h)
        outs.append(x)
    skip = torch.cat(outs + [fw, lookback, previous], dim=1)
    skip = glu(
        state, f'{prefix}.skip_glu',
        torch.tanh(skip @ state[f'{prefix}.skip_dense.weight'].T))
    output = torch.tanh(skip @ state[f'{prefix}.output_layer.weight'].T)
    return output, (*new_states, inputs)


def vocoder(state, feats, global_features, previous_samples):
    """FARGAN.forward fargan.py:21-57: feats (B, 114, F) -> (B, 1, 256 F)"""
    batch = feats.shape[0]
    dtype = feats.dtype
    states = (
        torch.zeros(batch, HOPSIZE, dtype=dtype), torch.zeros(batch, HOPSIZE, dtype=dtype),
        torch.zeros(batch, HOPSIZE, dtype=dtype),
        torch.zeros(batch, 4 * SUBFRAME_SIZE + 4, dtype=dtype))
    g = global_features.squeeze(2)
    signal = []
    for frame in feats.permute(2, 0, 1):
        period = torch.round(frame[:, -1]).to(torch.long)
        cond = conditioning(state, torch.cat((frame[:, :-1], g), dim=1))
        for sub in cond.reshape(batch, 2 * SUBFRAME_SIZE, SUBFRAMES).permute(2, 0, 1):
            out, states = subframe(state, sub, previous_samples, period, states)
            signal.append(out)
            previous_samples = torch.cat(
                [previous_samples[:, :, SUBFRAME_SIZE:], out[:, None]], dim=2)
    return torch.cat(signal, dim=1).unsqueeze(1)


def generator(
    state, loudness, pitch, periodicity, ppg, speakers, spectral_balance_ratios,
    loudness_ratios, previous_samples=None
):
    """Generator.forward generator.py:116-135 with MODEL='fargan'"""
    x = features.prepare_features(state, loudness, pitch, periodicity, ppg, fargan=True)
    g = features.prepare_global_features(state, speakers, spectral_balance_ratios, loudness_ratios)
    if previous_samples is None:
        previous_samples = torch.zeros(x.shape[0], 1, PREVIOUS_SAMPLES, dtype=x.dtype)
    return vocoder(state, x, g, previous_samples)
