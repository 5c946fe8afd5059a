"""Oracle: generator feature assembly (test infrastructure, see oracle/__init__).

Restates promonet/model/generator.py:49-70 (global features) and :137-197
(frame features) as pure functions of a reference state_dict.
"""
import torch

FMIN = 50.          # promonet/config/defaults.py:27
FMAX = 550.         # :28
MIN_DB = -100.      # :37
REF_DB = 20.        # :46
SAMPLE_RATE = 22050  # :49
LOUDNESS_BANDS = 8  # :90
PITCH_BINS = 256    # :96


def sparsify(ppg, method='percentile', threshold=torch.tensor(0.85)):
    """ppgs.sparsify restatement (third-party, un-vendored: PARITY UNPINNED).

    Call site promonet/model/generator.py:140-147.  percentile: per-frame
    quantile over the phoneme axis (linear interpolation), keep strictly
    greater, renormalise through softmax(log(p + 1e-8)).
    """
    if method == 'percentile':
        cutoff = torch.quantile(
            ppg, torch.as_tensor(threshold, dtype=ppg.dtype), dim=-2, keepdim=True)
        ppg = torch.where(ppg > cutoff, ppg, torch.zeros_like(ppg))
    elif method == 'constant':
        ppg = torch.where(ppg > threshold, ppg, torch.zeros_like(ppg))
    elif method == 'topk':
        k = int(threshold)
        kth = torch.topk(ppg, k, dim=-2).values[..., -1:, :]
        ppg = torch.where(ppg >= kth, ppg, torch.zeros_like(ppg))
    else:
        raise ValueError(f'Sparsification method {method} is not defined')
    return torch.softmax(torch.log(ppg + 1e-8), dim=-2)


def band_average(loudness, bands=LOUDNESS_BANDS):
    """generator.py:172-181 / preprocess/loudness.py:84-111"""
    step = loudness.shape[-2] / bands
    return torch.stack(
        [
            loudness[..., int(band * step):int((band + 1) * step), :].mean(dim=-2)
            for band in range(bands)
        ],
        dim=-2)


def normalize(loudness):
    """preprocess/loudness.py:144-146"""
    return (loudness - MIN_DB) / (REF_DB - MIN_DB)


def pitch_bins(pitch, pitch_distribution):
    """generator.py:153-157: clip, searchsorted (side=left), clip"""
    hz = torch.clip(pitch, FMIN, FMAX)
    bins = torch.searchsorted(pitch_distribution, hz)
    return torch.clip(bins, 0, PITCH_BINS - 1)


def prepare_features(state, loudness, pitch, periodicity, ppg, fargan=False):
    """generator.py:137-197 -> (B, 113 [+1], F)"""
    features = sparsify(ppg, 'percentile', state['ppg_threshold'])
    bins = pitch_bins(pitch, state['pitch_distribution'])
    embedding = torch.nn.functional.embedding(
        bins, state['pitch_embedding.weight']).permute(0, 2, 1)
    features = torch.cat((features, embedding), dim=1)
    features = torch.cat(
        (features, normalize(band_average(loudness))), dim=1)
    features = torch.cat((features, periodicity[:, None]), dim=1)
    if fargan:
        period = SAMPLE_RATE / torch.clip(pitch, FMIN, FMAX)
        features = torch.cat((features, period[:, None]), dim=1)
    return features


def prepare_global_features(state, speakers, spectral_balance_ratios, loudness_ratios):
    """generator.py:49-70 -> (B, 258, 1)"""
    g = torch.nn.functional.embedding(
        speakers, state['speaker_embedding.weight']).unsqueeze(-1)
    g = torch.cat((g, spectral_balance_ratios[:, None, None].to(g.dtype)), dim=1)
    return torch.cat((g, loudness_ratios[:, None, None].to(g.dtype)), dim=1)


def grid_sample(sequence, grid, method='linear'):
    """promonet.edit.grid.sample, promonet/edit/grid.py:12-43 (pinned: in-repo code)"""
    if method == 'linear':
        xp = torch.arange(sequence.shape[-1], device=sequence.device)
        i = torch.searchsorted(xp, grid, side='right')
        fp = torch.nn.functional.pad(sequence, (0, 1), mode='replicate')
        xp = torch.cat((xp, xp[-1:] + 1))
        return fp[..., i - 1] * (xp[i] - grid) + fp[..., i] * (grid - xp[i - 1])
    if method == 'nearest':
        return sequence[..., torch.round(grid).to(torch.long)]
    raise ValueError(f'Grid sampling method {method} is not defined')


def grid_of_length(frames, length):
    """ppgs.edit.grid.of_length (un-vendored, PARITY UNPINNED): `length` positions over [0, T - 1]"""
    return torch.linspace(0., frames - 1., int(length))


def resample_ppg(ppg, length):
    """promonet/preprocess/core.py:97-103: grid resample to `length` frames, then
    softmax(log(p + 1e-8)) over the phoneme axis (of_length: ppgs, un-vendored, PARITY UNPINNED:
    restated as linspace(0, T - 1, length))"""
    grid = grid_of_length(ppg.shape[-1], length)
    return torch.softmax(torch.log(grid_sample(ppg, grid) + 1e-8), -2)
