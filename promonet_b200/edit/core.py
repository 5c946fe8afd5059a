"""Speech-representation editing: drop-in for promonet.edit.from_features
(promonet/edit/core.py:17-132) for constant-ratio time-stretching, pitch-shifting
and loudness-scaling — the edits train.evaluate applies to every validation item
(promonet/train/core.py:640-799).  Each contour is one `pmn_edit_contour` launch
(resample, log2 domain for pitch, scale, shift and clip fused), the PPG one
`pmn_grid_sample` launch."""
import torch

from promonet_b200 import _lib, config
from promonet_b200.edit import grid as grids

__all__ = ['contour', 'from_features', 'from_file']


def contour(sequence, grid=None, log2_domain=False, scale=1., shift=0., clip=None):
    """One edited contour (..., T): resampled at `grid` (in the log2 domain if asked), then
    scale * value + shift, clipped to `clip` = (lo, hi) if given"""
    if not sequence.is_cuda:
        raise RuntimeError('promonet_b200 requires CUDA tensors; there is no CPU path')
    sequence = sequence.to(torch.float32).contiguous()
    t_in = sequence.shape[-1]
    t_out = t_in if grid is None else grid.shape[-1]
    out = torch.empty(*sequence.shape[:-1], t_out, device=sequence.device)
    if sequence.numel() == 0:
        return out
    lo, hi = (0., 0.) if clip is None else clip
    with torch.cuda.device(sequence.device):
        _lib.check(_lib.library().pmn_edit_contour(
            _lib.ptr(sequence), _lib.ptr(grid), _lib.ptr(out), sequence.numel() // t_in, t_in,
            t_out, int(log2_domain), scale, shift, lo, hi, _lib.stream()))
    return out


def from_features(
    loudness,
    pitch,
    periodicity,
    ppg,
    pitch_shift_cents=None,
    time_stretch_ratio=None,
    loudness_scale_db=None,
    stretch_unvoiced=True,
    stretch_silence=True,
    return_grid=False
):
    """Edit speech representation (arguments and order of operations of edit/core.py:17-132;
    inputs are not modified in place).  Leading batch axes are allowed.

    Returns
        edited_loudness, edited_pitch, edited_periodicity, edited_ppg (, grid)
    """
    grid = None
    if time_stretch_ratio is not None:
        if not (stretch_unvoiced and stretch_silence):
            raise NotImplementedError(
                'phoneme-selective stretching needs the phoneme tables of the un-vendored ppgs '
                'package (edit/core.py:59-110); only constant-ratio stretching is accelerated')
        grid = grids.constant(ppg, time_stretch_ratio).to(ppg.device)
        ppg = grids.sample(ppg, grid, config.PPG_INTERP_METHOD)
    shifting = pitch_shift_cents is not None
    if grid is not None or shifting:
        pitch = contour(
            pitch, grid, log2_domain=True,
            scale=2. ** (pitch_shift_cents / 1200.) if shifting else 1.,       # convert.py:64-66
            clip=(config.FMIN, config.FMAX) if shifting else None)
    if grid is not None:
        periodicity = contour(periodicity, grid)
    if grid is not None or loudness_scale_db is not None:
        loudness = contour(loudness, grid, shift=loudness_scale_db or 0.)
    if return_grid:
        return loudness, pitch, periodicity, ppg, grid
    return loudness, pitch, periodicity, ppg


def from_file(loudness_file, pitch_file, periodicity_file, ppg_file, *args, gpu=None, **kwargs):
    """Edit speech representation on disk (edit/core.py:135-179)"""
    from promonet_b200 import load
    device = torch.device('cuda', torch.cuda.current_device() if gpu is None else gpu)
    pitch = torch.load(pitch_file).to(device)
    return from_features(
        torch.load(loudness_file).to(device), pitch, torch.load(periodicity_file).to(device),
        load.ppg(ppg_file, pitch.shape[-1]).to(device), *args, **kwargs)
