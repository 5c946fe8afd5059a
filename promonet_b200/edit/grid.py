"""1-D grid sampling: drop-in for promonet.edit.grid.sample / of_length / constant
(promonet/edit/grid.py:12-69), the resampling step in front of
promonet.synthesize.from_features on the file path (promonet/load.py:172-188,
promonet/synthesize/core.py:95)."""
import torch

from promonet_b200 import _lib


def sample(sequence, grid, method='linear', renormalize=False):
    """Perform 1D grid-based sampling of sequence (..., T) at the (fractional) frame
    positions `grid` (T_out,).  renormalize=True also applies the distribution-preserving
    softmax(log(p + 1e-8)) over dim -2 (promonet/preprocess/core.py:97-103) in the same kernel."""
    if method not in ('linear', 'nearest'):
        raise ValueError(f'Grid sampling method {method} is not defined')
    if not sequence.is_cuda:
        raise RuntimeError('promonet_b200 requires CUDA tensors; there is no CPU path')
    sequence = sequence.to(torch.float32).contiguous()
    grid = grid.to(sequence.device, torch.float32).contiguous()
    t_in, t_out = sequence.shape[-1], grid.shape[-1]
    channels = sequence.shape[-2] if (renormalize and sequence.ndim >= 2) else 1
    items = sequence.numel() // (channels * t_in)
    out = torch.empty(*sequence.shape[:-1], t_out, device=sequence.device)
    _lib.check(_lib.library().pmn_grid_sample(
        _lib.ptr(sequence), _lib.ptr(grid), _lib.ptr(out), items, channels, t_in, t_out,
        int(method == 'nearest'), int(renormalize), _lib.stream()))
    return out


def of_length(tensor, length):
    """Time-stretch grid of a specified length (ppgs.edit.grid.of_length: un-vendored,
    restated as the uniform grid over [0, T - 1])"""
    return torch.linspace(
        0., tensor.shape[-1] - 1., int(length), dtype=torch.float32, device=tensor.device)


def constant(tensor, ratio):
    """Grid for constant-ratio time-stretching (ppgs.edit.grid.constant: un-vendored, [RECALLED]
    of_length with round(T / ratio + 1e-4) frames -- the frame count the reference's own
    selective-stretch branch uses, promonet/edit/core.py:82)"""
    return of_length(tensor, round(tensor.shape[-1] / ratio + 1e-4))
