"""Viterbi decoding: drop-in for torbi.from_probabilities as the reference calls it
(promonet/preprocess/harmonics.py:270-276); runs in viterbi.cu"""
import torch

from promonet_b200 import _lib

__all__ = ['from_probabilities']


def from_probabilities(
    observation,
    batch_frames=None,
    transition=None,
    initial=None,
    log_probs=False,
    gpu=None,
    num_threads=None
):
    """Decode a time-varying categorical distribution

    observation (B, T, S); batch_frames (B,) valid lengths; transition (S, S)
    [row i -> column j]; initial (S,); returns indices (B, T) int32 on the device"""
    if not torch.cuda.is_available():
        raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
    device = torch.device(
        'cuda', torch.cuda.current_device() if gpu is None else gpu
    ) if not observation.is_cuda else observation.device
    if observation.ndim != 3:
        raise ValueError('observation must be (batch, frames, states)')
    observation = observation.to(device, torch.float32).contiguous()
    batch, frames, states = observation.shape
    uniform = -torch.log(torch.tensor(float(states))) if log_probs else 1. / states
    if transition is None:
        transition = torch.full((states, states), float(uniform))
    if initial is None:
        initial = torch.full((states,), float(uniform))
    transition = transition.to(device, torch.float32).contiguous()
    initial = initial.to(device, torch.float32).contiguous()
    if transition.shape != (states, states) or initial.shape != (states,):
        raise ValueError('transition must be (states, states) and initial (states,)')
    if batch_frames is not None:
        batch_frames = batch_frames.to(device, torch.int32).contiguous()
    indices = torch.empty(batch, frames, dtype=torch.int32, device=device)
    lib = _lib.library()
    with torch.cuda.device(device):
        size = lib.pmn_viterbi_workspace_bytes(batch, frames, states)
        workspace = torch.empty(size, dtype=torch.uint8, device=device)
        _lib.check(lib.pmn_viterbi_decode(
            observation.data_ptr(), _lib.ptr(batch_frames), transition.data_ptr(),
            initial.data_ptr(), int(bool(log_probs)), indices.data_ptr(), batch, frames, states,
            workspace.data_ptr(), size, _lib.stream()))
    return indices
