"""Pitch and periodicity: drop-in for the `penn.from_audio` call at
promonet/preprocess/core.py:71-81 (FCNF0++, center='half-hop', Viterbi decoding).

penn is a third-party dependency whose source and pretrained weights are not
available offline: the pipeline follows the published algorithm (SURVEY
Appendix C; exact choices written down in oracle/penn.py) and `checkpoint=None`
means a seeded random initialisation.  Everything numerical runs in pitch.cu /
viterbi.cu through the C ABI."""
import ctypes
import math
from collections import OrderedDict

import torch

from promonet_b200 import _lib, config

__all__ = ['from_audio', 'Model', 'init_state', 'transition_matrix', 'initial_distribution']

SAMPLE_RATE = 8000
WINDOW_SIZE = 1024
PITCH_BINS = 1440
CENTS_PER_BIN = 5.
OCTAVE = 1200.
MAX_OCTAVES_PER_SECOND = 32.
LAYERS = (  # (c_in, c_out, length after the block)
    (1, 256, 481), (256, 32, 225), (32, 32, 97), (32, 128, 66), (128, 256, 35), (256, 512, 4))


def init_state(seed=config.RANDOM_SEED):
    """Random FCNF0++ parameters: torch default Conv1d init, LayerNorm affine
    perturbed around (1, 0) so that it is exercised"""
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    state = OrderedDict()
    for i, (c_in, c_out, length) in enumerate(LAYERS):
        conv = torch.nn.Conv1d(c_in, c_out, 32)
        state[f'layers.{i}.conv.weight'] = conv.weight.detach().clone()
        state[f'layers.{i}.conv.bias'] = conv.bias.detach().clone()
        state[f'layers.{i}.norm.weight'] = 1. + 0.1 * torch.randn(c_out, length)
        state[f'layers.{i}.norm.bias'] = 0.1 * torch.randn(c_out, length)
    conv = torch.nn.Conv1d(512, PITCH_BINS, 4)
    state['layers.6.weight'] = conv.weight.detach().clone()
    state['layers.6.bias'] = conv.bias.detach().clone()
    torch.random.set_rng_state(rng_state)
    return state


def convert_state(state):
    """Checkpoint state dict -> this module's names.  Accepts our own names
    (`layers.N.conv.weight`, ...) and the names of upstream penn's FCNF0++ as far
    as they can be known offline (penn is not installed: [RECALLED] its model is a
    Sequential of Sequential blocks, so keys look like `N.M.weight` with the Conv1d
    at M = 0 and the LayerNorm after the ReLU / MaxPool, and the final Conv1d at
    `6.weight`, possibly under a `model.` / `module.` prefix).  The mapping goes by
    block index and tensor rank -- rank 3 = convolution, rank 2 = LayerNorm
    (C, length) -- not by M, and every tensor must land on a slot of the right shape."""
    import re
    expected = init_state(0)
    if all(name in state for name in expected):
        return OrderedDict((name, state[name]) for name in expected)
    out = OrderedDict()
    for name, tensor in state.items():
        if not torch.is_tensor(tensor):
            continue
        match = re.search(r'(?:^|\.)(\d+)\.(?:(\d+)\.)?(weight|bias)$', name)
        if not match:
            continue
        block, kind = int(match.group(1)), match.group(3)
        if block == 6:
            target = f'layers.6.{kind}'
        elif tensor.ndim == 3 or (kind == 'bias' and tensor.ndim == 1):
            target = f'layers.{block}.conv.{kind}'
        else:
            target = f'layers.{block}.norm.{kind}'
        out[target] = tensor
    missing = [name for name in expected if name not in out]
    wrong = [name for name in expected if name in out and out[name].shape != expected[name].shape]
    if missing or wrong:
        raise ValueError(
            f'not an FCNF0++ checkpoint: missing {missing[:4]}, wrong shape {wrong[:4]} '
            f'(keys seen: {list(state)[:6]})')
    return OrderedDict((name, out[name]) for name in expected)


def load_checkpoint(file):
    """torchutil.checkpoint file ({'model': state_dict, ...}) or a bare state dict"""
    state = torch.load(file, map_location='cpu')
    for key in ('model', 'state_dict'):
        if isinstance(state, dict) and key in state and isinstance(state[key], dict):
            state = state[key]
    return convert_state(state)


def transition_matrix(hopsize_seconds):
    """Triangular-band pitch transition: max(0, max_bins_per_frame - |i - j|), rows normalised"""
    max_bins = MAX_OCTAVES_PER_SECOND * hopsize_seconds * (OCTAVE / CENTS_PER_BIN) + 1
    index = torch.arange(PITCH_BINS)
    transition = torch.clip(
        max_bins - (index[:, None] - index[None]).abs().float(), min=0.)
    return transition / transition.sum(dim=1, keepdim=True)


def initial_distribution():
    return torch.full((PITCH_BINS,), 1. / PITCH_BINS)


class Model:
    """Owns a libpromonet_b200 pitch-network handle"""

    def __init__(self, device=None, state=None, math=_lib.MATH_BF16X3_TC):
        self.math = math
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
        self.device = torch.device(
            'cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.handle = None
        self._workspace = None
        self._tables = {}
        self.load_state_dict(init_state() if state is None else state)

    def __del__(self):
        handle, self.handle = getattr(self, 'handle', None), None
        if handle and _lib is not None and getattr(_lib, '_library', None) is not None:
            _lib._library.pmn_pitch_destroy(handle)  # no-op at interpreter shutdown

    def load_state_dict(self, state):
        lib = _lib.library()
        if self.handle:
            lib.pmn_pitch_destroy(self.handle)
        handle = ctypes.c_void_p()
        _lib.check(lib.pmn_pitch_create(ctypes.byref(handle)))
        self.handle = handle
        with torch.cuda.device(self.device):
            keep = []
            for name, tensor in state.items():
                value = tensor.detach().to(self.device, torch.float32).contiguous()
                keep.append(value)
                shape = (ctypes.c_int64 * max(1, value.ndim))(*value.shape)
                _lib.check(lib.pmn_pitch_set_tensor(
                    handle, name.encode(), value.data_ptr(), shape, value.ndim, _lib.stream()))
            _lib.check(lib.pmn_pitch_finalize(handle, self.math, _lib.stream()))
            torch.cuda.current_stream().synchronize()
        return self

    def decoder_tables(self, hopsize_seconds):
        key = round(hopsize_seconds * 1e9)
        if key not in self._tables:
            self._tables[key] = (
                transition_matrix(hopsize_seconds).to(self.device).contiguous(),
                initial_distribution().to(self.device).contiguous())
        return self._tables[key]

    def __call__(
        self,
        audio,
        sample_rate=config.SAMPLE_RATE,
        hopsize=config.HOPSIZE / config.SAMPLE_RATE,
        fmin=config.FMIN,
        fmax=config.FMAX,
        batch_size=2048,
        transition=None,
        initial=None,
        return_intermediates=False
    ):
        """audio (B, T) -> pitch (B, F) Hz, periodicity (B, F)"""
        audio = audio.to(self.device, torch.float32).contiguous()
        if audio.ndim != 2:
            raise ValueError('audio must be (batch, samples)')
        batch, samples = audio.shape
        lib = _lib.library()
        frames = lib.pmn_pitch_frames(samples, int(sample_rate), float(hopsize))
        default_transition, default_initial = self.decoder_tables(hopsize)
        transition = default_transition if transition is None else transition.to(
            self.device, torch.float32).contiguous()
        initial = default_initial if initial is None else initial.to(
            self.device, torch.float32).contiguous()
        pitch = torch.empty(batch, frames, device=self.device)
        periodicity = torch.empty(batch, frames, device=self.device)
        logits = bins = None
        if return_intermediates:
            logits = torch.empty(batch, frames, PITCH_BINS, device=self.device)
            bins = torch.empty(batch, frames, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            size = lib.pmn_pitch_workspace_bytes(
                batch, samples, int(sample_rate), float(hopsize), int(batch_size))
            if self._workspace is None or self._workspace.numel() < size:
                self._workspace = torch.empty(size, dtype=torch.uint8, device=self.device)
            _lib.check(lib.pmn_pitch_forward(
                self.handle, audio.data_ptr(), batch, samples, int(sample_rate), float(hopsize),
                float(fmin), float(fmax), transition.data_ptr(), initial.data_ptr(),
                pitch.data_ptr(), periodicity.data_ptr(), _lib.ptr(logits), _lib.ptr(bins),
                int(batch_size), self._workspace.data_ptr(), self._workspace.numel(),
                _lib.stream()))
        if return_intermediates:
            return pitch, periodicity, logits, bins
        return pitch, periodicity


def from_audio(
    audio,
    sample_rate=config.SAMPLE_RATE,
    hopsize=config.HOPSIZE / config.SAMPLE_RATE,
    fmin=config.FMIN,
    fmax=config.FMAX,
    checkpoint=None,
    batch_size=2048,
    center='half-hop',
    decoder='viterbi',
    interp_unvoiced_at=None,
    gpu=None
):
    """penn.from_audio: audio (1, T) -> pitch (1, F), periodicity (1, F)"""
    if center != 'half-hop' or decoder != 'viterbi' or interp_unvoiced_at is not None:
        raise ValueError(
            "only center='half-hop', decoder='viterbi', interp_unvoiced_at=None "
            '(the configuration promonet uses) is implemented')
    device = torch.device('cuda', torch.cuda.current_device() if gpu is None else gpu)
    if (
        not hasattr(from_audio, 'model') or
        from_audio.checkpoint != checkpoint or
        from_audio.device != device
    ):
        state = None if checkpoint is None else load_checkpoint(checkpoint)
        from_audio.model = Model(device=device, state=state)
        from_audio.checkpoint = checkpoint
        from_audio.device = device
    return from_audio.model(
        audio, sample_rate, hopsize, fmin, fmax, batch_size or 2048)
