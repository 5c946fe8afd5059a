"""Host-side FARGAN generator (config/fargan.py): same call contract as
promonet.model.Generator with MODEL='fargan' (promonet/model/generator.py:116-135,
promonet/model/fargan.py:21-57); the arithmetic is fargan.cu"""
import ctypes

import torch

from promonet_b200 import _lib, config
from promonet_b200.model import init

NUM_PREVIOUS_SAMPLES = 2 * config.HOPSIZE  # HOPSIZE * FARGAN_PREVIOUS_FRAMES, static.py:69-70


class FarganGenerator:

    def __init__(self, device=None, state=None):
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200 requires a CUDA device; there is no CPU path')
        self.device = torch.device(
            'cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.handle = None
        self._workspace = None
        self.default_previous_samples = torch.zeros(
            1, 1, NUM_PREVIOUS_SAMPLES, device=self.device)
        self.load_state_dict(init.fargan_state() if state is None else state)

    def __del__(self):
        handle, self.handle = getattr(self, 'handle', None), None
        if handle and _lib is not None and getattr(_lib, '_library', None) is not None:
            _lib._library.pmn_fargan_destroy(handle)  # no-op at interpreter shutdown

    def load_state_dict(self, state):
        """Accepts promonet.model.Generator().state_dict() keys under config/fargan.py"""
        lib = _lib.library()
        if self.handle:
            lib.pmn_fargan_destroy(self.handle)
        handle = ctypes.c_void_p()
        _lib.check(lib.pmn_fargan_create(ctypes.byref(handle)))
        self.handle = handle
        with torch.cuda.device(self.device):
            keep = []
            for name, tensor in state.items():
                if not tensor.is_floating_point():
                    continue
                value = tensor.detach().to(self.device, torch.float32).contiguous()
                keep.append(value)
                shape = (ctypes.c_int64 * max(1, value.ndim))(*value.shape)
                _lib.check(lib.pmn_fargan_set_tensor(
                    handle, name.encode(), value.data_ptr(), shape, value.ndim, _lib.stream()))
            _lib.check(lib.pmn_fargan_finalize(handle, _lib.stream()))
            torch.cuda.current_stream().synchronize()
        self._state = {k: v.detach().cpu() for k, v in state.items()}
        return self

    def state_dict(self):
        return self._state

    def __call__(
        self,
        loudness,
        pitch,
        periodicity,
        ppg,
        speakers,
        spectral_balance_ratios,
        loudness_ratios,
        previous_samples=None
    ):
        batch, rows, frames = loudness.shape
        if (
            pitch.shape != (batch, frames) or periodicity.shape != (batch, frames) or
            ppg.shape != (batch, config.PPG_CHANNELS, frames) or speakers.shape != (batch,)
        ):
            raise ValueError('inconsistent generator input shapes')
        f32 = dict(device=self.device, dtype=torch.float32)
        loudness = loudness.to(**f32).contiguous()
        pitch = pitch.to(**f32).contiguous()
        periodicity = periodicity.to(**f32).contiguous()
        ppg = ppg.to(**f32).contiguous()
        speakers = speakers.to(self.device, torch.int64).contiguous()
        sbr = spectral_balance_ratios.to(**f32).contiguous()
        lr = loudness_ratios.to(**f32).contiguous()
        previous = None
        if previous_samples is not None and previous_samples.abs().sum() != 0:
            previous = previous_samples.to(**f32).reshape(-1, NUM_PREVIOUS_SAMPLES)
            if previous.shape[0] == 1 and batch > 1:
                previous = previous.expand(batch, -1)
            if previous.shape[0] != batch:
                raise ValueError('previous_samples must be (batch, 1, 512)')
            previous = previous.contiguous()
        audio = torch.empty(batch, 1, frames * config.HOPSIZE, **f32)
        if batch == 0 or frames == 0:
            return audio
        lib = _lib.library()
        with torch.cuda.device(self.device):
            size = lib.pmn_fargan_workspace_bytes(self.handle, batch, frames)
            if self._workspace is None or self._workspace.numel() < size:
                self._workspace = torch.empty(size, dtype=torch.uint8, device=self.device)
            _lib.check(lib.pmn_fargan_forward(
                self.handle, loudness.data_ptr(), rows, pitch.data_ptr(), periodicity.data_ptr(),
                ppg.data_ptr(), speakers.data_ptr(), sbr.data_ptr(), lr.data_ptr(),
                _lib.ptr(previous), audio.data_ptr(), batch, frames,
                self._workspace.data_ptr(), self._workspace.numel(), _lib.stream()))
        return audio

    forward = __call__
