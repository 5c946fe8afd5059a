"""Host-side Generator: owns a libpromonet_b200 generator handle

Mirrors the interface of promonet.model.Generator
(promonet/model/generator.py:84-135): `Generator()` is randomly initialised,
`load_state_dict` accepts the reference's checkpoint keys (weight_g/weight_v
pairs), and calling it takes the same eight arguments and returns (B, 1, T).
All arithmetic happens in the CUDA library; torch only owns the memory.
"""
import os

import torch

from promonet_b200 import _lib, config
from promonet_b200.model import init


# "fp16 + 2 x fp8" residual blocks at C = 128 (pmn_generator_set_f8) unless asked otherwise
F8_DEFAULT = True


class Generator:

    def __init__(self, device=None, math=_lib.MATH_BF16X3_TC, state=None, pair_mask=None, f8=None):
        """pair_mask: which residual blocks run as fused pair kernels
        (pmn_generator_set_pair_mask); None = the library default, or the
        PMN_PAIR_MASK environment variable (an experiment knob: outputs agree
        within fp32 rounding for every mask).
        f8: residual blocks of the C = 128 stage with "fp16 + 2 x fp8" operands
        (pmn_generator_set_f8); None = the PMN_GENERATOR_F8 environment variable, else F8_DEFAULT"""
        if not torch.cuda.is_available():
            raise RuntimeError(
                'promonet_b200.model.Generator needs a CUDA device (sm_100a); '
                'there is no CPU path')
        self.device = torch.device(
            'cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.math = math
        if pair_mask is None and os.environ.get('PMN_PAIR_MASK'):
            pair_mask = int(os.environ['PMN_PAIR_MASK'], 0)
        self.pair_mask = pair_mask
        if f8 is None:
            f8 = os.environ.get('PMN_GENERATOR_F8', '1' if F8_DEFAULT else '0') == '1'
        self.f8 = bool(f8) and math == _lib.MATH_BF16X3_TC
        self.handle = None
        self.default_previous_samples = torch.zeros(1, 1, 1, device=self.device)
        self._workspace = None
        self._staging = None
        self._streaming = None
        self._state = None
        self.load_state_dict(init.hifigan_state() if state is None else state)

    def __del__(self):
        self._destroy()

    def _destroy(self):
        handle, self.handle = getattr(self, 'handle', None), None
        if handle and _lib is not None and getattr(_lib, '_library', None) is not None:
            _lib._library.pmn_generator_destroy(handle)  # no-op at interpreter shutdown

    ###########################################################################
    # Weights
    ###########################################################################

    def state_dict(self):
        return self._state

    def load_state_dict(self, state):
        """Accepts promonet.model.Generator().state_dict() (hifigan) keys"""
        import ctypes
        lib = _lib.library()
        self._destroy()
        handle = ctypes.c_void_p()
        _lib.check(lib.pmn_generator_create(ctypes.byref(handle)))
        self.handle = handle
        with torch.cuda.device(self.device):
            stream = _lib.stream()
            keep = []
            for name, tensor in state.items():
                if not tensor.is_floating_point():
                    continue
                value = tensor.detach().to(
                    self.device, torch.float32).contiguous()
                keep.append(value)
                shape = (ctypes.c_int64 * max(1, value.ndim))(*value.shape)
                _lib.check(lib.pmn_generator_set_tensor(
                    handle, name.encode(), value.data_ptr(), shape, value.ndim,
                    stream))
            _lib.check(lib.pmn_generator_finalize(handle, self.math, stream))
            if self.pair_mask is not None:
                _lib.check(lib.pmn_generator_set_pair_mask(handle, self.pair_mask))
            if self.f8:
                _lib.check(lib.pmn_generator_set_f8(handle, 1))
            torch.cuda.current_stream().synchronize()
        self._state = {k: v.detach().cpu() for k, v in state.items()}
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('promonet_b200 has no CPU path')
        if device != self.device:
            self.device = device
            self.default_previous_samples = self.default_previous_samples.to(device)
            self._workspace = self._staging = self._streaming = None
            self.load_state_dict(self._state)
        return self

    def eval(self):
        return self

    ###########################################################################
    # Forward
    ###########################################################################

    def workspace(self, batch, frames):
        size = _lib.library().pmn_generator_workspace_bytes(
            self.handle, batch, frames)
        if self._workspace is None or self._workspace.numel() < size:
            self._workspace = torch.empty(
                size, dtype=torch.uint8, device=self.device)
        return self._workspace

    def _check_inputs(self, loudness, pitch, periodicity, ppg, speakers, sbr, lr):
        if loudness.ndim != 3 or ppg.ndim != 3 or pitch.ndim != 2:
            raise ValueError(
                'expected loudness (B, 8|513, F), pitch (B, F), '
                'periodicity (B, F), ppg (B, 40, F)')
        batch, _, frames = loudness.shape
        if (
            pitch.shape != (batch, frames) or
            periodicity.shape != (batch, frames) or
            ppg.shape != (batch, config.PPG_CHANNELS, frames) or
            speakers.shape != (batch,) or
            sbr.shape != (batch,) or
            lr.shape != (batch,)
        ):
            raise ValueError('inconsistent generator input shapes')
        return batch, frames

    def __call__(
        self,
        loudness,
        pitch,
        periodicity,
        ppg,
        speakers,
        spectral_balance_ratios,
        loudness_ratios,
        previous_samples=None,
        out=None
    ):
        """Generator.forward (generator.py:116-135) on device tensors; `out` (B, 1, 256 F)
        receives the audio when given"""
        batch, frames = self._check_inputs(
            loudness, pitch, periodicity, ppg, speakers,
            spectral_balance_ratios, loudness_ratios)
        f32 = dict(device=self.device, dtype=torch.float32)
        loudness = loudness.to(**f32).contiguous()
        pitch = pitch.to(**f32).contiguous()
        periodicity = periodicity.to(**f32).contiguous()
        ppg = ppg.to(**f32).contiguous()
        speakers = speakers.to(self.device, torch.int64).contiguous()
        sbr = spectral_balance_ratios.to(**f32).contiguous()
        lr = loudness_ratios.to(**f32).contiguous()
        audio = torch.empty(
            batch, 1, frames * config.HOPSIZE, **f32) if out is None else out
        if out is not None and (
                out.shape != (batch, 1, frames * config.HOPSIZE) or out.dtype != torch.float32 or
                out.device != self.device or not out.is_contiguous()):
            raise ValueError('out must be a contiguous float32 (B, 1, 256 F) tensor on the model device')
        if batch == 0 or frames == 0:
            return audio
        with torch.cuda.device(self.device):
            workspace = self.workspace(batch, frames)
            _lib.check(_lib.library().pmn_generator_forward(
                self.handle,
                loudness.data_ptr(), loudness.shape[1],
                pitch.data_ptr(), periodicity.data_ptr(), ppg.data_ptr(),
                speakers.data_ptr(), sbr.data_ptr(), lr.data_ptr(),
                audio.data_ptr(), batch, frames,
                workspace.data_ptr(), workspace.numel(), _lib.stream()))
        return audio

    forward = __call__

    def forward_host(
        self,
        loudness,
        pitch,
        periodicity,
        ppg,
        speakers,
        spectral_balance_ratios,
        loudness_ratios,
        out=None
    ):
        """Same forward over HOST tensors (ideally pinned): the library copies
        inputs H2D, synthesizes, and copies the audio D2H into `out`"""
        batch, frames = self._check_inputs(
            loudness, pitch, periodicity, ppg, speakers,
            spectral_balance_ratios, loudness_ratios)
        tensors = [
            t.to(torch.float32).contiguous() for t in
            (loudness, pitch, periodicity, ppg, spectral_balance_ratios, loudness_ratios)]
        speakers = speakers.to(torch.int64).contiguous()
        if any(t.is_cuda for t in tensors) or speakers.is_cuda:
            raise ValueError('forward_host takes host tensors')
        if out is None:
            out = torch.empty(
                batch, 1, frames * config.HOPSIZE, dtype=torch.float32,
                pin_memory=True)
        if batch == 0 or frames == 0:
            return out
        lib = _lib.library()
        with torch.cuda.device(self.device):
            size = lib.pmn_generator_staging_bytes(batch, frames, loudness.shape[1])
            if self._staging is None or self._staging.numel() < size:
                self._staging = torch.empty(size, dtype=torch.uint8, device=self.device)
            workspace = self.workspace(batch, frames)
            loud, pit, per, pp, sbr, lr = tensors
            _lib.check(lib.pmn_generator_forward_host(
                self.handle,
                loud.data_ptr(), loud.shape[1],
                pit.data_ptr(), per.data_ptr(), pp.data_ptr(),
                speakers.data_ptr(), sbr.data_ptr(), lr.data_ptr(),
                out.data_ptr(), batch, frames,
                self._staging.data_ptr(), self._staging.numel(),
                workspace.data_ptr(), workspace.numel(), _lib.stream()))
            torch.cuda.current_stream().synchronize()
        return out

    def stream_host(self, batches, depth=2):
        """Synthesize a stream of HOST batches (tuples of the seven pinned host tensors that
        forward_host takes), yielding one pinned host audio tensor (B, 1, 256 F) per batch, in
        order.  Copies run on their own streams: the device-to-host copy of batch i and the
        host-to-device copy of batch i + 1 overlap the synthesis of batch i + 1 / i + 2, so the
        GPU never waits for PCIe.  A yielded tensor is reused `depth` batches later: consume
        (or copy) it before advancing the generator that far."""
        if depth < 2:
            raise ValueError('depth must be at least 2')
        compute = torch.cuda.current_stream(self.device)
        # streams and the slots' device / pinned buffers live with the model: allocating pinned
        # memory is slow and synchronizes the device, so a second stream of batches reuses them
        if getattr(self, '_streaming', None) is None or len(self._streaming['slots']) != depth:
            self._streaming = dict(
                copy_in=torch.cuda.Stream(self.device), copy_out=torch.cuda.Stream(self.device),
                slots=[dict(inputs=None, audio=None, host=None, done=None) for _ in range(depth)])
        copy_in, copy_out = self._streaming['copy_in'], self._streaming['copy_out']
        slots = self._streaming['slots']
        for slot in slots:
            if slot['done'] is not None:
                slot['done'].synchronize()
        pending = []
        for index, batch in enumerate(batches):
            slot = slots[index % depth]
            if slot['done'] is not None:
                # its previous audio has been yielded (depth >= 2) and its copy has finished
                slot['done'].synchronize()
            batch_size, frames = self._check_inputs(*batch)
            with torch.cuda.stream(copy_in):
                # the slot's device inputs were last read by a forward that has been waited for
                slot['inputs'] = [
                    t.to(self.device, torch.int64 if t.dtype in (torch.int32, torch.int64) else torch.float32,
                         non_blocking=True) for t in batch]
                ready = torch.cuda.Event()
                ready.record(copy_in)
            shape = (batch_size, 1, frames * config.HOPSIZE)
            if slot['audio'] is None or slot['audio'].shape != shape:
                slot['audio'] = torch.empty(shape, device=self.device)
                slot['host'] = torch.empty(shape, pin_memory=True)
            compute.wait_event(ready)
            self(*slot['inputs'], out=slot['audio'])
            for tensor in slot['inputs']:
                tensor.record_stream(compute)
            computed = torch.cuda.Event()
            computed.record(compute)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(computed)
                slot['host'].copy_(slot['audio'], non_blocking=True)
                slot['done'] = torch.cuda.Event()
                slot['done'].record(copy_out)
            # the next forward may not overwrite this slot's audio before its copy out: with
            # depth slots that is guaranteed by the synchronize at the top of the loop
            pending.append(slot)
            if len(pending) == depth:
                first = pending.pop(0)
                first['done'].synchronize()
                yield first['host']
        for slot in pending:
            slot['done'].synchronize()
            yield slot['host']

    def features(self, loudness, pitch, periodicity, ppg):
        """Generator.prepare_features (generator.py:137-197) -> (B, 113, F)"""
        batch, _, frames = loudness.shape
        f32 = dict(device=self.device, dtype=torch.float32)
        loudness = loudness.to(**f32).contiguous()
        pitch = pitch.to(**f32).contiguous()
        periodicity = periodicity.to(**f32).contiguous()
        ppg = ppg.to(**f32).contiguous()
        out = torch.empty(batch, config.NUM_FEATURES, frames, **f32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.library().pmn_generator_features(
                self.handle, loudness.data_ptr(), loudness.shape[1],
                pitch.data_ptr(), periodicity.data_ptr(), ppg.data_ptr(),
                out.data_ptr(), batch, frames, _lib.stream()))
        return out
