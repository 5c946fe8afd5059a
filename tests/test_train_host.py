"""Host-side logic of the training path that needs no GPU: flat parameter storage,
convolution geometry, weight-table packing (no kernel is launched here)"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from promonet_b200 import _lib
from promonet_b200.model import init
from promonet_b200.train import ops
from promonet_b200.train.params import ALIGN, ParamSet


def test_param_set_layout_and_round_trip():
    state = init.hifigan_state(7)
    buffers = ('default_previous_samples', 'ppg_threshold', 'pitch_distribution')
    params = ParamSet(state, 'cpu', buffers)
    assert params.peers is None
    names = [k for k in state if k not in buffers]
    assert params.names() == names
    total = 0
    for name in names:
        offset, shape = params.index[name]
        assert offset % ALIGN == 0 and offset >= total      # 16-byte aligned, in order, no overlap
        assert tuple(state[name].shape) == shape
        assert torch.equal(params[name], state[name])
        total = offset + state[name].numel()
    assert params.numel >= total and params.numel % ALIGN == 0
    assert sum(v.numel() for k, v in state.items() if k not in buffers) == 14230784
    # views alias the flat buffers
    params.gradient('model.input_feature_conv.bias').fill_(3.)
    offset, _ = params.index['model.input_feature_conv.bias']
    assert float(params.grad[offset]) == 3.
    params.zero_grad()
    assert float(params.grad.abs().sum()) == 0.
    restored = params.state_dict()
    assert list(restored) == names + list(buffers)
    assert all(torch.equal(restored[k], state[k]) for k in state)
    other = ParamSet(init.hifigan_state(8), 'cpu', buffers)
    other.load_state_dict(restored)
    assert torch.equal(other.data, params.data)


@pytest.mark.parametrize('size,kernel,stride,dilation,padding', [
    ((700, 1), (11, 1), 1, (5, 1), (25, 0)),
    ((8193, 2), (5, 1), (3, 1), 1, (2, 0)),
    ((64, 129), (3, 9), (1, 2), 1, (1, 4)),
    ((64, 17), (3, 3), 1, 1, (1, 1)),
    ((16384, 1), (41, 1), (4, 1), 1, (20, 0)),
])
def test_geometry_matches_torch_output_shape(size, kernel, stride, dilation, padding):
    y = F.conv2d(torch.zeros(1, 1, *size), torch.zeros(1, 1, *kernel), None, stride, padding, dilation)
    g = ops.geometry(3, 1, 1, size, kernel, stride, dilation, padding)
    assert (g.h_out, g.w_out) == tuple(y.shape[2:])
    assert (g.batch, g.channel_stride, g.position_stride, g.batch_stride) == (3, 0, 0, 0)
    assert ctypes.sizeof(g) == 18 * ctypes.sizeof(ctypes.c_int)    # pmn_conv_geometry


def test_transposed_convolution_geometry():
    """A ConvTranspose1d (k = 2 s, padding s / 2) is described by the convolution it transposes"""
    for k, s, t in ((16, 8, 40), (4, 2, 300)):
        y = F.conv_transpose1d(torch.zeros(1, 1, t), torch.zeros(1, 1, k), None, s, (k - s) // 2)
        g = ops.geometry(1, 1, 1, (t * s, 1), (k, 1), (s, 1), 1, ((k - s) // 2, 0), size_out=(t, 1))
        assert y.shape[-1] == g.h_in == t * s and g.h_out == t


def test_weight_table_descriptor_layout():
    assert ctypes.sizeof(_lib.WeightDesc) == 7 * ctypes.sizeof(ctypes.c_void_p) + 4 * ctypes.sizeof(ctypes.c_int)
    v, g = torch.zeros(8, 4, 3), torch.zeros(8, 1, 1)
    table = ops.weight_table(
        [{'v': v, 'g': g, 'w': torch.zeros(8, 4, 3), 'dim0': 8, 'dim1': 4, 'taps': 3}], 'cpu')
    assert table.dtype == torch.uint8 and table.numel() == ctypes.sizeof(_lib.WeightDesc)
    desc = _lib.WeightDesc.from_buffer_copy(bytes(table.numpy()))
    assert desc.v == v.data_ptr() and desc.g == g.data_ptr() and desc.packed is None
    assert (desc.dim0, desc.dim1, desc.taps, desc.groups) == (8, 4, 3, 1)


def test_packed_weight_sizes():
    # [128-row tile][tap][32-channel block][8][rows][4]: rows and channels padded
    assert ops.packed_floats(1024, 1024, 5) == 1024 * 5 * 1024
    assert ops.packed_floats(512, 113, 7) == 512 * 7 * 128
    assert ops.packed_floats(1, 32, 7) == 32 * 7 * 32
    assert ops.packed_floats(100, 40, 3) == 128 * 3 * 64
    assert ops.channel_pad(113) == 128


@pytest.mark.parametrize('flags,modules,convs', [
    ({}, 5, 5 * 6 + 26), ({'multi_scale': True}, 6, 5 * 6 + 7 + 26),
    ({'multi_resolution': True}, 8, 5 * 6 + 3 * 6 + 26),
    ({'multi_scale': True, 'multi_resolution': True}, 9, 5 * 6 + 7 + 3 * 6 + 26)])
def test_discriminator_module_graph_builds_for_every_flag(monkeypatch, flags, modules, convs):
    """Construction only (flat storage, layer views, weight table: no kernel is launched), on the
    CPU with the CUDA check bypassed: the default graph and the flagged ones stay wired"""
    from promonet_b200.train import discriminator
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    state = init.discriminator_state(3, **flags)
    for math in ('fp32', 'tf32'):
        module = discriminator.Discriminator(state, 'cpu', math)
        assert len(module.modules) == modules and len(module.layers.layers) == convs
        assert module.cmb.post.prefix == f'discriminators.{modules}.conv_post'
        kinds = [type(m).__name__ for m in module.modules]
        assert kinds == ['Period'] * 5 + ['Scale'] * ('multi_scale' in flags) + \
            ['Resolution'] * (3 * ('multi_resolution' in flags))
        for layer in module.layers.layers:
            assert layer.w is not None and layer.gw is not None
            assert (layer.packed is not None) == (math == 'tf32')
        restored = module.state_dict()
        assert all(torch.equal(restored[k], state[k]) for k in state)
    resolutions = [m for m in module.modules if isinstance(m, discriminator.Resolution)]
    assert [(m.n_fft, m.hop, m.win) for m in resolutions] == (
        [(1024, 120, 600), (2048, 240, 1200), (512, 50, 240)] if resolutions else [])


def test_trainer_wiring_for_every_discriminator_flag(monkeypatch):
    """Trainer construction on the CPU (CUDA check bypassed, no kernel launched): the flags reach
    the discriminator, the default stays 5 x period + complex multi-band"""
    from promonet_b200.train import Trainer
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    names = lambda trainer: [type(m).__name__ for m in trainer.discriminators.modules]
    assert names(Trainer(device='cpu')) == ['Period'] * 5
    flagged = Trainer(
        device='cpu', math='fp32', multi_scale_discriminator=True,
        multi_resolution_discriminator=True, spectral_convergence_loss=False)
    assert names(flagged) == ['Period'] * 5 + ['Scale'] + ['Resolution'] * 3
    assert flagged.world == 1 and flagged.generator.params.peers is None
