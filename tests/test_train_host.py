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


def test_optimizer_state_interchanges_with_torch_adamw():
    """Checkpoints carry torch.optim.AdamW.state_dict() (torchutil.checkpoint, train/core.py:426-438):
    a reference optimizer's state loads into the flat buffers and ours loads into a reference optimizer"""
    from promonet_b200 import config
    torch.manual_seed(0)
    state = {'a.weight': torch.randn(3, 5), 'a.bias': torch.randn(3), 'b.weight_g': torch.randn(7, 1, 1)}
    params = ParamSet(state, 'cpu')
    leaves = [torch.nn.Parameter(v.clone()) for v in state.values()]
    reference = torch.optim.AdamW(
        leaves, lr=config.LEARNING_RATE, betas=config.ADAM_BETAS, eps=config.ADAM_EPS,
        weight_decay=config.WEIGHT_DECAY)
    for _ in range(3):
        for leaf in leaves:
            leaf.grad = torch.randn_like(leaf)
        reference.step()
    params.load_optimizer_state(reference.state_dict())
    assert params.steps == 3 and float(params.steps_device) == 3.
    for leaf, name in zip(leaves, state):
        assert torch.equal(params._view(params.exp_avg, name), reference.state[leaf]['exp_avg'])
        assert torch.equal(params._view(params.exp_avg_sq, name), reference.state[leaf]['exp_avg_sq'])
    exported = params.optimizer_state()
    assert exported['param_groups'][0]['lr'] == config.LEARNING_RATE
    fresh = torch.optim.AdamW([torch.nn.Parameter(v.clone()) for v in state.values()])
    fresh.load_state_dict(exported)                     # what torchutil.checkpoint.load does
    for position, leaf in enumerate(fresh.param_groups[0]['params']):
        assert torch.equal(fresh.state[leaf]['exp_avg'], reference.state[leaves[position]]['exp_avg'])
        assert float(fresh.state[leaf]['step']) == 3.
    # the flat layout of the first checkpoints still loads; anything else is refused loudly
    params.load_optimizer_state({'exp_avg': params.exp_avg.clone(), 'exp_avg_sq': params.exp_avg_sq.clone(), 'step': 5})
    assert params.steps == 5
    with pytest.raises(ValueError):
        params.load_optimizer_state({'moments': 1})
    with pytest.raises(ValueError):
        params.load_optimizer_state({'state': {}, 'param_groups': [{'params': [0]}]})


def test_pitch_checkpoint_key_mapping():
    """FCNF0++ checkpoints: our names, and upstream penn's Sequential-of-Sequential names
    ([RECALLED]: `N.0.weight` conv, `N.2|3.weight` LayerNorm, `6.weight` head), mapped by
    block index and tensor rank; anything that does not cover the network is refused"""
    from promonet_b200.preprocess import penn
    ours = penn.init_state(3)
    assert all(torch.equal(v, ours[k]) for k, v in penn.convert_state(ours).items())
    upstream = {}
    for i in range(6):
        norm = 3 if i < 3 else 2                        # after Conv1d, ReLU (, MaxPool1d)
        upstream[f'{i}.0.weight'] = ours[f'layers.{i}.conv.weight']
        upstream[f'{i}.0.bias'] = ours[f'layers.{i}.conv.bias']
        upstream[f'{i}.{norm}.weight'] = ours[f'layers.{i}.norm.weight']
        upstream[f'{i}.{norm}.bias'] = ours[f'layers.{i}.norm.bias']
    upstream['6.weight'], upstream['6.bias'] = ours['layers.6.weight'], ours['layers.6.bias']
    for prefix in ('', 'module.'):
        mapped = penn.convert_state({prefix + k: v for k, v in upstream.items()})
        assert list(mapped) == list(ours)
        assert all(torch.equal(mapped[k], ours[k]) for k in ours)
    del upstream['4.2.bias']
    with pytest.raises(ValueError):
        penn.convert_state(upstream)
