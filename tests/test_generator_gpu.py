"""Parity of the CUDA synthesis path (through the C ABI) against the oracle
and the golden vectors frozen from the reference.  Tolerance: 1e-4 relative to
max|reference| (BASELINE.json north_star), written at each assert."""
import ctypes

import pytest
import torch

from conftest import relative_error
from oracle import features as oracle_features
from oracle import hifigan, inputs

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-4


@pytest.fixture(scope='module')
def lib():
    from promonet_b200 import _lib
    return _lib


@pytest.fixture(scope='module')
def state():
    from promonet_b200.model import init
    return init.hifigan_state(1234)


@pytest.fixture(scope='module', params=['fp32', 'bf16x3', 'bf16x3+f8'])
def model(state, request, lib):
    """Every math mode must meet the same parity bar: fp32 FMAs, bf16 x 3 on the tensor cores, and
    the latter with the C = 128 residual blocks on fp16 + 2 x fp8 operands (the default)"""
    import promonet_b200
    math = lib.MATH_FP32_SIMT if request.param == 'fp32' else lib.MATH_BF16X3_TC
    return promonet_b200.model.Generator(state=state, math=math, f8=request.param.endswith('+f8'))


def conv1d(lib, x, weight, bias=None, bias2=None, residual=None, dilation=1,
           padding=0, in_slope=1., out_act=0, accum=None, accum_mode=0,
           accum_scale=1., want_out=True):
    x = x.cuda().contiguous()
    weight = weight.cuda().contiguous()
    c_out, c_in, k = weight.shape
    packed = torch.empty(c_in, k, c_out, device='cuda')
    lib.check(lib.library().pmn_pack_conv1d_weight(
        weight.data_ptr(), packed.data_ptr(), c_out, c_in, k, lib.stream()))
    batch, _, t_in = x.shape
    t_out = t_in + 2 * padding - (k - 1) * dilation
    out = torch.empty(batch, c_out, t_out, device='cuda') if want_out else None
    keep = [t.cuda().contiguous() if t is not None else None for t in (bias, bias2, residual)]
    lib.check(lib.library().pmn_conv1d(
        x.data_ptr(), packed.data_ptr(), *[lib.ptr(t) for t in keep],
        lib.ptr(out), lib.ptr(accum), accum_mode, accum_scale,
        batch, c_in, c_out, t_in, t_out, k, dilation, padding, in_slope, out_act,
        lib.stream()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('c_in,c_out,k,dilation,t', [
    (32, 32, 3, 1, 700), (32, 32, 11, 5, 1030), (64, 64, 7, 3, 513),
    (128, 128, 11, 1, 300), (256, 256, 3, 5, 129), (113, 512, 7, 1, 86),
    (1, 256, 32, 1, 993), (48, 40, 5, 2, 77), (512, 1440, 4, 1, 4)])
def test_conv1d_matches_torch(lib, c_in, c_out, k, dilation, t):
    torch.manual_seed(c_in + k)
    x = torch.randn(3, c_in, t)
    w = torch.randn(c_out, c_in, k) / (c_in * k) ** .5
    b = torch.randn(c_out)
    same = k % 2 == 1
    padding = dilation * (k - 1) // 2 if same else 0
    expected = torch.nn.functional.conv1d(
        torch.nn.functional.leaky_relu(x.double(), 0.1), w.double(), b.double(),
        padding=padding, dilation=dilation)
    actual = conv1d(lib, x, w, b, dilation=dilation, padding=padding, in_slope=0.1)
    assert relative_error(actual, expected) < 1e-5


def test_conv1d_epilogue(lib):
    torch.manual_seed(0)
    x = torch.randn(2, 64, 333)
    w = torch.randn(64, 64, 7) / 21.
    b, b2, r = torch.randn(64), torch.randn(2, 64), torch.randn(2, 64, 333)
    y = torch.nn.functional.conv1d(
        x.double(), w.double(), b.double(), padding=3) + b2[:, :, None] + r
    accum = torch.ones(2, 64, 333, device='cuda')
    out = conv1d(lib, x, w, b, b2, r, padding=3, out_act=1, accum=accum,
                 accum_mode=2, accum_scale=0.5)
    assert relative_error(out, torch.tanh(y)) < 1e-5
    assert relative_error(accum, 1. + 0.5 * torch.tanh(y)) < 1e-5
    accum2 = torch.zeros(2, 64, 333, device='cuda')
    assert conv1d(lib, x, w, b, padding=3, accum=accum2, accum_mode=1,
                  accum_scale=1 / 3, want_out=False) is None
    assert relative_error(accum2, torch.nn.functional.conv1d(
        x.double(), w.double(), b.double(), padding=3) / 3) < 1e-5


def test_conv1d_rejects_bad_arguments(lib):
    x = torch.zeros(1, 4, 8, device='cuda')
    status = lib.library().pmn_conv1d(
        x.data_ptr(), x.data_ptr(), None, None, None, x.data_ptr(), None, 0, 1.,
        1, 4, 4, 8, 9, 3, 1, 0, 1., 0, lib.stream())
    assert status == -1 and b't_out' in lib.library().pmn_last_error()


@pytest.mark.parametrize('c_in,c_out,k,s,t', [
    (512, 256, 16, 8, 86), (256, 128, 16, 8, 130), (128, 64, 4, 2, 1000),
    (64, 32, 4, 2, 257), (24, 20, 16, 8, 5)])
def test_conv_transpose1d_matches_torch(lib, c_in, c_out, k, s, t):
    torch.manual_seed(k + t)
    x = torch.randn(2, c_in, t)
    w = torch.randn(c_in, c_out, k) / c_in ** .5
    b = torch.randn(c_out)
    expected = torch.nn.functional.conv_transpose1d(
        torch.nn.functional.leaky_relu(x.double(), 0.1), w.double(), b.double(),
        stride=s, padding=(k - s) // 2)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = torch.empty(2, c_out, s * t, device='cuda')
    lib.check(lib.library().pmn_conv_transpose1d(
        xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(),
        2, c_in, c_out, t, k, s, 0.1, lib.stream()))
    assert relative_error(out, expected) < 1e-5


def test_weight_norm_fold(lib):
    v = torch.randn(37, 5, 11)
    g = torch.rand(37, 1, 1) + .5
    expected = hifigan.fold_weight_norm(g, v)
    vd, gd = v.cuda(), g.cuda()
    out = torch.empty_like(vd)
    lib.check(lib.library().pmn_weight_norm_fold(
        vd.data_ptr(), gd.data_ptr(), out.data_ptr(), 37, 55, lib.stream()))
    assert relative_error(out, expected) < 1e-6


@pytest.mark.parametrize('tag,kernel', [('c32k3', 3), ('c64k11', 11)])
def test_block_matches_reference_golden(lib, golden, tag, kernel):
    """Block.forward hifigan.py:198-210 against the reference's own output"""
    g = golden('block')
    x = g[f'{tag}_x'].cuda()
    cur = x
    for i, dilation in enumerate((1, 3, 5)):
        def folded(name):
            return hifigan.fold_weight_norm(
                g[f'{tag}_{name}.{i}.weight_g'], g[f'{tag}_{name}.{i}.weight_v'])
        xt = conv1d(lib, cur, folded('convs1'), g[f'{tag}_convs1.{i}.bias'],
                    dilation=dilation, padding=dilation * (kernel - 1) // 2, in_slope=0.1)
        cur = conv1d(lib, xt, folded('convs2'), g[f'{tag}_convs2.{i}.bias'],
                     residual=cur, padding=(kernel - 1) // 2, in_slope=0.1)
    assert relative_error(cur, g[f'{tag}_y']) < TOLERANCE


@pytest.mark.parametrize('tag', ['r8', 'r513'])
def test_features_match_reference_golden(model, golden, tag):
    g = golden('generator')
    actual = model.features(*[g[f'{tag}_{k}'] for k in (
        'loudness', 'pitch', 'periodicity', 'ppg')])
    expected = g[f'{tag}_features']
    # the pitch embedding rows are gathered: exact; the rest is fp32 arithmetic
    assert torch.equal(actual[:, 40:104].cpu(), expected[:, 40:104])
    assert relative_error(actual, expected) < 1e-6


@pytest.mark.parametrize('tag', ['r8', 'r513'])
def test_generator_matches_reference_golden(model, golden, tag):
    g = golden('generator')
    args = [g[f'{tag}_{k}'] for k in (
        'loudness', 'pitch', 'periodicity', 'ppg', 'speakers', 'sbr', 'lr')]
    audio = model(*[a.cuda() for a in args])
    assert audio.shape == g[f'{tag}_audio'].shape
    assert relative_error(audio, g[f'{tag}_audio']) < TOLERANCE


@pytest.mark.parametrize('batch,frames', [(1, 86), (3, 37), (2, 1)])
def test_generator_matches_oracle(model, state, batch, frames):
    args = inputs.synthesis(batch, frames, seed=batch * 100 + frames)
    with torch.no_grad():
        expected = hifigan.generator(hifigan.to_double(state), *[
            a.double() if a.is_floating_point() else a for a in args])
    audio = model(*[a.cuda() for a in args])
    error = relative_error(audio, expected)
    assert error < TOLERANCE, error
    # per utterance too
    for i in range(batch):
        assert relative_error(audio[i], expected[i]) < TOLERANCE


def test_batch_items_are_independent(model):
    """Sharding invariant: an utterance's audio does not depend on its batch"""
    args = inputs.synthesis(4, 40, seed=5)
    full = model(*[a.cuda() for a in args])
    half = model(*[a[2:].cuda() for a in args])
    assert torch.equal(full[2:], half)


def test_forward_host_equals_device_forward(model):
    args = inputs.synthesis(2, 33, seed=9)
    device = model(*[a.cuda() for a in args]).cpu()
    host = model.forward_host(*[a.pin_memory() for a in args])
    assert torch.equal(host, device)


def test_stream_host_yields_every_batch_in_order(model):
    """The pipelined host entry (copies on their own streams) returns what the synchronous
    device forward returns, for batches of different shapes and more batches than slots"""
    batches = [
        [a.pin_memory() for a in inputs.synthesis(batch, frames, seed=20 + i)]
        for i, (batch, frames) in enumerate([(2, 33), (2, 33), (1, 50), (3, 17), (2, 33)])]
    expected = [model(*[a.cuda() for a in batch]).cpu() for batch in batches]
    count = 0
    for audio, reference in zip(model.stream_host(iter(batches)), expected):
        assert audio.is_pinned() and torch.equal(audio, reference)
        count += 1
    assert count == len(batches)
    assert list(model.stream_host(iter([]))) == []
    with pytest.raises(ValueError):
        list(model.stream_host(iter(batches), depth=1))


def test_from_features_signature_and_parity(state):
    """promonet.synthesize.from_features (synthesize/core.py:18-59)"""
    import promonet_b200
    loud, pitch, per, ppg, spk, sbr, lr = inputs.synthesis(1, 50, seed=3)
    audio = promonet_b200.synthesize.from_features(
        loud[0], pitch, per, ppg, speaker=int(spk[0]),
        spectral_balance_ratio=float(sbr[0]), loudness_ratio=float(lr[0]), gpu=0)
    assert audio.shape == (1, 50 * 256) and audio.dtype == torch.float32
    with torch.no_grad():
        expected = hifigan.generator(state, loud, pitch, per, ppg, spk, sbr, lr)[0]
    assert relative_error(audio, expected) < TOLERANCE


def test_generator_reports_errors(model, lib):
    with pytest.raises(ValueError):
        model(torch.zeros(1, 8, 4), torch.zeros(1, 5), torch.zeros(1, 4),
              torch.zeros(1, 40, 4), torch.zeros(1, dtype=torch.long),
              torch.ones(1), torch.ones(1))
    handle = ctypes.c_void_p()
    lib.check(lib.library().pmn_generator_create(ctypes.byref(handle)))
    with pytest.raises(lib.Error, match='missing tensor'):
        lib.check(lib.library().pmn_generator_finalize(handle, 0, lib.stream()))
    lib.library().pmn_generator_destroy(handle)


def test_from_files_to_files_batches_equal_lengths(tmp_path):
    """synthesize/core.py:62-201: features on disk -> wav files, utterances of equal length in one
    batch, each identical to its own batch-1 synthesis"""
    import wave
    import promonet_b200
    from promonet_b200 import synthesize
    files = {k: [] for k in ('loudness', 'pitch', 'periodicity', 'ppg', 'output')}
    expected = []
    for index, frames in enumerate((20, 33, 20, 20)):
        loud, pitch, per, ppg, _, _, _ = inputs.synthesis(1, frames, seed=50 + index)
        stretched = torch.softmax(2. * torch.randn(40, frames + 7), dim=-2)   # needs resampling
        for name, value in (('loudness', loud[0]), ('pitch', pitch), ('periodicity', per),
                            ('ppg', stretched)):
            file = tmp_path / f'{index}-{name}.pt'
            torch.save(value, file)
            files[name].append(file)
        files['output'].append(tmp_path / 'out' / f'{index}.wav')
        expected.append(synthesize.from_file(
            files['loudness'][-1], files['pitch'][-1], files['periodicity'][-1], files['ppg'][-1],
            speaker=index))
    synthesize.from_files_to_files(
        files['loudness'], files['pitch'], files['periodicity'], files['ppg'], files['output'],
        speakers=[0, 1, 2, 3])
    for file, audio in zip(files['output'], expected):
        with wave.open(str(file), 'rb') as handle:
            assert handle.getframerate() == promonet_b200.SAMPLE_RATE
            data = torch.frombuffer(bytearray(handle.readframes(handle.getnframes())), dtype=torch.int16)
        assert data.shape[0] == audio.shape[-1]
        quantized = (audio.cpu().reshape(-1).clamp(-1., 1.) * 32767.).round()
        assert (data.float() - quantized).abs().max() <= 1.
