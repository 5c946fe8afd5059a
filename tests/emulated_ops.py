"""A CPU test double of promonet_b200.train.ops (TEST INFRASTRUCTURE ONLY, like oracle/).

`install(monkeypatch)` replaces the functions of promonet_b200.train.ops that launch kernels
of libpromonet_b200 by plain-torch functions with the semantics include/promonet_b200.h
documents for them (math='fp32' path, ungrouped layers).  It exists to run the HOST logic of the
trainable modules — which launches they make, on which views, with which flags, in which order —
without a GPU: tests/test_train_emulated.py first holds the double to the oracle through the
multi-period and complex multi-band discriminators, then uses it on the other modules and on
the whole Trainer.step (new sequencing can be developed against it before GPU time is spent).
It says nothing about the CUDA kernels themselves and nothing under promonet_b200/ may import it."""
import torch
import torch.nn.functional as F

from promonet_b200.train import ops

ACT_NONE, ACT_LRELU, ACT_LRELU_MASK, ACT_TANH_MASK = 0, 1, 2, 3
OUT_NONE, OUT_LRELU, OUT_TANH = 0, 1, 2
_tables = {}     # data_ptr of a weight table -> its entries (tensors)
_audio = {}      # data_ptr of a spectrum -> the audio it came from (for the STFT backward)


def _act(value, companion, act, slope):
    if act == ACT_NONE:
        return value
    if act == ACT_LRELU:
        return F.leaky_relu(value, slope)
    if act == ACT_LRELU_MASK:
        return value * torch.where(companion.reshape(value.shape) > 0, 1., slope)
    if act == ACT_TANH_MASK:
        return value * (1. - companion.reshape(value.shape) ** 2)
    raise ValueError(act)


def _input_view(g, a):
    """a as (B, c_in, h_in, w_in), honouring the optional element strides of the geometry"""
    if g.channel_stride or g.position_stride or g.batch_stride:
        return a.as_strided(
            (g.batch, g.c_in, g.h_in, g.w_in),
            (g.batch_stride, g.channel_stride, g.w_in * g.position_stride, g.position_stride))
    return a.reshape(g.batch, g.c_in, g.h_in, g.w_in)


def _store(out, value, accumulate):
    flat = out.view(-1)
    if flat.numel() != value.numel():
        raise ValueError(f'output holds {flat.numel()} values, the operator made {value.numel()}')
    if accumulate:
        flat.add_(value.reshape(-1))
    else:
        flat.copy_(value.reshape(-1))
    return out


def conv_gemm(g, transposed, a, wmat, out, a_companion=None, a_act=ACT_NONE, a_slope=1.,
              bias=None, bias2=None, out_act=OUT_NONE, out_slope=1., mask_src=None,
              mask_slope=1., residual=None, alpha=1., accumulate=False):
    stride, dilation, padding = (g.sh, g.sw), (g.dh, g.dw), (g.ph, g.pw)
    if not transposed:
        x = _act(_input_view(g, a), a_companion, a_act, a_slope)
        value = F.conv2d(
            x, wmat.reshape(g.c_out, g.c_in, g.kh, g.kw), None, stride, padding, dilation)
        rows = g.c_out
        assert tuple(value.shape[2:]) == (g.h_out, g.w_out)
    else:
        dy = _act(a.reshape(g.batch, g.c_out, g.h_out, g.w_out), a_companion, a_act, a_slope)
        weight = wmat.reshape(g.c_in, g.c_out, g.kh, g.kw).transpose(0, 1)
        value = torch.nn.grad.conv2d_input(
            (g.batch, g.c_in, g.h_in, g.w_in), weight, dy, stride, padding, dilation)
        rows = g.c_in
    if bias is not None:
        value = value + bias.reshape(1, rows, 1, 1)
    if bias2 is not None:
        value = value + bias2.reshape(g.batch, rows, 1, 1)
    if out_act == OUT_LRELU:
        value = F.leaky_relu(value, out_slope)
    elif out_act == OUT_TANH:
        value = torch.tanh(value)
    if mask_src is not None:
        value = value * torch.where(mask_src.reshape(value.shape) > 0, 1., mask_slope)
    if residual is not None:
        value = value + residual.reshape(value.shape)
    return _store(out, alpha * value, accumulate)


def conv_wgrad(g, dy, x, gw, gbias=None, dy_companion=None, dy_act=ACT_NONE, dy_slope=1.,
               x_companion=None, x_act=ACT_NONE, x_slope=1.):
    dy = _act(dy.reshape(g.batch, g.c_out, g.h_out, g.w_out), dy_companion, dy_act, dy_slope)
    x = _act(_input_view(g, x), x_companion, x_act, x_slope)
    value = torch.nn.grad.conv2d_weight(
        x, (g.c_out, g.c_in, g.kh, g.kw), dy, (g.sh, g.sw), (g.ph, g.pw), (g.dh, g.dw))
    _store(gw, value, True)
    if gbias is not None:
        gbias.add_(dy.sum(dim=(0, 2, 3)))


def weight_table(entries, device):
    table = torch.zeros(max(1, len(entries)), dtype=torch.uint8)
    _tables[table.data_ptr()] = (table, entries)      # keeps the table alive, hence the key unique
    return table


def prepare_weights(table, layers, max_dim0):
    _, entries = _tables[table.data_ptr()]
    assert len(entries) == layers
    for entry in entries:
        dim0, dim1, taps, groups = entry['dim0'], entry['dim1'], entry['taps'], entry.get('groups', 1)
        assert dim0 <= max_dim0
        weight = entry['v'].reshape(dim0, -1)
        if entry['g'] is not None:
            weight = entry['g'].reshape(dim0, 1) * weight / weight.norm(dim=1, keepdim=True)
            entry['w'].view(-1).copy_(weight.reshape(-1))
        if groups > 1:
            # block-diagonal dense form of a grouped weight (dim0, dim1 / groups, taps)
            rows, cols = dim0 // groups, dim1 // groups
            dense = torch.zeros(dim0, dim1, taps)
            for group in range(groups):
                dense[group * rows:(group + 1) * rows, group * cols:(group + 1) * cols] = \
                    weight.reshape(dim0, cols, taps)[group * rows:(group + 1) * rows]
            weight = dense
            entry['dense'].view(-1).copy_(dense.reshape(-1))
        if entry.get('wt') is not None:
            transpose_weight(weight, entry['wt'], dim0, dim1, taps)
        assert entry.get('packed') is None and entry.get('packed_t') is None    # fp32 path only


def extract_grouped(dense, gw, dim0, dim1, taps, groups):
    rows, cols = dim0 // groups, dim1 // groups
    dense = dense.reshape(dim0, dim1, taps)
    blocks = [
        dense[group * rows:(group + 1) * rows, group * cols:(group + 1) * cols]
        for group in range(groups)]
    return _store(gw, torch.cat(blocks), False)


def transpose_weight(w, wt, dim0, dim1, taps):
    return _store(wt, w.reshape(dim0, dim1, taps).transpose(0, 1), False)


def weight_norm_backward(v, g, gw, gv, gg, dim0, inner):
    v, gw = v.reshape(dim0, inner), gw.reshape(dim0, inner)
    norm = v.norm(dim=1, keepdim=True)
    projection = (gw * v).sum(dim=1, keepdim=True)
    _store(gg, projection / norm, False)
    _store(gv, g.reshape(dim0, 1) / norm * (gw - v * projection / norm ** 2), False)


def weight_norm_table(entries, device):
    table = torch.zeros(max(1, len(entries)), dtype=torch.uint8)
    _tables[table.data_ptr()] = (table, entries)
    return table


def weight_norm_backward_table(table, layers, max_dim0):
    _, entries = _tables[table.data_ptr()]
    assert len(entries) == layers
    for entry in entries:
        assert entry['dim0'] <= max_dim0
        weight_norm_backward(
            entry['v'], entry['g'], entry['gw'], entry['gv'], entry['gg'], entry['dim0'], entry['inner'])


def reflect_pad(x, left, right):
    t = x.shape[-1]
    return F.pad(x.reshape(1, -1, t), (left, right), mode='reflect').reshape(
        *x.shape[:-1], left + t + right)


def reflect_pad_backward(gout, gx, left, right, accumulate=False):
    t = gx.shape[-1]
    leaf = torch.zeros(1, gx.numel() // t, t, requires_grad=True)
    F.pad(leaf, (left, right), mode='reflect').backward(gout.reshape(1, -1, left + t + right))
    return _store(gx, leaf.grad, accumulate)


def axpby(a, x, b, y):
    y.copy_((a * x.reshape(y.shape) if x is not None else 0.) + (b * y if b != 0. else 0.))
    return y


def mse_to_target(x, target, weight, loss, grad=None):
    if loss is not None:
        loss.add_(weight * ((x - target) ** 2).mean())
    if grad is not None:
        _store(grad, weight * 2. * (x - target) / x.numel(), False)


def l1_mean(fake, real, weight, loss, gfake=None, accumulate=False):
    if loss is not None:
        loss.add_(weight * (fake - real).abs().mean())
    if gfake is not None:
        _store(gfake, weight * torch.sign(fake - real) / fake.numel(), accumulate)


def _stft_magnitude(audio, window, eps):
    """reflect pad 384, 1024-point STFT at hop 256 (rectangular or periodic hann window) ->
    (B, 513, F) magnitude sqrt(re^2 + im^2 + eps) and the complex spectrum"""
    padded = F.pad(audio[:, None], (384, 384), mode='reflect')[:, 0]
    spectrum = torch.stft(
        padded, 1024, 256, 1024, torch.hann_window(1024) if window == 'hann' else torch.ones(1024),
        center=False, return_complex=True)
    real = torch.view_as_real(spectrum)
    return torch.sqrt(real[..., 0] ** 2 + real[..., 1] ** 2 + eps), spectrum


def _banded(magnitude):
    """layout 2: the five bands of (B, F, 513) one after the other, each contiguous"""
    from promonet_b200 import config
    rows = magnitude.transpose(1, 2)
    return torch.cat([rows[..., lo:hi].reshape(-1) for lo, hi in config.CMB_BANDS])


def stft_magnitude(audio, window='hann', eps=1e-6, layout=0, want_spectrum=True):
    magnitude, spectrum = _stft_magnitude(audio, window, eps)
    spectrum = torch.view_as_real(spectrum.transpose(1, 2)).contiguous()       # (B, F, 513, 2)
    _audio[spectrum.data_ptr()] = (spectrum, audio.clone())
    result = {0: magnitude, 1: magnitude.transpose(1, 2).contiguous(), 2: _banded(magnitude)}[layout]
    return result.contiguous(), spectrum if want_spectrum else None


def stft_magnitude_backward(gmagnitude, spectrum, gaudio, window='hann', eps=1e-6, layout=0,
                            accumulate=False):
    base = spectrum
    while base._base is not None:          # a batch slice of the forward's spectrum
        base = base._base
    whole, audio = _audio[base.data_ptr()]
    first = (spectrum.data_ptr() - whole.data_ptr()) // (spectrum[0].numel() * 4)
    leaf = audio[first:first + spectrum.shape[0]].clone().requires_grad_()
    magnitude, _ = _stft_magnitude(leaf, window, eps)
    result = {0: magnitude, 1: magnitude.transpose(1, 2), 2: _banded(magnitude)}[layout]
    result.backward(gmagnitude.reshape(result.shape))
    return _store(gaudio, leaf.grad, accumulate)


def copy_columns(src, dst, src_offset, dst_offset, cols, accumulate=False):
    source = src[..., src_offset:src_offset + cols].reshape(-1, cols)
    target = dst.view(-1, dst.shape[-1])[:, dst_offset:dst_offset + cols]
    if accumulate:
        target.add_(source)
    else:
        target.copy_(source)
    return dst


def dft_basis_rect(n_fft, win_length, device):
    bins = n_fft // 2 + 1
    window = torch.zeros(n_fft, dtype=torch.float64)
    left = (n_fft - win_length) // 2
    window[left:left + win_length] = 1.
    k = torch.arange(bins, dtype=torch.float64)[:, None]
    phase = 2 * torch.pi * ((k * torch.arange(n_fft, dtype=torch.float64)[None]) % n_fft) / n_fft
    return torch.cat([window * torch.cos(phase), -window * torch.sin(phase)]).float()


def dft_basis(n_fft, device):
    """pmn_dft_basis: periodic hann window over the whole frame"""
    bins = n_fft // 2 + 1
    window = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    k = torch.arange(bins, dtype=torch.float64)[:, None]
    phase = 2 * torch.pi * ((k * torch.arange(n_fft, dtype=torch.float64)[None]) % n_fft) / n_fft
    return torch.cat([window * torch.cos(phase), -window * torch.sin(phase)]).float()


def spectral_convergence(spec, batch, weight, sums, loss, gspec=None):
    """pmn_spectral_convergence (promonet/train/loss.py:61-122)"""
    bins = spec.shape[1] // 2
    root = lambda z: torch.sqrt(torch.clamp(torch.sqrt(z[:, :bins] ** 2 + z[:, bins:] ** 2), min=1e-7))
    leaf = spec[batch:].detach().clone().requires_grad_()
    target = root(spec[:batch])
    difference, reference = (target - root(leaf)).abs().sum(), target.sum()
    sums.copy_(torch.stack([difference.detach(), reference]))
    if loss is not None:
        loss.add_(weight * (difference / reference).detach())
    if gspec is not None:
        (weight * difference / reference).backward()
        _store(gspec, leaf.grad, False)


def complex_magnitude(spec):
    items, rows, frames = spec.shape
    bins = rows // 2
    return torch.sqrt(spec[:, :bins] ** 2 + spec[:, bins:] ** 2)[:, None].contiguous()


def complex_magnitude_backward(gmagnitude, spec):
    items, rows, frames = spec.shape
    bins = rows // 2
    norm = torch.sqrt(spec[:, :bins] ** 2 + spec[:, bins:] ** 2)
    scale = torch.where(norm > 0, gmagnitude.reshape(norm.shape) / norm.clamp_min(1e-38), 0.)
    return torch.cat([scale * spec[:, :bins], scale * spec[:, bins:]], dim=1).contiguous()


def frame_overlap_add(gframes, gsignal, hop):
    batch, n_fft, frames = gframes.shape
    for frame in range(frames):
        gsignal[:, frame * hop:frame * hop + n_fft] += gframes[:, :, frame]
    return gsignal


def conv_transpose1d(x, weight, bias, stride, in_slope):
    k = weight.shape[-1]
    return F.conv_transpose1d(
        F.leaky_relu(x, in_slope), weight, bias, stride=stride, padding=(k - stride) // 2)


def features(loudness, pitch, periodicity, ppg, pitch_distribution, pitch_embedding, threshold):
    from oracle import features as oracle_features
    state = {
        'ppg_threshold': torch.tensor(threshold), 'pitch_distribution': pitch_distribution,
        'pitch_embedding.weight': pitch_embedding}
    return oracle_features.prepare_features(state, loudness, pitch, periodicity, ppg).contiguous()


def pitch_bins(pitch, edges, fmin, fmax):
    return torch.clip(torch.searchsorted(edges, torch.clip(pitch, fmin, fmax)), 0, edges.numel() - 1)


def global_features(speaker_embedding, speakers, sbr, lr):
    return torch.cat([speaker_embedding[speakers], sbr[:, None], lr[:, None]], dim=1)


def embedding_backward(gout, index, gtable, channel_offset=0):
    channels = gtable.shape[1]
    rows = gout[:, channel_offset:channel_offset + channels].permute(0, 2, 1).reshape(-1, channels)
    gtable.index_put_((index.reshape(-1),), rows, accumulate=True)


def row_sum(x, out, rows, cols, accumulate=False):
    return _store(out, x.reshape(rows, cols).sum(dim=1), accumulate)


def channel_sum(x, out, accumulate=False):
    batch, channels = x.shape[:2]
    return _store(out, x.reshape(batch, channels, -1).sum(dim=(0, 2)), accumulate)


def _mel_basis():
    from oracle import dsp
    return torch.from_numpy(dsp.mel_basis()).float()


def linear_to_mel(magnitude, floor=float('-inf')):
    return torch.log(_mel_basis() @ magnitude).clamp_min(floor)


def mel_loss(magnitude, target_mels, weight, loss, gmagnitude=None, grad_weight=None):
    leaf = magnitude.detach().clone().requires_grad_()
    value = (torch.log(_mel_basis() @ leaf) - target_mels).abs().mean()
    if loss is not None:
        loss.add_(weight * value.detach())
    if gmagnitude is not None:
        value.backward()
        _store(gmagnitude, (weight if grad_weight is None else grad_weight) * leaf.grad, False)


EMULATED = (
    conv_transpose1d, features, pitch_bins, global_features, embedding_backward, row_sum,
    channel_sum, linear_to_mel, mel_loss, extract_grouped, dft_basis, spectral_convergence,
    conv_gemm, conv_wgrad, weight_table, prepare_weights, transpose_weight, weight_norm_backward,
    weight_norm_table, weight_norm_backward_table,
    reflect_pad, reflect_pad_backward, axpby, mse_to_target, l1_mean, stft_magnitude,
    stft_magnitude_backward, copy_columns, dft_basis_rect, complex_magnitude,
    complex_magnitude_backward, frame_overlap_add)


def install(monkeypatch):
    """Route promonet_b200.train.ops through the double and let the modules build on the CPU"""
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    for function in EMULATED:
        monkeypatch.setattr(ops, function.__name__, function)

    def refuse(*args, **kwargs):
        raise AssertionError('a tensor-core operator was called on the fp32 path')
    for name in ('conv_gemm_tc', 'conv_wgrad_tc', 'pack_weight_taps'):
        monkeypatch.setattr(ops, name, refuse)
