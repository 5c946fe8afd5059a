"""Tensor-level wrappers of the training-step operators of libpromonet_b200
(include/promonet_b200.h, "Training step").  Every function launches CUDA
kernels on the current stream; torch only owns the memory."""
import ctypes

import torch

from promonet_b200 import _lib
from promonet_b200._lib import ConvGeometry

ACT_NONE, ACT_LRELU, ACT_LRELU_MASK, ACT_TANH_MASK = 0, 1, 2, 3
OUT_NONE, OUT_LRELU, OUT_TANH = 0, 1, 2


def _pair(value):
    return tuple(value) if isinstance(value, (tuple, list)) else (value, value)


def geometry(batch, c_in, c_out, size_in, kernel, stride=1, dilation=1, padding=0,
             size_out=None, strides=(0, 0, 0)):
    """pmn_conv_geometry of a forward convolution; sizes are (H, W)"""
    (h_in, w_in), (kh, kw) = _pair(size_in), _pair(kernel)
    (sh, sw), (dh, dw), (ph, pw) = _pair(stride), _pair(dilation), _pair(padding)
    if size_out is None:
        h_out = (h_in + 2 * ph - dh * (kh - 1) - 1) // sh + 1
        w_out = (w_in + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    else:
        h_out, w_out = _pair(size_out)
    return ConvGeometry(
        batch, c_in, c_out, h_in, w_in, h_out, w_out, kh, kw, sh, sw, dh, dw, ph, pw, *strides)


def _check(status):
    _lib.check(status)


def conv_gemm(geom, transposed, a, wmat, out, a_companion=None, a_act=ACT_NONE, a_slope=1.,
              bias=None, bias2=None, out_act=OUT_NONE, out_slope=1., mask_src=None,
              mask_slope=1., residual=None, alpha=1., accumulate=False):
    _check(_lib.library().pmn_conv_gemm(
        ctypes.byref(geom), int(transposed), _lib.ptr(a), _lib.ptr(a_companion), a_act, a_slope,
        _lib.ptr(wmat), _lib.ptr(bias), _lib.ptr(bias2), out_act, out_slope,
        _lib.ptr(mask_src), mask_slope, _lib.ptr(residual), alpha, int(accumulate),
        _lib.ptr(out), _lib.stream()))
    return out


def conv_gemm_tc(geom, transposed, a, wpacked, out, a_companion=None, a_act=ACT_NONE, a_slope=1.,
                 bias=None, bias2=None, out_act=OUT_NONE, out_slope=1., mask_src=None,
                 mask_slope=1., residual=None, alpha=1., accumulate=False):
    """conv_gemm on tcgen05 (tf32); wpacked from pack_weight_taps"""
    _check(_lib.library().pmn_conv_gemm_tc(
        ctypes.byref(geom), int(transposed), _lib.ptr(a), _lib.ptr(a_companion), a_act, a_slope,
        _lib.ptr(wpacked), _lib.ptr(bias), _lib.ptr(bias2), out_act, out_slope,
        _lib.ptr(mask_src), mask_slope, _lib.ptr(residual), alpha, int(accumulate),
        _lib.ptr(out), _lib.stream()))
    return out


def channel_pad(channels):
    return (channels + 31) // 32 * 32


def packed_floats(rows, reduce, taps):
    """Size of pack_weight_taps' output for a GEMM with `rows` output channels"""
    return int(_lib.library().pmn_conv_tc_packed_floats(rows, reduce, taps))


def pack_weight_taps(w, out, d0, d1, taps, transposed):
    _check(_lib.library().pmn_pack_weight_taps(
        _lib.ptr(w), _lib.ptr(out), d0, d1, taps, int(transposed), _lib.stream()))
    return out


def conv_wgrad(geom, dy, x, gw, gbias=None, dy_companion=None, dy_act=ACT_NONE, dy_slope=1.,
               x_companion=None, x_act=ACT_NONE, x_slope=1.):
    _check(_lib.library().pmn_conv_wgrad(
        ctypes.byref(geom), _lib.ptr(dy), _lib.ptr(dy_companion), dy_act, dy_slope,
        _lib.ptr(x), _lib.ptr(x_companion), x_act, x_slope, _lib.ptr(gw), _lib.ptr(gbias),
        _lib.stream()))


def conv_wgrad_tc(geom, dy, x, gw, gbias=None, dy_companion=None, dy_act=ACT_NONE, dy_slope=1.,
                  x_companion=None, x_act=ACT_NONE, x_slope=1.):
    _check(_lib.library().pmn_conv_wgrad_tc(
        ctypes.byref(geom), _lib.ptr(dy), _lib.ptr(dy_companion), dy_act, dy_slope,
        _lib.ptr(x), _lib.ptr(x_companion), x_act, x_slope, _lib.ptr(gw), _lib.ptr(gbias),
        _lib.stream()))


def weight_table(entries, device):
    """Device copy of a pmn_weight_desc array; entries are dicts of tensors / ints"""
    table = (_lib.WeightDesc * len(entries))()
    for desc, entry in zip(table, entries):
        for name in ('v', 'g', 'w', 'packed', 'packed_t', 'wt', 'dense'):
            tensor = entry.get(name)
            setattr(desc, name, _lib.ptr(tensor) if tensor is not None else None)
        desc.dim0, desc.dim1, desc.taps = entry['dim0'], entry['dim1'], entry['taps']
        desc.groups = entry.get('groups', 1)
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).clone()
    return raw.to(device)


def prepare_weights(table, layers, max_dim0):
    _check(_lib.library().pmn_prepare_weights(_lib.ptr(table), layers, max_dim0, _lib.stream()))


def extract_grouped(dense, gw, dim0, dim1, taps, groups):
    _check(_lib.library().pmn_extract_grouped(
        _lib.ptr(dense), _lib.ptr(gw), dim0, dim1, taps, groups, _lib.stream()))
    return gw


def transpose_weight(w, wt, dim0, dim1, taps):
    _check(_lib.library().pmn_transpose_weight(
        _lib.ptr(w), _lib.ptr(wt), dim0, dim1, taps, _lib.stream()))
    return wt


def weight_norm_fold(v, g, w, dim0, inner):
    _check(_lib.library().pmn_weight_norm_fold(
        _lib.ptr(v), _lib.ptr(g), _lib.ptr(w), dim0, inner, _lib.stream()))
    return w


def weight_norm_table(entries, device):
    """Device copy of a pmn_weight_norm_desc array; entries are dicts of tensors / ints"""
    table = (_lib.WeightNormDesc * len(entries))()
    for desc, entry in zip(table, entries):
        for name in ('v', 'g', 'gw', 'gv', 'gg'):
            setattr(desc, name, _lib.ptr(entry[name]))
        desc.dim0, desc.inner = entry['dim0'], entry['inner']
    raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).clone()
    return raw.to(device)


def weight_norm_backward_table(table, layers, max_dim0):
    _check(_lib.library().pmn_weight_norm_backward_table(_lib.ptr(table), layers, max_dim0, _lib.stream()))


def weight_norm_backward(v, g, gw, gv, gg, dim0, inner):
    _check(_lib.library().pmn_weight_norm_backward(
        _lib.ptr(v), _lib.ptr(g), _lib.ptr(gw), _lib.ptr(gv), _lib.ptr(gg), dim0, inner,
        _lib.stream()))


def reflect_pad(x, left, right):
    """x (..., T) -> (..., left + T + right)"""
    t_in = x.shape[-1]
    rows = x.numel() // t_in
    out = torch.empty(*x.shape[:-1], left + t_in + right, device=x.device, dtype=x.dtype)
    _check(_lib.library().pmn_reflect_pad(
        _lib.ptr(x), _lib.ptr(out), rows, t_in, left, right, _lib.stream()))
    return out


def reflect_pad_backward(gout, gx, left, right, accumulate=False):
    t_in = gx.shape[-1]
    rows = gx.numel() // t_in
    _check(_lib.library().pmn_reflect_pad_backward(
        _lib.ptr(gout), _lib.ptr(gx), rows, t_in, left, right, int(accumulate), _lib.stream()))
    return gx


def axpby(a, x, b, y):
    _check(_lib.library().pmn_axpby(a, _lib.ptr(x), b, _lib.ptr(y), y.numel(), _lib.stream()))
    return y


def mse_to_target(x, target, weight, loss, grad=None):
    _check(_lib.library().pmn_mse_to_target(
        _lib.ptr(x), x.numel(), target, weight, _lib.ptr(loss), _lib.ptr(grad), _lib.stream()))


def l1_mean(fake, real, weight, loss, gfake=None, accumulate=False):
    _check(_lib.library().pmn_l1_mean(
        _lib.ptr(fake), _lib.ptr(real), fake.numel(), weight, _lib.ptr(loss), _lib.ptr(gfake),
        int(accumulate), _lib.stream()))


def adamw(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step, grad_scale=1.,
          step_device=None):
    _check(_lib.library().pmn_adamw(
        _lib.ptr(param), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), param.numel(),
        lr, betas[0], betas[1], eps, weight_decay, step, grad_scale, _lib.ptr(step_device),
        _lib.stream()))


def adamw_peer(grad_ptrs, param_ptrs, rank, exp_avg, exp_avg_sq, begin, end, lr, betas, eps,
               weight_decay, step, step_device=None):
    """Fused reduce-scatter + AdamW + all-gather over peer memory (pmn_adamw_peer); grad_ptrs /
    param_ptrs are the per-rank device addresses (ints) of the flat buffers"""
    world = len(grad_ptrs)
    grads = (ctypes.c_void_p * world)(*grad_ptrs)
    params = (ctypes.c_void_p * world)(*param_ptrs)
    _check(_lib.library().pmn_adamw_peer(
        grads, params, world, rank, _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), begin, end, lr,
        betas[0], betas[1], eps, weight_decay, step, _lib.ptr(step_device), _lib.stream()))


def row_sum(x, out, rows, cols, accumulate=False):
    _check(_lib.library().pmn_row_sum(
        _lib.ptr(x), _lib.ptr(out), rows, cols, int(accumulate), _lib.stream()))
    return out


def features(loudness, pitch, periodicity, ppg, pitch_distribution, pitch_embedding, threshold):
    batch, rows, frames = loudness.shape
    out = torch.empty(batch, 113, frames, device=loudness.device)
    _check(_lib.library().pmn_features(
        _lib.ptr(loudness), rows, _lib.ptr(pitch), _lib.ptr(periodicity), _lib.ptr(ppg),
        _lib.ptr(pitch_distribution), _lib.ptr(pitch_embedding), threshold, _lib.ptr(out),
        batch, frames, _lib.stream()))
    return out


def pitch_bins(pitch, edges, fmin, fmax):
    bins = torch.empty(pitch.shape, dtype=torch.int64, device=pitch.device)
    _check(_lib.library().pmn_pitch_bins(
        _lib.ptr(pitch), _lib.ptr(edges), _lib.ptr(bins), pitch.numel(), edges.numel(),
        fmin, fmax, _lib.stream()))
    return bins


def embedding_backward(gout, index, gtable, channel_offset=0):
    """gtable[index[b, f]] += gout[b, offset:offset + channels, f]"""
    batch, out_channels, frames = gout.shape
    rows, channels = gtable.shape
    _check(_lib.library().pmn_embedding_backward(
        _lib.ptr(gout), _lib.ptr(index), _lib.ptr(gtable), batch, channels, frames, rows,
        out_channels, channel_offset, _lib.stream()))


def global_features(speaker_embedding, speakers, sbr, lr):
    num_speakers, channels = speaker_embedding.shape
    out = torch.empty(speakers.shape[0], channels + 2, device=speaker_embedding.device)
    _check(_lib.library().pmn_global_features(
        _lib.ptr(speaker_embedding), _lib.ptr(speakers), _lib.ptr(sbr), _lib.ptr(lr),
        _lib.ptr(out), speakers.shape[0], channels, num_speakers, _lib.stream()))
    return out


def stft_magnitude(audio, window='hann', eps=1e-6, layout=0, want_spectrum=True):
    """audio (B, T) -> magnitude ((B, 513, F) or (B, F, 513)), spectrum (B, F, 513, 2)"""
    batch, samples = audio.shape
    frames = samples // 256
    shape = {0: (batch, 513, frames), 1: (batch, frames, 513), 2: (batch * frames * 513,)}[layout]
    magnitude = torch.empty(shape, device=audio.device)
    spectrum = torch.empty(batch, frames, 513, 2, device=audio.device) if want_spectrum else None
    _check(_lib.library().pmn_stft_magnitude(
        _lib.ptr(audio), batch, samples, 0 if window == 'hann' else 1, eps, layout,
        _lib.ptr(spectrum), _lib.ptr(magnitude), _lib.stream()))
    return magnitude, spectrum


def stft_magnitude_backward(gmagnitude, spectrum, gaudio, window='hann', eps=1e-6, layout=0,
                            accumulate=False):
    batch, samples = gaudio.shape
    _check(_lib.library().pmn_stft_magnitude_backward(
        _lib.ptr(gmagnitude), _lib.ptr(spectrum), batch, samples,
        0 if window == 'hann' else 1, eps, layout, _lib.ptr(gaudio), int(accumulate),
        _lib.stream()))
    return gaudio


def mel_loss(magnitude, target_mels, weight, loss, gmagnitude=None, grad_weight=None):
    batch, _, frames = magnitude.shape
    _check(_lib.library().pmn_mel_loss(
        _lib.ptr(magnitude), _lib.ptr(target_mels), batch, frames, weight,
        weight if grad_weight is None else grad_weight, _lib.ptr(loss),
        _lib.ptr(gmagnitude), _lib.stream()))


def channel_sum(x, out, accumulate=False):
    """out[c] (+)= sum over batch and trailing axes of x (B, C, ...)"""
    batch, channels = x.shape[:2]
    _check(_lib.library().pmn_channel_sum(
        _lib.ptr(x), _lib.ptr(out), batch, channels, x.numel() // (batch * channels),
        int(accumulate), _lib.stream()))
    return out


def copy_columns(src, dst, src_offset, dst_offset, cols, accumulate=False):
    """dst[..., dst_offset:dst_offset + cols] (+)= src[..., src_offset:src_offset + cols]"""
    rows = src.numel() // src.shape[-1]
    _check(_lib.library().pmn_copy_columns(
        _lib.ptr(src), src.shape[-1], src_offset, _lib.ptr(dst), dst.shape[-1], dst_offset,
        rows, cols, int(accumulate), _lib.stream()))
    return dst


def linear_to_mel(magnitude, floor=float('-inf')):
    batch, _, frames = magnitude.shape
    mels = torch.empty(batch, 80, frames, device=magnitude.device)
    _check(_lib.library().pmn_linear_to_mel(
        _lib.ptr(magnitude), _lib.ptr(mels), floor, batch, frames, _lib.stream()))
    return mels


def conv_transpose1d(x, weight, bias, stride, in_slope):
    """LeakyReLU + ConvTranspose1d (kernel 2 * stride, padding stride / 2), hifigan.py:97-106"""
    batch, c_in, t_in = x.shape
    _, c_out, k = weight.shape
    out = torch.empty(batch, c_out, t_in * stride, device=x.device)
    _check(_lib.library().pmn_conv_transpose1d(
        _lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(out), batch, c_in, c_out, t_in,
        k, stride, in_slope, _lib.stream()))
    return out


def dft_basis(n_fft, device):
    """(2 bins, n_fft) windowed DFT weight (loss.py:61-80 as a 1 x 1 convolution over frames)"""
    out = torch.empty(2 * (n_fft // 2 + 1), n_fft, device=device)
    _check(_lib.library().pmn_dft_basis(_lib.ptr(out), n_fft, _lib.stream()))
    return out


def spectral_convergence(spec, batch, weight, sums, loss, gspec=None):
    """spec (2 B, 2 bins, frames): target items first; see pmn_spectral_convergence"""
    _, rows, frames = spec.shape
    _check(_lib.library().pmn_spectral_convergence(
        _lib.ptr(spec), batch, rows // 2, frames, weight, _lib.ptr(sums), _lib.ptr(loss),
        _lib.ptr(gspec), _lib.stream()))


def frame_overlap_add(gframes, gsignal, hop):
    batch, n_fft, frames = gframes.shape
    _check(_lib.library().pmn_frame_overlap_add(
        _lib.ptr(gframes), _lib.ptr(gsignal), batch, n_fft, frames, hop, gsignal.shape[-1],
        _lib.stream()))
    return gsignal


def dft_basis_rect(n_fft, win_length, device):
    """(2 bins, n_fft) DFT weight with a centred rectangular window of win_length samples
    (torch.stft(window=None, win_length=...), model/discriminator.py:134-140)"""
    out = torch.empty(2 * (n_fft // 2 + 1), n_fft, device=device)
    _check(_lib.library().pmn_dft_basis_rect(_lib.ptr(out), n_fft, win_length, _lib.stream()))
    return out


def complex_magnitude(spec):
    """spec (N, 2 bins, frames) -> |X| as (N, 1, bins, frames) (discriminator.py:141)"""
    items, rows, frames = spec.shape
    out = torch.empty(items, 1, rows // 2, frames, device=spec.device)
    _check(_lib.library().pmn_complex_magnitude(
        _lib.ptr(spec), _lib.ptr(out), items, rows // 2, frames, _lib.stream()))
    return out


def complex_magnitude_backward(gmagnitude, spec):
    """-> gradient of spec (N, 2 bins, frames)"""
    items, rows, frames = spec.shape
    gspec = torch.empty_like(spec)
    _check(_lib.library().pmn_complex_magnitude_backward(
        _lib.ptr(gmagnitude), _lib.ptr(spec), _lib.ptr(gspec), items, rows // 2, frames,
        _lib.stream()))
    return gspec
