"""Numerics experiment (CPU, no GPU): could the residual-block convolutions run on the INT8 tensor
cores (tcgen05 kind::i8, twice the bf16 rate on sm_100a) inside the 1e-4 parity bar?

Ozaki-style slicing: every operand is a 16-bit fixed-point number relative to a scale, stored as
two int8 slices x = (x_hi 2^8 + x_lo) s_x; the products x_hi w_hi, x_hi w_lo, x_lo w_hi accumulate
EXACTLY in int32 (8 + 8 bits x at most 256 channels x 11 taps < 2^31) and the dropped x_lo w_lo term
is 2^-16 of the leading one.  Three int8 MMAs = 1.5 bf16 MMAs instead of the 3 of the bf16 x 3
scheme, and the 4 KB A tile of one MMA then holds K = 32 channels instead of 16.
The catch is the scale: an activation row of the GEMM is one time step, and the taps of a
convolution add different time steps into the same accumulator, so the scale cannot vary along
time inside a tile.  Tried here: one scale per utterance and layer (max |x| over channels and time,
which the producing epilogue can deliver with an atomicMax), weights scaled per output channel.

    python profiles/debug/int8_slices_numerics.py [frames]

Prints max|a - b| / max|b| of the generated audio against the fp64 oracle for: exact fp32,
bf16 x 3 (the shipped scheme), int8 x 3 slices as above, the same with one activation scale per
utterance and input channel (which can be folded into a per-utterance copy of the weights), that
with the fourth product kept, and int8 with only one slice of the activations (2 products).
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import hifigan, inputs  # noqa: E402
from promonet_b200.model import init  # noqa: E402

conv1d, conv_transpose1d = F.conv1d, F.conv_transpose1d


def fixed16(value, scale):
    """16-bit signed fixed point relative to scale, returned as (hi, lo) int8-valued slices"""
    q = torch.clamp(torch.round(value / scale * 32767.), -32767., 32767.)
    hi = torch.floor((q + 128.) / 256.)
    return hi, q - 256. * hi          # lo in [-128, 127]


def split_bf16(value):
    hi = value.float().bfloat16().double()
    return hi, (value - hi).float().bfloat16().double()


def make(scheme):
    def operands(x, w, transpose):
        if scheme == 'bf16x3':
            x_hi, x_lo = split_bf16(x)
            w_hi, w_lo = split_bf16(w)
            return ((x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi)), 1.
        # per-utterance scale of the activations, per-output-channel scale of the weights
        if 'channel' in scheme:     # one scale per utterance AND input channel (foldable into the weights)
            sx = x.abs().amax(dim=2, keepdim=True).clamp_min(1e-30)
        else:
            sx = x.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30)
        dims = (0, 2) if transpose else (1, 2)
        sw = w.abs().amax(dim=dims, keepdim=True).clamp_min(1e-30)
        x_hi, x_lo = fixed16(x, sx)
        w_hi, w_lo = fixed16(w, sw)
        unit = 1. / 32767.
        x_hi, x_lo = x_hi * 256. * sx * unit, x_lo * sx * unit
        w_hi, w_lo = w_hi * 256. * sw * unit, w_lo * sw * unit
        if scheme in ('int8x3', 'int8x3-channel'):
            return ((x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi)), 1.
        if scheme == 'int8x4-channel':
            return ((x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi), (x_lo, w_lo)), 1.
        if scheme == 'int8x2':
            return ((x_hi, w_hi), (x_hi, w_lo)), 1.
        raise ValueError(scheme)

    def conv(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
        if w.shape[1] < 32:            # the 113 -> 512 input conv stays fp32 in the product too
            return conv1d(x, w, bias, stride, padding, dilation, groups)
        products, _ = operands(x, w, False)
        out = sum(conv1d(a, b, None, stride, padding, dilation, groups) for a, b in products)
        return out if bias is None else out + bias[None, :, None]

    def conv_transpose(x, w, bias=None, stride=1, padding=0, **kwargs):
        products, _ = operands(x, w, True)
        out = sum(conv_transpose1d(a, b, None, stride, padding, **kwargs) for a, b in products)
        return out if bias is None else out + bias[None, :, None]
    return conv, conv_transpose


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    state = init.hifigan_state(1234)
    args = inputs.synthesis(2, frames, seed=3)
    double = hifigan.to_double(state)
    dargs = [a.double() if a.is_floating_point() else a for a in args]
    with torch.no_grad():
        exact = hifigan.generator(double, *dargs)
        error = lambda audio: float(
            ((audio.double() - exact).abs().amax(dim=(1, 2)) / exact.abs().amax(dim=(1, 2))).max())
        print(f'fp32 (torch CPU)      {error(hifigan.generator(state, *args)):.2e}')
        for scheme in ('bf16x3', 'int8x3', 'int8x3-channel', 'int8x4-channel', 'int8x2'):
            F.conv1d, F.conv_transpose1d = make(scheme)
            try:
                audio = hifigan.generator(double, *dargs)
            finally:
                F.conv1d, F.conv_transpose1d = conv1d, conv_transpose1d
            print(f'{scheme:21s} {error(audio):.2e}')


if __name__ == '__main__':
    main()
