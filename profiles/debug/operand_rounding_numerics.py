"""Numerics experiment (CPU, no GPU): which operand roundings keep the generator inside the 1e-4
parity bar, and what would each cost on the tensor cores?  The whole HiFi-GAN generator runs in
fp64 with the operands of every residual-block convolution and transposed convolution rounded the
way a scheme would round them (products and sums exact, as fp32 / int32 accumulation is to this
precision); the error of the generated audio against the fp64 oracle is printed per scheme.

    python profiles/debug/operand_rounding_numerics.py [frames] [scheme ...]      (SEED=n for other inputs)
    python profiles/debug/operand_rounding_numerics.py pitch                      (the pitch network)

Schemes (MMA cost per MAC in units of one bf16 MMA; results in DESIGN.md section 7):
  bf16x3             the shipped one: x_hi w_hi + x_hi w_lo + x_lo w_hi on bf16 halves          3
  fp16+fp8x3         the same three products on fp16 halves                                       3
  fp16+fp8x2         fp16 main term; the two corrections (2^-11 of it) in fp8 e4m3, one
                     power-of-two scale per tensor                                                2
  fp16+mxfp8x2       ... one scale per 32 channels (tcgen05.mma.kind::mxf8f6f4.block_scale)       2
  fp16+fp8x2-static  ... activation scales fixed in advance                                       2
  int8x3             Ozaki-style: operands as two int8 slices of a 16-bit fixed-point number,
                     x = (x_hi 2^8 + x_lo) s; three products, exact in int32 (kind::i8 runs at
                     twice the bf16 rate).  An activation row of the GEMM is one time step and the
                     taps add different time steps into one accumulator, so the scale cannot
                     follow time: one per utterance and layer, weights per output channel          1.5
  int8x3-channel     ... one activation scale per utterance and input channel (foldable into a
                     per-utterance copy of the weights)                                          1.5
  int8x4-channel     ... with the x_lo w_lo product kept                                          2
  int8x2             one int8 slice of the activations                                            1
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


def fp8(value, scale=None):
    """value rounded to e4m3 (3 mantissa bits, largest finite value 448) after a power-of-two
    scaling that puts the tensor's largest magnitude in [128, 256); returned unscaled"""
    if scale is None:
        scale = 2. ** torch.floor(torch.log2(256. / value.abs().max().clamp_min(1e-30)))
    return (value * scale).float().to(torch.float8_e4m3fn).double() / scale


def mxfp8(value, axis):
    """e4m3 with one power-of-two scale per block of 32 consecutive elements along `axis` (the
    reduction axis: channels), as tcgen05.mma.kind::mxf8f6f4.block_scale consumes it"""
    moved = value.transpose(axis, -1)
    shape = moved.shape
    blocks = moved.reshape(*shape[:-1], -1, 32) if shape[-1] % 32 == 0 else moved.reshape(*shape[:-1], 1, -1)
    scale = 2. ** torch.floor(torch.log2(256. / blocks.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)))
    rounded = (blocks * scale).float().to(torch.float8_e4m3fn).double() / scale
    return rounded.reshape(shape).transpose(axis, -1)


def make(scheme):
    def operands(x, w, transpose):
        if scheme == 'fp16+mxfp8x2':
            x_hi, w_hi = x.float().half().double(), w.float().half().double()
            x_lo, w_lo = x - x_hi, w - w_hi
            axis = 0 if transpose else 1          # the input-channel axis of the weight
            return ((x_hi, w_hi), (mxfp8(x_hi, 1), mxfp8(w_lo, axis)),
                    (mxfp8(x_lo, 1), mxfp8(w_hi, axis))), 1.
        if scheme.startswith('fp16+fp8'):
            # main term in fp16 (11-bit operands, one bf16-rate MMA), the two 2^-11 corrections in
            # fp8 e4m3 (half an MMA each): 2 units instead of 3.  Power-of-two scales keep the
            # residuals in e4m3's range.
            x_hi, w_hi = x.float().half().double(), w.float().half().double()
            x_lo, w_lo = x - x_hi, w - w_hi
            if scheme == 'fp16+fp8x2':
                return ((x_hi, w_hi), (fp8(x_hi), fp8(w_lo)), (fp8(x_lo), fp8(w_hi))), 1.
            if scheme == 'fp16+fp8x2-static':
                # activation scales fixed in advance (no pass over the tensor): |x| < 3584 assumed;
                # whatever falls under e4m3's subnormals is flushed (it only feeds corrections)
                return ((x_hi, w_hi), (fp8(x_hi, 2. ** -3), fp8(w_lo)),
                        (fp8(x_lo, 2. ** 8), fp8(w_hi))), 1.
            return ((x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi)), 1.      # fp16 x 3, for reference
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
    args = inputs.synthesis(2, frames, seed=int(__import__('os').environ.get('SEED', '3')))
    double = hifigan.to_double(state)
    dargs = [a.double() if a.is_floating_point() else a for a in args]
    with torch.no_grad():
        exact = hifigan.generator(double, *dargs)
        error = lambda audio: float(
            ((audio.double() - exact).abs().amax(dim=(1, 2)) / exact.abs().amax(dim=(1, 2))).max())
        print(f'fp32 (torch CPU)      {error(hifigan.generator(state, *args)):.2e}')
        schemes = sys.argv[2:] or (
            'bf16x3', 'int8x3', 'int8x3-channel', 'int8x4-channel', 'int8x2', 'fp16+fp8x3', 'fp16+fp8x2')
        for scheme in schemes:
            F.conv1d, F.conv_transpose1d = make(scheme)
            try:
                audio = hifigan.generator(double, *dargs)
            finally:
                F.conv1d, F.conv_transpose1d = conv1d, conv_transpose1d
            print(f'{scheme:21s} {error(audio):.2e}')




def pitch_network(frames=24, schemes=('bf16x3', 'fp16+mxfp8x2')):
    """The same question for the FCNF0++ pitch network of the preprocess chain (oracle/penn.py:
    395 MFLOP per frame, every layer on conv1d_tc_kernel): error of the logits, bar 1e-4"""
    from oracle import penn
    state = penn.init_state(1234)
    double = {k: v.double() for k, v in state.items()}
    audio = inputs.audio(1, 256 * frames + 1024, seed=2)
    with torch.no_grad():
        x = penn.frames(audio, 22050, 256 / 22050).double()
        exact = penn.infer(double, x)
        error = lambda logits: float((logits.double() - exact).abs().max() / exact.abs().max())
        print(f'pitch network, {x.shape[0]} frames: fp32 {error(penn.infer(state, x.float())):.2e}')
        for scheme in schemes:
            conv, _ = make(scheme)
            # the first layer (1 input channel) is an im2col 1 x 1 conv in the product: also rounded
            F.conv1d = lambda x, w, b=None, *a, **k: conv(x, w, b, *a, **k) if w.shape[1] >= 32 \
                else sum(conv1d(p, q, None, *a, **k) for p, q in split_products(scheme, x, w)) + b[None, :, None]
            try:
                logits = penn.infer(double, x)
            finally:
                F.conv1d = conv1d
            print(f'  {scheme:19s} {error(logits):.2e}')


def split_products(scheme, x, w):
    """The k = 32, one-channel first layer as the three (or 1 + 2) products of the scheme, with
    the block scales taken along the taps (its reduction axis)"""
    if scheme == 'bf16x3':
        x_hi, x_lo = split_bf16(x)
        w_hi, w_lo = split_bf16(w)
        return (x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi)
    x_hi, w_hi = x.float().half().double(), w.float().half().double()
    return (x_hi, w_hi), (fp8(x_hi), fp8(w - w_hi)), (fp8(x - x_hi), fp8(w_hi))


if __name__ == '__main__':
    if 'pitch' in sys.argv:
        pitch_network()
    else:
        main()
