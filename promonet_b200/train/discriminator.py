"""Trainable discriminators: 5 x multi-period + complex multi-band
(promonet/model/discriminator.py:13-93,146-208 under config/promonet.py), optionally the
multi-scale (:211-239) and the three multi-resolution (:96-141) ones.

`forward(x)` runs every sub-discriminator on a batch whose first half is real
audio and second half generated audio (the reference calls each sub-discriminator
twice, :41-43; the layers have no batch statistics, so one pass over the
concatenation is the same arithmetic) and keeps the feature maps.  `backward`
walks the layers in reverse over a slice of that batch: all of it with weight
gradients for the discriminator step (train/core.py:239-256), the generated
half with input gradients only for the generator step (:272-338).
"""
import torch

from promonet_b200 import config
from promonet_b200.model import init
from promonet_b200.train import ops
from promonet_b200.train.layers import Layers
from promonet_b200.train.params import ParamSet

SLOPE = config.LRELU_SLOPE
PAD = (config.NUM_FFT - config.HOPSIZE) // 2


class Period:
    """DiscriminatorP discriminator.py:57-93"""

    def __init__(self, layers, prefix, period):
        self.period = period
        self.convs = [layers.conv(f'{prefix}.convs.{i}') for i in range(5)]
        self.post = layers.conv(f'{prefix}.conv_post')

    def forward(self, x):
        """x (N, 1, T) -> record with maps[0] = input view, maps[1..6] = feature maps"""
        n, _, t = x.shape
        pad = (self.period - t % self.period) % self.period
        if pad:
            x = ops.reflect_pad(x, 0, pad)
        height = (t + pad) // self.period
        maps = [x.view(n, 1, height, self.period)]
        geometries = []
        for i, conv in enumerate(self.convs + [self.post]):
            kernel, stride, padding = ((5, 1), (3, 1) if i < 4 else 1, (2, 0)) if i < 5 \
                else ((3, 1), 1, (1, 0))
            source = maps[-1]
            geometry = ops.geometry(
                n, conv.dim1, conv.dim0, source.shape[2:], kernel, stride, 1, padding)
            out = torch.empty(n, conv.dim0, geometry.h_out, geometry.w_out, device=x.device)
            conv.apply(geometry, False, source, out, bias=conv.bias,
                out_act=ops.OUT_LRELU if i < 5 else ops.OUT_NONE, out_slope=SLOPE)
            maps.append(out)
            geometries.append(geometry)
        return {'maps': maps, 'geometries': geometries, 'pad': pad, 'samples': t}

    def backward(self, record, gmaps, lo, hi, weights, gaudio):
        """gmaps[i]: gradient of feature map i over items lo:hi, or None until something
        writes it; gmaps[5] (the logits) must be set.  gaudio (hi - lo, 1, T) accumulates."""
        layers = self.convs + [self.post]
        n = hi - lo
        for i in reversed(range(6)):
            geometry = _with_batch(record['geometries'][i], n)
            y, x = record['maps'][i + 1][lo:hi], record['maps'][i][lo:hi]
            act = (ops.ACT_LRELU_MASK, SLOPE) if i < 5 else (ops.ACT_NONE, 1.)
            g = gmaps[i]
            if weights:
                layers[i].wgrad(geometry, g, x, dy_companion=y, dy_act=act[0], dy_slope=act[1])
            if i == 0 and gaudio is None:
                break
            target = gmaps[i - 1] if i > 0 else None
            accumulate = target is not None
            if target is None:
                target = torch.empty_like(x)
            layers[i].apply_transposed(geometry, True, g, target, a_companion=y,
                          a_act=act[0], a_slope=act[1], accumulate=accumulate)
            if i > 0:
                gmaps[i - 1] = target
            else:
                ginput = target.view(n, 1, -1)
                if record['pad']:
                    ops.reflect_pad_backward(ginput, gaudio, 0, record['pad'], accumulate=True)
                else:
                    ops.axpby(1., ginput, 1., gaudio)


class Scale:
    """DiscriminatorS discriminator.py:211-239 (MULTI_SCALE_DISCRIMINATOR): grouped Conv1d stack
    over the waveform; the groups run as block-diagonal dense convolutions (train/layers.py)"""

    def __init__(self, layers, prefix):
        self.specs = init.MULTI_SCALE_CONVS
        self.convs = [
            layers.conv(f'{prefix}.convs.{i}', groups=spec[4]) for i, spec in enumerate(self.specs)]
        self.post = layers.conv(f'{prefix}.conv_post')

    def forward(self, x):
        n, _, t = x.shape
        maps = [x.view(n, 1, t, 1)]
        geometries = []
        specs = [(k, s, p) for _, _, k, s, _, p in self.specs] + [(3, 1, 1)]
        for i, (conv, (kernel, stride, padding)) in enumerate(zip(self.convs + [self.post], specs)):
            source = maps[-1]
            geometry = ops.geometry(
                n, conv.dim1, conv.dim0, source.shape[2:], (kernel, 1), (stride, 1), 1, (padding, 0))
            out = torch.empty(n, conv.dim0, geometry.h_out, 1, device=x.device)
            conv.apply(geometry, False, source, out, bias=conv.bias,
                       out_act=ops.OUT_LRELU if i < 6 else ops.OUT_NONE, out_slope=SLOPE)
            maps.append(out)
            geometries.append(geometry)
        return {'maps': maps, 'geometries': geometries, 'pad': 0, 'samples': t}

    def backward(self, record, gmaps, lo, hi, weights, gaudio):
        layers = self.convs + [self.post]
        n = hi - lo
        last = len(layers) - 1
        for i in reversed(range(len(layers))):
            geometry = _with_batch(record['geometries'][i], n)
            y, x = record['maps'][i + 1][lo:hi], record['maps'][i][lo:hi]
            act = (ops.ACT_LRELU_MASK, SLOPE) if i < last else (ops.ACT_NONE, 1.)
            g = gmaps[i]
            if weights:
                layers[i].wgrad(geometry, g, x, dy_companion=y, dy_act=act[0], dy_slope=act[1])
            if i == 0 and gaudio is None:
                break
            target = gmaps[i - 1] if i > 0 else None
            accumulate = target is not None
            if target is None:
                target = torch.empty_like(x)
            layers[i].apply_transposed(
                geometry, True, g, target, a_companion=y, a_act=act[0], a_slope=act[1],
                accumulate=accumulate)
            if i > 0:
                gmaps[i - 1] = target
            else:
                ops.axpby(1., target.view(n, 1, -1), 1., gaudio)


class Resolution:
    """DiscriminatorR discriminator.py:96-141 (MULTI_RESOLUTION_DISCRIMINATOR, off by default).
    The STFT (window=None, win_length < n_fft, hop not a divisor of n_fft) is a 1 x 1
    convolution over the reflect-padded signal read in place as overlapping frames, with the
    DFT basis as its weight: the same construction as the spectral-convergence loss
    (train/losses.py), so its forward and backward are the conv kernels."""

    SLOPE = 0.2     # discriminator.py:121 (not LRELU_SLOPE)

    def __init__(self, layers, prefix, resolution, math):
        self.n_fft, self.hop, self.win = resolution
        self.math = math
        self.convs = [layers.conv(f'{prefix}.convs.{i}') for i in range(5)]
        self.post = layers.conv(f'{prefix}.conv_post')
        self.basis = None

    def _basis(self, device):
        """(rows, forward operand, backward operand) of the DFT convolution, built once"""
        if self.basis is None:
            basis = ops.dft_basis_rect(self.n_fft, self.win, device)     # (2 bins, n_fft)
            rows = basis.shape[0]
            if self.math == 'tf32':
                forward = ops.pack_weight_taps(
                    basis, torch.empty(ops.packed_floats(rows, self.n_fft, 1), device=device),
                    rows, self.n_fft, 1, False)
                backward = ops.pack_weight_taps(
                    basis, torch.empty(ops.packed_floats(self.n_fft, rows, 1), device=device),
                    rows, self.n_fft, 1, True)
            else:
                forward = basis
                backward = ops.transpose_weight(basis, torch.empty_like(basis), rows, self.n_fft, 1)
            self.basis = (rows, forward, backward)
        return self.basis

    def _dft(self, geometry, a, weight, out):
        if self.math == 'tf32':
            return ops.conv_gemm_tc(geometry, False, a, weight, out)
        return ops.conv_gemm(geometry, False, a, weight, out)

    def forward(self, x):
        """x (N, 1, T) -> record with maps[0] = |STFT| (N, 1, bins, frames), maps[1..6] = feature maps"""
        n, _, t = x.shape
        pad = (self.n_fft - self.hop) // 2                               # :129-132
        padded = ops.reflect_pad(x.view(n, t), pad, pad)
        length = t + 2 * pad
        frames = 1 + (length - self.n_fft) // self.hop                   # center=False
        rows, forward, _ = self._basis(x.device)
        geometry = ops.geometry(
            n, self.n_fft, rows, (frames, 1), (1, 1), strides=(1, self.hop, length))
        spec = self._dft(geometry, padded, forward, torch.empty(n, rows, frames, device=x.device))
        maps = [ops.complex_magnitude(spec)]
        geometries = []
        for i, conv in enumerate(self.convs + [self.post]):
            kernel = (3, 9) if i < 4 else (3, 3)
            stride = (1, 2) if 1 <= i <= 3 else (1, 1)
            source = maps[-1]
            geometry = ops.geometry(
                n, conv.dim1, conv.dim0, source.shape[2:], kernel, stride, 1, (1, kernel[1] // 2))
            out = torch.empty(n, conv.dim0, geometry.h_out, geometry.w_out, device=x.device)
            conv.apply(geometry, False, source, out, bias=conv.bias,
                       out_act=ops.OUT_LRELU if i < 5 else ops.OUT_NONE, out_slope=self.SLOPE)
            maps.append(out)
            geometries.append(geometry)
        return {
            'maps': maps, 'geometries': geometries, 'spec': spec, 'pad': pad, 'samples': t,
            'frames': frames, 'length': length}

    def backward(self, record, gmaps, lo, hi, weights, gaudio):
        """As Period.backward; the gradient of maps[0] goes back through the magnitude, the DFT
        convolution, the overlap-add of the frames and the reflect padding"""
        layers = self.convs + [self.post]
        n = hi - lo
        for i in reversed(range(6)):
            geometry = _with_batch(record['geometries'][i], n)
            y, x = record['maps'][i + 1][lo:hi], record['maps'][i][lo:hi]
            act = (ops.ACT_LRELU_MASK, self.SLOPE) if i < 5 else (ops.ACT_NONE, 1.)
            g = gmaps[i]
            if weights:
                layers[i].wgrad(geometry, g, x, dy_companion=y, dy_act=act[0], dy_slope=act[1])
            if i == 0 and gaudio is None:
                break
            target = gmaps[i - 1] if i > 0 else None
            accumulate = target is not None
            if target is None:
                target = torch.empty_like(x)
            layers[i].apply_transposed(
                geometry, True, g, target, a_companion=y, a_act=act[0], a_slope=act[1],
                accumulate=accumulate)
            if i > 0:
                gmaps[i - 1] = target
                continue
            frames, length, pad = record['frames'], record['length'], record['pad']
            rows, _, backward = self._basis(target.device)
            gspec = ops.complex_magnitude_backward(target, record['spec'][lo:hi])
            geometry = ops.geometry(n, rows, self.n_fft, (frames, 1), (1, 1))
            gframes = self._dft(
                geometry, gspec, backward, torch.empty(n, self.n_fft, frames, device=target.device))
            gpadded = torch.zeros(n, length, device=target.device)
            ops.frame_overlap_add(gframes, gpadded, self.hop)
            ops.reflect_pad_backward(gpadded, gaudio.view(n, -1), pad, pad, accumulate=True)


class ComplexMultiBand:
    """DiscriminatorCMB discriminator.py:146-208"""

    def __init__(self, layers, prefix):
        self.bands = [
            [layers.conv(f'{prefix}.band_convs.{b}.{i}.0') for i in range(5)]
            for b, _ in enumerate(config.CMB_BANDS)]
        self.post = layers.conv(f'{prefix}.conv_post')

    def forward(self, x):
        n, _, t = x.shape
        # rectangular-window STFT magnitude, written band by band: (N, 1, F, width) each
        banded, spectrum = ops.stft_magnitude(x.view(n, t), 'rect', 0., layout=2)
        frames = t // config.HOPSIZE
        maps, geometries, outputs = [], [], []
        for (lo, hi), stack in zip(config.CMB_BANDS, self.bands):
            width = hi - lo
            band = banded[n * frames * lo:n * frames * hi].view(n, 1, frames, width)
            chain, chain_geometry = [band], []
            for i, conv in enumerate(stack):
                kernel = (3, 9) if i < 4 else (3, 3)
                stride = (1, 2) if 1 <= i <= 3 else (1, 1)
                geometry = ops.geometry(
                    n, conv.dim1, conv.dim0, chain[-1].shape[2:], kernel, stride, 1,
                    (1, kernel[1] // 2))
                out = torch.empty(n, conv.dim0, geometry.h_out, geometry.w_out, device=x.device)
                conv.apply(geometry, False, chain[-1], out, bias=conv.bias,
                              out_act=ops.OUT_LRELU, out_slope=SLOPE)
                chain.append(out)
                chain_geometry.append(geometry)
            maps.append(chain)
            geometries.append(chain_geometry)
            outputs.append(chain[-1])
        total = sum(o.shape[-1] for o in outputs)
        joined = torch.empty(n, 32, frames, total, device=x.device)
        column = 0
        for o in outputs:
            ops.copy_columns(o, joined, 0, column, o.shape[-1])
            column += o.shape[-1]
        post_geometry = ops.geometry(n, 32, 1, (frames, total), (3, 3), 1, 1, (1, 1))
        logits = torch.empty(n, 1, frames, total, device=x.device)
        self.post.apply(post_geometry, False, joined, logits, bias=self.post.bias)
        return {
            'maps': maps, 'geometries': geometries, 'joined': joined, 'logits': logits,
            'post_geometry': post_geometry, 'spectrum': spectrum, 'samples': t, 'frames': frames}

    @staticmethod
    def feature_maps(record):
        """In the reference's order (:201-207): per band its 5 maps, then the post map"""
        maps = [m for chain in record['maps'] for m in chain[1:]]
        return maps + [record['logits']]

    def backward(self, record, gmaps, lo, hi, weights, gaudio):
        """gmaps: gradients in feature_maps() order over items lo:hi (None = not yet written);
        the last entry (logits) must be set"""
        n, frames = hi - lo, record['frames']
        glogits = gmaps[-1]
        geometry = _with_batch(record['post_geometry'], n)
        joined = record['joined'][lo:hi]
        if weights:
            self.post.wgrad(geometry, glogits, joined)
        gjoined = self.post.apply_transposed(geometry, True, glogits, torch.empty_like(joined))
        gbanded = torch.zeros(n * frames * 513, device=gjoined.device) if gaudio is not None else None
        column = 0
        for b, ((first, last), stack) in enumerate(zip(config.CMB_BANDS, self.bands)):
            chain = record['maps'][b]
            width = chain[-1].shape[-1]
            g = gmaps[5 * b + 4]
            if g is None:
                g = torch.empty(n, 32, frames, width, device=gjoined.device)
                ops.copy_columns(gjoined, g, column, 0, width)
            else:
                ops.copy_columns(gjoined, g, column, 0, width, accumulate=True)
            column += width
            for i in reversed(range(5)):
                geometry = _with_batch(record['geometries'][b][i], n)
                y, x = chain[i + 1][lo:hi], chain[i][lo:hi]
                if weights:
                    stack[i].wgrad(
                        geometry, g, x, dy_companion=y, dy_act=ops.ACT_LRELU_MASK, dy_slope=SLOPE)
                if i == 0 and gaudio is None:
                    break
                if i > 0:
                    target = gmaps[5 * b + i - 1]
                    accumulate = target is not None
                    if target is None:
                        target = torch.empty_like(x)
                else:
                    target = gbanded[n * frames * first:n * frames * last].view(x.shape)
                    accumulate = False
                stack[i].apply_transposed(geometry, True, g, target, a_companion=y,
                              a_act=ops.ACT_LRELU_MASK, a_slope=SLOPE, accumulate=accumulate)
                g = target
        if gaudio is not None:
            ops.stft_magnitude_backward(
                gbanded, record['spectrum'][lo:hi], gaudio.view(n, -1), 'rect', 0., layout=2,
                accumulate=True)


def _with_batch(geometry, batch):
    if geometry.batch == batch:
        return geometry
    copy = type(geometry).from_buffer_copy(geometry)
    copy.batch = batch
    return copy


def sub_discriminators(state):
    """The sub-discriminators a state dict holds, in order (discriminator.py:15-34): 'p' x 5, 's' if
    MULTI_SCALE_DISCRIMINATOR, 'r' x 3 if MULTI_RESOLUTION_DISCRIMINATOR, 'cmb'.  Anything else
    (FARGAN_DISCRIMINATOR, a subset) is not built and raises."""
    kinds = []
    for index in range(len({k.split('.')[1] for k in state})):
        prefix = f'discriminators.{index}'
        first = state.get(f'{prefix}.convs.0.weight_v')
        if f'{prefix}.band_convs.0.0.0.weight_v' in state:
            kinds.append('cmb')
        elif first is None:
            kinds.append('?')
        elif first.ndim == 3:
            kinds.append('s')
        elif tuple(first.shape[1:]) == (1, 3, 9):
            kinds.append('r')
        elif tuple(first.shape[1:]) == (1, 5, 1):
            kinds.append('p')
        else:
            kinds.append('?')
    periods = len(config.DISCRIMINATOR_PERIODS)
    allowed = [
        ['p'] * periods + scale + resolution + ['cmb']
        for scale in ([], ['s']) for resolution in ([], ['r'] * len(init.MULTI_RESOLUTIONS))]
    if kinds not in allowed:
        raise NotImplementedError(
            f'sub-discriminators {kinds}: only 5 x DiscriminatorP (+ DiscriminatorS) '
            '(+ 3 x DiscriminatorR) + DiscriminatorCMB are built (DESIGN.md section 7)')
    return kinds


class Discriminator:

    def __init__(self, state=None, device=None, math='tf32', multi_scale=False, peer_group=None,
                 multi_resolution=False):
        """multi_scale = MULTI_SCALE_DISCRIMINATOR (config/defaults.py:180), multi_resolution =
        MULTI_RESOLUTION_DISCRIMINATOR (:177); both are inferred from the keys when a state dict
        is given"""
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200.train needs a CUDA device (sm_100a); there is no CPU path')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if state is None:
            state = init.discriminator_state(multi_scale=multi_scale, multi_resolution=multi_resolution)
        kinds = sub_discriminators(state)
        multi_scale = 's' in kinds
        self.params = ParamSet(state, self.device, peer_group=peer_group)
        self.layers = Layers(self.params, math)
        self.modules = [
            Period(self.layers, f'discriminators.{i}', period)
            for i, period in enumerate(config.DISCRIMINATOR_PERIODS)]
        if multi_scale:
            self.modules.append(Scale(self.layers, f'discriminators.{len(self.modules)}'))
        for resolution in init.MULTI_RESOLUTIONS if 'r' in kinds else ():
            self.modules.append(Resolution(
                self.layers, f'discriminators.{len(self.modules)}', resolution, math))
        self.cmb = ComplexMultiBand(self.layers, f'discriminators.{len(self.modules)}')
        self.layers.allocate()

    def state_dict(self):
        return self.params.state_dict()

    def load_state_dict(self, state):
        self.params.load_state_dict(state)

    def refresh(self):
        self.layers.refresh()

    def forward(self, x):
        """x (N, 1, T) -> records (one per sub-discriminator)"""
        return [m.forward(x) for m in self.modules] + [self.cmb.forward(x)]

    def logits(self, records):
        """Flattened logits per sub-discriminator, (N, n_i) each (discriminator.py:93,208)"""
        n = records[0]['maps'][0].shape[0]
        return [r['maps'][-1].view(n, -1) for r in records[:-1]] + \
               [records[-1]['logits'].view(n, -1)]

    def feature_maps(self, records):
        """Lists of feature maps per sub-discriminator in the reference's order"""
        return [r['maps'][1:] for r in records[:-1]] + [self.cmb.feature_maps(records[-1])]

    def backward(self, records, gmaps, lo, hi, weights, gaudio=None):
        """gmaps: per sub-discriminator, a list aligned with feature_maps() holding the
        gradient of each map over items lo:hi or None"""
        for module, record, g in zip(self.modules + [self.cmb], records, gmaps):
            module.backward(record, g, lo, hi, weights, gaudio)
        if weights:
            self.layers.finish()

    __call__ = forward
