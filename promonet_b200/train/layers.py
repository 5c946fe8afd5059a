"""Convolution layers of the trainable modules

A layer holds views into a ParamSet (weight_g / weight_v / bias, or weight / bias),
the folded weight `w = g v / ||v||` (model/core.py:43-45; recomputed whenever the
parameters change, as torch's weight_norm hook does every forward), its transpose
for the data gradient, and a scratch buffer the weight-gradient kernel accumulates
into before the weight-norm backward turns it into gradients of g and v.
"""
import torch

from promonet_b200.train import ops


class Conv:

    def __init__(self, params, prefix):
        self.params = params
        self.prefix = prefix
        self.weight_norm = f'{prefix}.weight_v' in params
        key = f'{prefix}.weight_v' if self.weight_norm else f'{prefix}.weight'
        self.shape = params.index[key][1]
        self.dim0, self.dim1 = self.shape[0], self.shape[1]
        self.taps = 1
        for s in self.shape[2:]:
            self.taps *= s
        self.numel = self.dim0 * self.dim1 * self.taps
        self.has_bias = f'{prefix}.bias' in params
        # assigned by Layers.allocate()
        self.w = self.wt = self.gw = self.packed = self.packed_t = None
        self.tensor_cores = False

    @property
    def bias(self):
        return self.params[f'{self.prefix}.bias'] if self.has_bias else None

    @property
    def gbias(self):
        return self.params.gradient(f'{self.prefix}.bias') if self.has_bias else None

    def refresh(self):
        """Fold weight norm and transpose (after every optimizer step)"""
        if self.weight_norm:
            ops.weight_norm_fold(
                self.params[f'{self.prefix}.weight_v'], self.params[f'{self.prefix}.weight_g'],
                self.w, self.dim0, self.dim1 * self.taps)
        ops.transpose_weight(self.w, self.wt, self.dim0, self.dim1, self.taps)
        if self.packed is not None:
            ops.pack_weight_taps(self.w, self.packed, self.dim0, self.dim1, self.taps, False)
        if self.packed_t is not None:
            ops.pack_weight_taps(self.w, self.packed_t, self.dim0, self.dim1, self.taps, True)

    def apply(self, geometry, transposed, a, out, **kwargs):
        """Implicit GEMM with rows = dim 0 of the weight, reducing over dim 1 (forward of a
        Conv, data gradient of a ConvTranspose).  `transposed` is the gather direction."""
        if self.packed is not None and kwargs.get('out_act') != ops.OUT_TANH:
            return ops.conv_gemm_tc(geometry, transposed, a, self.packed, out, **kwargs)
        return ops.conv_gemm(geometry, transposed, a, self.w, out, **kwargs)

    TENSOR_CORE_WGRAD = {
        (ops.ACT_NONE, ops.ACT_NONE), (ops.ACT_NONE, ops.ACT_LRELU),
        (ops.ACT_LRELU_MASK, ops.ACT_NONE), (ops.ACT_LRELU, ops.ACT_NONE)}

    def wgrad(self, geometry, dy, x, bias=True, **kwargs):
        """Accumulate the weight (and bias) gradient of the convolution `geometry`"""
        gbias = self.gbias if bias else None
        pair = (kwargs.get('dy_act', ops.ACT_NONE), kwargs.get('x_act', ops.ACT_NONE))
        if self.packed is not None and pair in self.TENSOR_CORE_WGRAD:
            return ops.conv_wgrad_tc(geometry, dy, x, self.gw, gbias, **kwargs)
        return ops.conv_wgrad(geometry, dy, x, self.gw, gbias, **kwargs)

    def apply_transposed(self, geometry, transposed, a, out, **kwargs):
        """Rows = dim 1 of the weight, reducing over dim 0 (data gradient of a Conv)"""
        if self.packed_t is not None:
            return ops.conv_gemm_tc(geometry, transposed, a, self.packed_t, out, **kwargs)
        return ops.conv_gemm(geometry, transposed, a, self.wt, out, **kwargs)

    def finish(self):
        """Weight-norm backward: gw -> gradients of weight_g and weight_v"""
        if self.weight_norm:
            ops.weight_norm_backward(
                self.params[f'{self.prefix}.weight_v'], self.params[f'{self.prefix}.weight_g'],
                self.gw, self.params.gradient(f'{self.prefix}.weight_v'),
                self.params.gradient(f'{self.prefix}.weight_g'), self.dim0,
                self.dim1 * self.taps)


class Layers:
    """The convolutions of one module, with flat derived / scratch storage"""

    # a GEMM goes to the tensor cores when both its row and reduction channel counts reach this
    # (measured: even the 1-channel first and last layers are faster there than on the FMA tiles)
    TENSOR_CORE_MIN_CHANNELS = 1

    def __init__(self, params, math='tf32'):
        if math not in ('tf32', 'fp32'):
            raise ValueError(f'unknown math mode {math}')
        self.params = params
        self.math = math
        self.layers = []

    def conv(self, prefix):
        layer = Conv(self.params, prefix)
        self.layers.append(layer)
        return layer

    def allocate(self):
        device = self.params.device
        total = sum(layer.numel for layer in self.layers)
        normed = sum(layer.numel for layer in self.layers if layer.weight_norm)
        self.folded = torch.empty(normed, device=device)
        self.transposed = torch.empty(total, device=device)
        self.scratch = torch.zeros(normed, device=device)
        packable = [
            layer for layer in self.layers
            if self.math == 'tf32' and min(layer.dim0, layer.dim1) >= self.TENSOR_CORE_MIN_CHANNELS]
        sizes = [
            (ops.packed_floats(layer.dim0, layer.dim1, layer.taps),
             ops.packed_floats(layer.dim1, layer.dim0, layer.taps)) for layer in packable]
        self.packed = torch.empty(sum(a + b for a, b in sizes), device=device)
        offset = 0
        for layer, (forward, backward) in zip(packable, sizes):
            layer.packed = self.packed[offset:offset + forward]
            layer.packed_t = self.packed[offset + forward:offset + forward + backward]
            offset += forward + backward
        f = t = 0
        for layer in self.layers:
            layer.wt = self.transposed[t:t + layer.numel]
            t += layer.numel
            if layer.weight_norm:
                layer.w = self.folded[f:f + layer.numel]
                layer.gw = self.scratch[f:f + layer.numel]
                f += layer.numel
            else:
                layer.w = self.params[f'{layer.prefix}.weight']
                layer.gw = self.params.gradient(f'{layer.prefix}.weight')
        entries = []
        for layer in self.layers:
            key = 'weight_v' if layer.weight_norm else 'weight'
            fma = layer.packed is None      # the FMA path needs the plain transpose
            entries.append({
                'v': self.params[f'{layer.prefix}.{key}'],
                'g': self.params[f'{layer.prefix}.weight_g'] if layer.weight_norm else None,
                'w': layer.w if layer.weight_norm else None,
                'packed': layer.packed, 'packed_t': layer.packed_t,
                'wt': layer.wt if fma else None,
                'dim0': layer.dim0, 'dim1': layer.dim1, 'taps': layer.taps})
        self.table = ops.weight_table(entries, device)
        self.max_dim0 = max(layer.dim0 for layer in self.layers)

    def refresh(self):
        """After an optimizer step: fold every weight norm and rebuild the operand
        packings of every layer (two launches for the whole module)"""
        ops.prepare_weights(self.table, len(self.layers), self.max_dim0)

    def zero_grad(self):
        self.params.zero_grad()
        self.scratch.zero_()

    def finish(self):
        for layer in self.layers:
            layer.finish()
