"""Convolution layers of the trainable modules

A layer holds views into a ParamSet (weight_g / weight_v / bias, or weight / bias),
the folded weight `w = g v / ||v||` (model/core.py:43-45; recomputed whenever the
parameters change, as torch's weight_norm hook does every forward), the operand
packings of the tensor-core kernels (or the plain transpose for the FMA path), and a
scratch buffer the weight-gradient kernel accumulates into before the weight-norm
backward turns it into gradients of g and v.

A grouped convolution (DiscriminatorS, model/discriminator.py:218-224) is run as the dense
convolution with a block-diagonal weight: its parameters keep torch's (C_out, C_in / groups, k)
shape, the packings are written with zeros off the diagonal, and the dense weight gradient
is reduced to its diagonal blocks before the weight-norm backward.
"""
import torch

from promonet_b200.train import ops


class Conv:

    def __init__(self, params, prefix, groups=1):
        self.params = params
        self.prefix = prefix
        self.groups = groups
        self.weight_norm = f'{prefix}.weight_v' in params
        key = f'{prefix}.weight_v' if self.weight_norm else f'{prefix}.weight'
        self.shape = params.index[key][1]
        self.dim0, self.dim1 = self.shape[0], self.shape[1] * groups   # dense GEMM dimensions
        self.taps = 1
        for s in self.shape[2:]:
            self.taps *= s
        self.numel = self.dim0 * self.dim1 * self.taps            # dense
        self.stored = self.numel // groups                         # as the parameters store it
        self.has_bias = f'{prefix}.bias' in params
        # assigned by Layers.allocate()
        self.w = self.wt = self.gw = self.packed = self.packed_t = None
        self.dense = self.gw_dense = None

    @property
    def bias(self):
        return self.params[f'{self.prefix}.bias'] if self.has_bias else None

    @property
    def gbias(self):
        return self.params.gradient(f'{self.prefix}.bias') if self.has_bias else None

    def apply(self, geometry, transposed, a, out, **kwargs):
        """Implicit GEMM with rows = dim 0 of the weight, reducing over dim 1 (forward of a
        Conv, data gradient of a ConvTranspose).  `transposed` is the gather direction."""
        if self.packed is not None and kwargs.get('out_act') != ops.OUT_TANH:
            return ops.conv_gemm_tc(geometry, transposed, a, self.packed, out, **kwargs)
        weight = self.dense if self.groups > 1 else self.w
        return ops.conv_gemm(geometry, transposed, a, weight, out, **kwargs)

    def apply_transposed(self, geometry, transposed, a, out, **kwargs):
        """Rows = dim 1 of the weight, reducing over dim 0 (data gradient of a Conv)"""
        if self.packed_t is not None:
            return ops.conv_gemm_tc(geometry, transposed, a, self.packed_t, out, **kwargs)
        return ops.conv_gemm(geometry, transposed, a, self.wt, out, **kwargs)

    TENSOR_CORE_WGRAD = {
        (ops.ACT_NONE, ops.ACT_NONE), (ops.ACT_NONE, ops.ACT_LRELU),
        (ops.ACT_LRELU_MASK, ops.ACT_NONE), (ops.ACT_LRELU, ops.ACT_NONE)}

    def wgrad(self, geometry, dy, x, bias=True, **kwargs):
        """Accumulate the weight (and bias) gradient of the convolution `geometry`"""
        gbias = self.gbias if bias else None
        target = self.gw_dense if self.groups > 1 else self.gw
        pair = (kwargs.get('dy_act', ops.ACT_NONE), kwargs.get('x_act', ops.ACT_NONE))
        if self.packed is not None and pair in self.TENSOR_CORE_WGRAD:
            return ops.conv_wgrad_tc(geometry, dy, x, target, gbias, **kwargs)
        return ops.conv_wgrad(geometry, dy, x, target, gbias, **kwargs)

    def finish(self):
        """Weight-norm backward: gw -> gradients of weight_g and weight_v"""
        if self.groups > 1:
            ops.extract_grouped(self.gw_dense, self.gw, self.dim0, self.dim1, self.taps, self.groups)
        if self.weight_norm:
            ops.weight_norm_backward(
                self.params[f'{self.prefix}.weight_v'], self.params[f'{self.prefix}.weight_g'],
                self.gw, self.params.gradient(f'{self.prefix}.weight_v'),
                self.params.gradient(f'{self.prefix}.weight_g'), self.dim0,
                self.stored // self.dim0)


class Layers:
    """The convolutions of one module, with flat derived / scratch storage"""

    def __init__(self, params, math='tf32'):
        if math not in ('tf32', 'fp32'):
            raise ValueError(f'unknown math mode {math}')
        self.params = params
        self.math = math
        self.layers = []

    def conv(self, prefix, groups=1):
        layer = Conv(self.params, prefix, groups)
        self.layers.append(layer)
        return layer

    def allocate(self):
        device = self.params.device
        tensor_cores = self.math == 'tf32'   # every layer: even the 1-channel ones are faster there
        flat = lambda count, zero=False: (torch.zeros if zero else torch.empty)(count, device=device)
        normed = [layer for layer in self.layers if layer.weight_norm]
        grouped = [layer for layer in self.layers if layer.groups > 1]
        self.folded = flat(sum(layer.stored for layer in normed))
        # one buffer cleared per backward: weight-gradient scratch of the weight-normed layers and
        # the dense gradients of the grouped ones
        self.scratch = flat(
            sum(layer.stored for layer in normed) + sum(layer.numel for layer in grouped), zero=True)
        sizes = [
            (ops.packed_floats(layer.dim0, layer.dim1, layer.taps),
             ops.packed_floats(layer.dim1, layer.dim0, layer.taps)) if tensor_cores else (0, 0)
            for layer in self.layers]
        if any(max(a, b) >= 2 ** 31 for a, b in sizes):
            raise ValueError('a packed weight exceeds 2^31 elements (pack_weights_kernel indexes with 32 bits)')
        self.packed = flat(sum(a + b for a, b in sizes))
        self.transposed = flat(0 if tensor_cores else sum(layer.numel for layer in self.layers))
        self.dense = flat(0 if tensor_cores else sum(layer.numel for layer in grouped))
        f = s = k = t = d = 0
        entries = []
        for layer, (forward, backward) in zip(self.layers, sizes):
            if tensor_cores:
                layer.packed = self.packed[k:k + forward]
                layer.packed_t = self.packed[k + forward:k + forward + backward]
                k += forward + backward
            else:
                layer.wt = self.transposed[t:t + layer.numel]
                t += layer.numel
                if layer.groups > 1:
                    layer.dense = self.dense[d:d + layer.numel]
                    d += layer.numel
            if layer.weight_norm:
                layer.w = self.folded[f:f + layer.stored]
                layer.gw = self.scratch[s:s + layer.stored]
                f += layer.stored
                s += layer.stored
            else:
                layer.w = self.params[f'{layer.prefix}.weight']
                layer.gw = self.params.gradient(f'{layer.prefix}.weight')
            if layer.groups > 1:
                layer.gw_dense = self.scratch[s:s + layer.numel]
                s += layer.numel
            key = 'weight_v' if layer.weight_norm else 'weight'
            entries.append({
                'v': self.params[f'{layer.prefix}.{key}'],
                'g': self.params[f'{layer.prefix}.weight_g'] if layer.weight_norm else None,
                'w': layer.w if layer.weight_norm else None,
                'packed': layer.packed, 'packed_t': layer.packed_t,
                'wt': layer.wt, 'dense': layer.dense,
                'dim0': layer.dim0, 'dim1': layer.dim1, 'taps': layer.taps,
                'groups': layer.groups})
        self.table = ops.weight_table(entries, device)
        self.max_dim0 = max(layer.dim0 for layer in self.layers)
        # weight-norm backward of every normed layer in one launch (finish)
        self.normed = normed
        self.norm_table = ops.weight_norm_table([{
            'v': self.params[f'{layer.prefix}.weight_v'], 'g': self.params[f'{layer.prefix}.weight_g'],
            'gw': layer.gw, 'gv': self.params.gradient(f'{layer.prefix}.weight_v'),
            'gg': self.params.gradient(f'{layer.prefix}.weight_g'),
            'dim0': layer.dim0, 'inner': layer.stored // layer.dim0} for layer in normed], device) \
            if normed else None

    def refresh(self):
        """After an optimizer step: fold every weight norm and rebuild the operand
        packings of every layer (two launches for the whole module)"""
        ops.prepare_weights(self.table, len(self.layers), self.max_dim0)

    def zero_grad(self):
        self.params.zero_grad()
        self.scratch.zero_()

    def finish(self):
        """Weight gradients -> parameter gradients: grouped layers keep the diagonal blocks of their
        dense gradient, then one launch runs the weight-norm backward of every normed layer"""
        for layer in self.layers:
            if layer.groups > 1:
                ops.extract_grouped(
                    layer.gw_dense, layer.gw, layer.dim0, layer.dim1, layer.taps, layer.groups)
        if self.normed:
            ops.weight_norm_backward_table(
                self.norm_table, len(self.normed), max(layer.dim0 for layer in self.normed))
