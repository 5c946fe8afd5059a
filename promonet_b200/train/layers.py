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
        self.w = self.wt = self.gw = None

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

    def __init__(self, params):
        self.params = params
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

    def refresh(self):
        for layer in self.layers:
            layer.refresh()

    def zero_grad(self):
        self.params.zero_grad()
        self.scratch.zero_()

    def finish(self):
        for layer in self.layers:
            layer.finish()
