"""Flat parameter storage for the training step

All parameters of a module live in ONE fp32 device buffer (`data`), with
matching `grad`, `exp_avg` and `exp_avg_sq` buffers, so that the optimizer is one
kernel launch (pmn_adamw) and the data-parallel gradient exchange is one NCCL
all-reduce per module (train/core.py:255,338 are where the reference would
need them).  Names and shapes are those of the reference state dict
(promonet.model.Generator / Discriminator .state_dict()), so checkpoints
interchange (torchutil.checkpoint format: train/core.py:426-438).
"""
from collections import OrderedDict

import torch

from promonet_b200.train import ops

ALIGN = 4  # floats: every tensor starts on a 16-byte boundary


class ParamSet:

    def __init__(self, state, device, buffers=()):
        self.device = torch.device(device)
        self.buffers = OrderedDict(
            (k, v.detach().to(self.device)) for k, v in state.items() if k in buffers)
        self.index = OrderedDict()
        offset = 0
        for name, value in state.items():
            if name in buffers:
                continue
            self.index[name] = (offset, tuple(value.shape))
            offset += (value.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = offset
        self.data = torch.zeros(offset, device=self.device)
        self.grad = torch.zeros(offset, device=self.device)
        self.exp_avg = torch.zeros(offset, device=self.device)
        self.exp_avg_sq = torch.zeros(offset, device=self.device)
        self.steps = 0
        # the step count also lives on the device (for CUDA-graph replay of the optimizer)
        self.steps_device = torch.zeros(1, device=self.device)
        self.one = torch.ones(1, device=self.device)
        self.load_state_dict(state)

    def _view(self, flat, name):
        offset, shape = self.index[name]
        numel = 1
        for s in shape:
            numel *= s
        return flat[offset:offset + numel].view(shape)

    def __contains__(self, name):
        return name in self.index

    def __getitem__(self, name):
        return self._view(self.data, name)

    def gradient(self, name):
        return self._view(self.grad, name)

    def names(self):
        return list(self.index)

    def load_state_dict(self, state):
        for name in self.index:
            self[name].copy_(state[name].detach().to(self.device, torch.float32))
        for name in self.buffers:
            if name in state:
                self.buffers[name] = state[name].detach().to(self.device)

    def state_dict(self):
        out = OrderedDict((name, self[name].detach().cpu().clone()) for name in self.index)
        for name, value in self.buffers.items():
            out[name] = value.detach().cpu().clone()
        return out

    def gradients(self):
        return OrderedDict((name, self.gradient(name)) for name in self.index)

    def zero_grad(self):
        self.grad.zero_()  # cudaMemsetAsync

    def optimizer_state(self):
        return {
            'exp_avg': self.exp_avg.cpu(), 'exp_avg_sq': self.exp_avg_sq.cpu(),
            'step': self.steps}

    def load_optimizer_state(self, state):
        self.exp_avg.copy_(state['exp_avg'])
        self.exp_avg_sq.copy_(state['exp_avg_sq'])
        self.steps = int(state['step'])
        self.steps_device.fill_(float(self.steps))

    def adamw(self, lr, betas, eps, weight_decay, grad_scale=1.):
        """torch.optim.AdamW.step over every parameter (one launch)"""
        self.steps += 1
        ops.axpby(1., self.one, 1., self.steps_device)
        ops.adamw(
            self.data, self.grad, self.exp_avg, self.exp_avg_sq, lr, betas, eps, weight_decay,
            self.steps, grad_scale, self.steps_device)
