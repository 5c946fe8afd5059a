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

    def __init__(self, state, device, buffers=(), peer_group=None):
        """peer_group: a torch.distributed group of more than one rank puts the parameter and
        gradient buffers in symmetric memory (every rank can address every rank's buffers over
        NVLink), which is what the fused data-parallel optimizer step (adamw_peer) needs"""
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
        self.peers = None
        if peer_group is not None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symmetric
            self.data = symmetric.empty(offset, dtype=torch.float32, device=self.device).zero_()
            self.grad = symmetric.empty(offset, dtype=torch.float32, device=self.device).zero_()
            self.peers = {
                'data': symmetric.rendezvous(self.data, group=peer_group),
                'grad': symmetric.rendezvous(self.grad, group=peer_group),
                'group': peer_group,
                'rank': dist.get_rank(peer_group), 'world': dist.get_world_size(peer_group)}
            # this rank's slice of the flat buffers (ZeRO-1: it keeps the moments of that slice)
            shard = (offset + self.peers['world'] - 1) // self.peers['world']
            shard = (shard + ALIGN - 1) // ALIGN * ALIGN
            self.peers['shard'] = shard
            self.peers['begin'] = min(self.peers['rank'] * shard, offset)
            self.peers['end'] = min(self.peers['begin'] + shard, offset)
        else:
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

    def _whole(self, moment):
        """The moment buffer with every rank's slice (a collective when the step is sharded)"""
        if self.peers is None:
            return moment
        import torch.distributed as dist
        shard, world = self.peers['shard'], self.peers['world']
        mine = torch.zeros(shard, device=self.device)
        count = self.peers['end'] - self.peers['begin']
        mine[:count].copy_(moment[self.peers['begin']:self.peers['end']])
        whole = torch.empty(shard * world, device=self.device)
        dist.all_gather_into_tensor(whole, mine, group=self.peers['group'])
        return whole[:self.numel]

    def optimizer_state(self):
        return {
            'exp_avg': self._whole(self.exp_avg).cpu(), 'exp_avg_sq': self._whole(self.exp_avg_sq).cpu(),
            'step': self.steps}

    def load_optimizer_state(self, state):
        self.exp_avg.copy_(state['exp_avg'])
        self.exp_avg_sq.copy_(state['exp_avg_sq'])
        self.steps = int(state['step'])
        self.steps_device.fill_(float(self.steps))

    def adamw_peer(self, lr, betas, eps, weight_decay):
        """The data-parallel step in one kernel over NVLink peer memory: this rank averages its
        slice of every rank's gradients, applies AdamW to it and writes the new parameters into
        every rank's buffer (pmn_adamw_peer).  The two barriers are device-side, on the current
        stream: gradients complete everywhere before, parameter writes landed everywhere after."""
        self.steps += 1
        ops.axpby(1., self.one, 1., self.steps_device)
        self.peers['grad'].barrier()
        ops.adamw_peer(
            self.peers['grad'].buffer_ptrs, self.peers['data'].buffer_ptrs, self.peers['rank'],
            self.exp_avg, self.exp_avg_sq, self.peers['begin'], self.peers['end'], lr, betas, eps,
            weight_decay, self.steps, self.steps_device)
        self.peers['data'].barrier()

    def adamw(self, lr, betas, eps, weight_decay, grad_scale=1.):
        """torch.optim.AdamW.step over every parameter (one launch)"""
        self.steps += 1
        ops.axpby(1., self.one, 1., self.steps_device)
        ops.adamw(
            self.data, self.grad, self.exp_avg, self.exp_avg_sq, lr, betas, eps, weight_decay,
            self.steps, grad_scale, self.steps_device)
