"""Flat parameter storage for the training step

All parameters of a module live in ONE fp32 device buffer (`data`), with
matching `grad`, `exp_avg` and `exp_avg_sq` buffers, so that the optimizer is one
kernel launch (pmn_adamw) and the data-parallel gradient exchange is one NCCL
all-reduce per module (train/core.py:255,338 are where the reference would
need them).  Names and shapes are those of the reference state dict
(promonet.model.Generator / Discriminator .state_dict()) and the optimizer moments are
saved and loaded in torch.optim.AdamW.state_dict() layout, so checkpoints interchange with
the reference both ways (torchutil.checkpoint format: train/core.py:426-438).
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
            # Symmetric memory needs NVLink peer access between every pair of ranks (and is a
            # private torch API): when any rank cannot set it up, all ranks fall back together
            # to plain buffers, i.e. to the NCCL all-reduce exchange (Trainer.optimize)
            try:
                import torch.distributed._symmetric_memory as symmetric
                data = symmetric.empty(offset, dtype=torch.float32, device=self.device).zero_()
                grad = symmetric.empty(offset, dtype=torch.float32, device=self.device).zero_()
                handles = (symmetric.rendezvous(data, group=peer_group),
                           symmetric.rendezvous(grad, group=peer_group))
                failure = None
            except Exception as error:               # noqa: BLE001 (any failure means "no peers")
                failure = error
            agreed = torch.tensor([0. if failure is None else 1.], device=self.device)
            dist.all_reduce(agreed, group=peer_group)
            if float(agreed) > 0:
                import warnings
                warnings.warn(
                    'promonet_b200.train: peer-memory optimizer step unavailable '
                    f'({failure!r}); using the NCCL all-reduce exchange', RuntimeWarning)
                peer_group = None
        if peer_group is not None:
            self.data, self.grad = data, grad
            self.peers = {
                'data': handles[0],
                'grad': handles[1],
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

    def optimizer_state(self, lr=None, betas=None, eps=None, weight_decay=None):
        """torch.optim.AdamW.state_dict() of the optimizer the reference builds over
        `module.parameters()` (promonet/train/core.py:62-63, saved by torchutil.checkpoint at
        :426-438): per-parameter `exp_avg` / `exp_avg_sq` / `step` keyed by the parameter's
        position, which is the order of the state dict; a reference run can resume from it"""
        from promonet_b200 import config
        exp_avg, exp_avg_sq = self._whole(self.exp_avg).cpu(), self._whole(self.exp_avg_sq).cpu()
        state = {}
        for position, name in enumerate(self.index):
            state[position] = {
                'step': torch.tensor(float(self.steps)),
                'exp_avg': self._view(exp_avg, name).clone(),
                'exp_avg_sq': self._view(exp_avg_sq, name).clone()}
        group = {
            'lr': config.LEARNING_RATE if lr is None else lr,
            'betas': tuple(config.ADAM_BETAS if betas is None else betas),
            'eps': config.ADAM_EPS if eps is None else eps,
            'weight_decay': config.WEIGHT_DECAY if weight_decay is None else weight_decay,
            'amsgrad': False, 'maximize': False, 'foreach': None, 'capturable': False,
            'differentiable': False, 'fused': None, 'decoupled_weight_decay': True,
            'params': list(range(len(self.index)))}
        return {'state': state, 'param_groups': [group]}

    def load_optimizer_state(self, state):
        """Accepts torch.optim.AdamW.state_dict() (a reference checkpoint or ours) and the flat
        layout of this library's first checkpoints; anything else raises instead of silently
        restarting the moments and the bias-correction step"""
        if 'exp_avg' in state and 'state' not in state:         # flat buffers (round-1 files)
            self.exp_avg.copy_(state['exp_avg'])
            self.exp_avg_sq.copy_(state['exp_avg_sq'])
            self.steps = int(state['step'])
        elif 'state' in state and 'param_groups' in state:
            positions = state['param_groups'][0]['params']
            names = list(self.index)
            if len(positions) != len(names):
                raise ValueError(
                    f'optimizer state has {len(positions)} parameters, the module {len(names)}')
            if not state['state']:
                return                                          # an optimizer that never stepped
            exp_avg, exp_avg_sq = torch.zeros(self.numel), torch.zeros(self.numel)
            steps = 0
            for position, name in zip(positions, names):
                entry = state['state'][position]
                if tuple(entry['exp_avg'].shape) != self.index[name][1]:
                    raise ValueError(f'optimizer state of {name} has shape {tuple(entry["exp_avg"].shape)}')
                self._view(exp_avg, name).copy_(entry['exp_avg'])
                self._view(exp_avg_sq, name).copy_(entry['exp_avg_sq'])
                steps = max(steps, int(entry['step']))
            self.exp_avg.copy_(exp_avg)
            self.exp_avg_sq.copy_(exp_avg_sq)
            self.steps = steps
        else:
            raise ValueError(
                'unrecognised optimizer state: expected torch.optim.AdamW.state_dict() '
                f'(keys state / param_groups), got keys {sorted(state)[:6]}')
        self.steps_device.fill_(float(self.steps))

    def release_peers(self):
        """Drop the symmetric-memory mappings (before torch.distributed.destroy_process_group, which
        can otherwise wait on the peers): the parameters and gradients move to plain device buffers
        and the module falls back to the NCCL exchange"""
        if self.peers is None:
            return
        data, grad = self.data.clone(), self.grad.clone()
        torch.cuda.synchronize(self.device)
        self.peers = None
        self.data, self.grad = data, grad

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
