"""Feasibility probe: torch symmetric memory (peer pointers over NVLink) on this box"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f'cuda:{rank}')
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(rank, 'ptrs', [hex(p) for p in hdl.buffer_ptrs], 'signal', [hex(p) for p in hdl.signal_pad_ptrs][:2],
      'rank', hdl.rank, hdl.world_size, flush=True)
t.fill_(rank + 1.)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, 'peer value', float(peer[0]), flush=True)
hdl.barrier()
dist.destroy_process_group()
