"""Two-rank probe of the fused peer-memory optimizer step (progress markers per rank)"""
import os, sys, faulthandler
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
import torch.distributed as dist

rank = int(os.environ['RANK'])
faulthandler.dump_traceback_later(40, exit=True)
def mark(text):
    print(f'[rank {rank}] {text}', flush=True)

torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
from promonet_b200.train.params import ParamSet
from promonet_b200.train import ops
state = {'a.weight': torch.randn(1000, 37), 'b.bias': torch.randn(513)}
params = ParamSet(state, f'cuda:{rank}', peer_group=dist.group.WORLD)
mark(f'paramset ok: numel {params.numel} shard {params.peers["begin"]}..{params.peers["end"]}')
dist.broadcast(params.data, 0)
params.grad.copy_(torch.full((params.numel,), float(rank + 1)))
torch.cuda.synchronize()
mark('before adamw_peer')
before = params.data.clone()
params.adamw_peer(2e-4, (.8, .99), 1e-9, .01)
torch.cuda.synchronize()
mark('after adamw_peer')
delta = (params.data - before)
mark(f'update min {float(delta.min()):.3e} max {float(delta.max()):.3e} (expect about -2e-4 everywhere)')
whole = params._whole(params.exp_avg)
mark(f'moments gathered: mean {float(whole.mean()):.4f} (expect 0.2 * 1.5 = 0.3)')
dist.barrier()
mark('done')
os._exit(0)
