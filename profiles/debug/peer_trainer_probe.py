"""Two-rank probe of Trainer with the fused peer-memory optimizer (progress markers per rank)"""
import os, sys, faulthandler, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
import torch.distributed as dist

rank = int(os.environ['RANK'])
faulthandler.dump_traceback_later(50, exit=True)
t0 = time.time()
def mark(text):
    print(f'[rank {rank} {time.time() - t0:5.1f}s] {text}', flush=True)

torch.cuda.set_device(rank)
device = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=device)
from oracle import train as oracle_train
from promonet_b200 import parallel
from promonet_b200.model import init
from promonet_b200.train.core import Trainer
states = init.hifigan_state(1234), init.discriminator_state(1234)
mark('states built')
trainer = Trainer(*states, device=device, math='tf32', peer_optimizer=True)
mark('trainer built')
trainer.broadcast_parameters()
torch.cuda.synchronize()
mark('broadcast done')
batch = [t.to(device).contiguous() for t in parallel.shard_tensors(list(oracle_train.batch(4, 8, seed=41)), rank, 2)]
for i in range(2):
    trainer._discriminator_phase(batch)
    torch.cuda.synchronize(); mark(f'step {i}: discriminator phase')
    trainer.optimize(trainer.discriminators.params)
    torch.cuda.synchronize(); mark(f'step {i}: D optimize')
    trainer._generator_phase(batch)
    torch.cuda.synchronize(); mark(f'step {i}: generator phase')
    trainer.optimize(trainer.generator.params)
    torch.cuda.synchronize(); mark(f'step {i}: G optimize')
    losses = trainer._final_phase()
    mark(f'step {i}: losses {[round(v, 4) for v in losses.tolist()]}')
checksum = float(trainer.generator.params.data.double().abs().sum())
mark(f'G checksum {checksum:.6f}')
trainer.save(Path(os.environ.get('GRAFT_REPO_ROOT', '.')) / 'gpurun_out' / 'peer_ckpt')
mark('saved')
dist.barrier()
mark('done')
os._exit(0)
