"""Does destroy_process_group() return when the symmetric-memory buffers are released first?

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 profiles/debug/destroy_probe.py [release]

Builds a 2-rank Trainer with the peer-memory optimizer, takes one optimizer step, then tears down: with
`release`, drops every reference to the symmetric tensors / handles (Trainer.close) before
the barrier and destroy_process_group; a watchdog thread reports a hang after 60 s and exits.
"""
import os
import sys
import threading
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import promonet_b200  # noqa: E402
from promonet_b200.train import Trainer  # noqa: E402


def main():
    release = len(sys.argv) > 1 and sys.argv[1] == 'release'
    rank = int(os.environ['RANK'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl')
    trainer = Trainer(device=f'cuda:{torch.cuda.current_device()}')
    assert trainer.generator.params.peers is not None, 'no peer memory on this box'
    trainer.optimize(trainer.generator.params)       # one fused exchange + AdamW over the peers
    torch.cuda.synchronize()

    def watchdog():
        time.sleep(60)
        print(f'rank {rank}: teardown HUNG (release={release})', flush=True)
        os._exit(3)
    threading.Thread(target=watchdog, daemon=True).start()
    if release:
        trainer.close()
    dist.barrier()
    started = time.perf_counter()
    dist.destroy_process_group()
    print(f'rank {rank}: destroy_process_group returned in {time.perf_counter() - started:.2f} s (release={release})',
          flush=True)


if __name__ == '__main__':
    main()
