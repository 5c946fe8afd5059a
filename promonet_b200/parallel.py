"""Multi-GPU host logic: one process per GPU, utterances sharded with no data-path
collective (synthesis and preprocessing shard over independent utterances);
training replicas exchange gradients with one all-reduce per optimizer.

torch.distributed is plumbing here: rendezvous, a barrier around timed regions and a
MAX reduction of per-rank device times.  Works on NCCL (GPU ranks) and gloo (CPU
tests of this logic)."""
import os

import torch


def environment():
    """(rank, local_rank, world_size) from the torchrun environment"""
    return (
        int(os.environ.get('RANK', '0')),
        int(os.environ.get('LOCAL_RANK', '0')),
        int(os.environ.get('WORLD_SIZE', '1')))


def initialize(backend=None, device=None):
    """Join the process group when launched by torchrun with more than one rank"""
    import torch.distributed as dist
    rank, _, world = environment()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl' and device is not None:
            kwargs['device_id'] = device
        dist.init_process_group(backend, **kwargs)
    return rank, world


def shard(count, rank, world):
    """Contiguous block of `count` utterances owned by `rank`: sizes differ by at
    most one and the blocks tile range(count) in rank order"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank {rank} of {world}')
    base, extra = divmod(count, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owns(index, rank, world):
    """Round-robin ownership of the items of a stream whose length is not known in advance
    (validation loaders): item i belongs to rank i mod world"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank {rank} of {world}')
    return index % world == rank


def shard_tensors(tensors, rank, world):
    """Slice every (B, ...) tensor to this rank's utterances"""
    block = shard(tensors[0].shape[0], rank, world)
    return [t[block.start:block.stop] for t in tensors]


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value, device='cpu'):
    """Slowest rank's time: the denominator of every multi-GPU throughput"""
    import torch.distributed as dist
    tensor = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
    return float(tensor)


def sum_over_ranks(value, device='cpu'):
    """Units processed by the whole job"""
    import torch.distributed as dist
    tensor = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return float(tensor)


def all_reduce_sum(flat, group=None):
    """Sum a module's flat gradient buffer over the data-parallel ranks in place (the
    exchange step of the training path: one collective per optimizer); the mean
    is taken by the optimizer kernel's grad_scale = 1 / world size"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def gather_utterances(local, count, rank, world, device='cpu'):
    """Reassemble per-rank results (B_rank, ...) in utterance order on every rank;
    only used when a caller wants the whole batch back (not on the timed path)"""
    import torch.distributed as dist
    if world == 1:
        return local
    sizes = [len(shard(count, r, world)) for r in range(world)]
    pieces = [
        torch.empty((size, *local.shape[1:]), dtype=local.dtype, device=device)
        for size in sizes]
    dist.all_gather(pieces, local.to(device).contiguous()) if len(set(sizes)) == 1 else \
        _all_gather_ragged(pieces, local.to(device).contiguous(), rank, world)
    return torch.cat(pieces, dim=0)


def _all_gather_ragged(pieces, local, rank, world):
    import torch.distributed as dist
    for source in range(world):
        if source == rank:
            pieces[source].copy_(local)
        dist.broadcast(pieces[source], src=source)


def shutdown(*trainers, grace=30.):
    """Orderly end of a multi-rank job: close the trainers (their peer-memory mappings), meet at a
    barrier and destroy the process group.  torch's teardown has been seen to wait on the peers
    with live symmetric-memory mappings; should it still be stuck after `grace` seconds, a watchdog
    ends the process (everything has been printed and synchronised by then)."""
    import sys
    import threading
    import time
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return
    for trainer in trainers:
        trainer.close()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    torch.distributed.barrier()
    sys.stdout.flush()
    sys.stderr.flush()

    def watchdog():
        time.sleep(grace)
        sys.stderr.write('promonet_b200.parallel.shutdown: destroy_process_group did not return; exiting\n')
        sys.stderr.flush()
        os._exit(0)
    threading.Thread(target=watchdog, daemon=True).start()
    torch.distributed.destroy_process_group()
