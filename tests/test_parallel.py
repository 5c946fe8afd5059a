"""N > 1 host logic on CPU: two gloo ranks exercise sharding, the MAX-over-ranks
timing reduction and result reassembly (the GPU ranks use the same code over NCCL)"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from promonet_b200 import parallel


def test_shard_tiles_the_batch():
    for count in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard(count, r, world) for r in range(world)]
            assert [i for b in blocks for i in b] == list(range(count))
            sizes = [len(b) for b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, count, queue):
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    got_rank, got_world = parallel.initialize('gloo')
    assert (got_rank, got_world) == (rank, world)
    utterances = torch.arange(count * 3, dtype=torch.float32).reshape(count, 3)
    mine, = parallel.shard_tensors([utterances], rank, world)
    processed = mine * 2.                       # stands in for the per-rank synthesis
    parallel.barrier()
    slowest = parallel.max_over_ranks(10. + rank)          # rank 1 is the slow one
    total = parallel.sum_over_ranks(mine.shape[0])
    whole = parallel.gather_utterances(processed, count, rank, world)
    queue.put((rank, slowest, total, whole))
    parallel.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('count', [8, 5])
def test_two_rank_sharding_over_gloo(count):
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_worker, args=(rank, 2, port, count, queue)) for rank in range(2)]
    for worker in workers:
        worker.start()
    results = [queue.get(timeout=120) for _ in workers]
    for worker in workers:
        worker.join(timeout=60)
        assert worker.exitcode == 0
    expected = torch.arange(count * 3, dtype=torch.float32).reshape(count, 3) * 2.
    for rank, slowest, total, whole in results:
        assert slowest == 11.      # max over ranks, not this rank's own time
        assert total == count      # every utterance processed exactly once
        assert torch.equal(whole, expected)
