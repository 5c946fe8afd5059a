"""Data-parallel training step over NCCL: two ranks, each with half of a batch, must end the
step with the parameters a single GPU reaches on the whole batch (the losses are batch means,
so averaging the rank gradients is the full-batch gradient).  Needs two GPUs."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, queue, peer_optimizer):
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from oracle import train as oracle_train
    from promonet_b200 import parallel
    from promonet_b200.model import init
    from promonet_b200.train.core import Trainer
    torch.cuda.set_device(rank)
    device = torch.device('cuda', rank)
    parallel.initialize('nccl', device)
    states = init.hifigan_state(1234), init.discriminator_state(1234)
    batch = oracle_train.batch(4, 8, seed=41)
    trainer = Trainer(*states, device=device, math='fp32', peer_optimizer=peer_optimizer)
    assert trainer.world == world
    assert (trainer.generator.params.peers is not None) == peer_optimizer
    trainer.broadcast_parameters()
    mine = [t.to(device).contiguous() for t in parallel.shard_tensors(list(batch), rank, world)]
    for _ in range(2):
        trainer.step(*mine)
    result = None
    if rank == 0:
        torch.distributed.barrier()
        # single-process yardstick on the same device, outside the process group's reach
        single = Trainer(*states, device=device, math='fp32', data_parallel=False)
        whole = [t.to(device).contiguous() for t in batch]
        for _ in range(2):
            single.step(*whole)
        result = []
        for a, b in ((trainer.generator, single.generator),
                     (trainer.discriminators, single.discriminators)):
            difference = (a.params.data - b.params.data).abs()
            moved = (b.params.data - b.params.data.new_tensor(0.)).abs().max()
            result.append((float(difference.max()), float((difference > 1e-5).float().mean()),
                           float(moved)))
    else:
        torch.distributed.barrier()
    queue.put((rank, result))
    # orderly teardown: peer mappings released, barrier, destroy_process_group (the helper ends
    # the process should torch's teardown still wait on the peer)
    parallel.shutdown(trainer)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('peer_optimizer', [True, False])
def test_two_rank_step_matches_single_gpu_full_batch(peer_optimizer):
    """peer_optimizer=True: gradient exchange fused with AdamW in one kernel over NVLink peer
    memory (pmn_adamw_peer); False: NCCL all-reduce + local AdamW"""
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_worker, args=(rank, 2, port, queue, peer_optimizer))
        for rank in range(2)]
    for worker in workers:
        worker.start()
    results = dict(queue.get(timeout=900) for _ in workers)
    for worker in workers:
        worker.join(timeout=120)
        assert worker.exitcode == 0
    for largest, fraction, _ in results[0]:
        # two AdamW steps move a parameter by at most 2 lr = 4e-4; sign flips of near-zero
        # gradients (different summation order across ranks) may move a few by that much
        assert largest <= 2 * 2e-4 * 1.01
        assert fraction < 2e-2


def _validation_loader(items, frames):
    from oracle import inputs
    loader = []
    for index in range(items):
        loudness, pitch, periodicity, ppg, speakers, _, _ = inputs.synthesis(1, frames + index, seed=60 + index)
        loader.append((
            None, loudness, pitch, periodicity * .5, ppg, speakers, None, None, None,
            torch.zeros(1, 1, (frames + index) * 256), None))
    return loader


def _validation_worker(rank, world, port, queue):
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import promonet_b200
    from promonet_b200 import parallel
    from promonet_b200.model import init
    from promonet_b200.train import evaluate
    torch.cuda.set_device(rank)
    device = torch.device('cuda', rank)
    parallel.initialize('nccl', device)
    generator = promonet_b200.model.Generator(device=device, state=init.hifigan_state(1234))
    loader = _validation_loader(3, 20)
    scalars, waveforms = evaluate(None, 1, generator, loader, rank)
    alone, _ = evaluate(None, 1, generator, loader, rank, data_parallel=False)
    queue.put((rank, scalars, alone, sorted(waveforms)))
    parallel.shutdown()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_validation_equals_single_gpu_validation():
    """Items dealt round-robin to two ranks + one all-reduce of the sums = the scalars one GPU
    computes over the whole loader"""
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_validation_worker, args=(rank, 2, port, queue)) for rank in range(2)]
    for worker in workers:
        worker.start()
    results = {rank: rest for rank, *rest in (queue.get(timeout=600) for _ in workers)}
    for worker in workers:
        worker.join(timeout=120)
        assert worker.exitcode == 0
    assert results[0][0].keys() == results[1][0].keys() == results[0][1].keys()
    for name, value in results[0][1].items():
        # the same scalars on both ranks, equal to what one GPU computes over the whole loader
        assert results[0][0][name] == pytest.approx(results[1][0][name], rel=1e-12, nan_ok=True), name
        assert results[0][0][name] == pytest.approx(value, rel=1e-9, nan_ok=True), name
    assert len(results[0][2]) == 2 * 7 and len(results[1][2]) == 7    # items 0, 2 and item 1
