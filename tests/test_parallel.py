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


def _gradient_worker(rank, world, port, queue):
    """Each rank differentiates the LSGAN discriminator loss (oracle, CPU autograd) on its
    shard of the batch, packs the gradients into the flat buffer the trainer all-reduces, and
    averages them the way the optimizer kernel does (grad_scale = 1 / world)"""
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    parallel.initialize('gloo')
    torch.set_num_threads(2)
    from oracle import train as oracle_train
    from promonet_b200.model import init
    from promonet_b200.train.params import ParamSet
    state = oracle_train.leaf_state(init.discriminator_state(1234))
    torch.manual_seed(7)
    audio = .3 * torch.randn(2, 1, 2048)
    generated = .3 * torch.randn(2, 1, 2048)

    def gradients(real, fake):
        for p in oracle_train.parameters(state).values():
            p.grad = None
        real_logits, fake_logits, _, _ = oracle_train.discriminator(state, real, fake)
        oracle_train.discriminator_loss(real_logits, fake_logits).backward()
        return {k: v.grad.clone() for k, v in oracle_train.parameters(state).items()}

    whole = gradients(audio, generated)
    mine = gradients(*parallel.shard_tensors([audio, generated], rank, world))
    params = ParamSet({k: v.detach() for k, v in state.items()}, 'cpu')
    for name, value in mine.items():
        params.gradient(name).copy_(value)
    parallel.all_reduce_sum(params.grad)
    params.grad.mul_(1. / world)
    worst = max(
        float((params.gradient(k) - v).abs().max() / v.abs().max().clamp_min(1e-30))
        for k, v in whole.items())
    queue.put((rank, worst))
    parallel.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_average_equals_the_full_batch_gradient():
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_gradient_worker, args=(rank, 2, port, queue)) for rank in range(2)]
    for worker in workers:
        worker.start()
    results = [queue.get(timeout=600) for _ in workers]
    for worker in workers:
        worker.join(timeout=60)
        assert worker.exitcode == 0
    for rank, worst in results:
        assert worst < 1e-4, (rank, worst)


###############################################################################
# Data-parallel validation: items dealt round-robin, one all-reduce of the sums
###############################################################################


def _oracle_sums(metrics):
    """The 12 running sums of pmn_metrics_update (include/promonet_b200.h) out of the oracle"""
    pairs = (
        metrics.loudness.both, metrics.loudness.loud, metrics.loudness.quiet, metrics.periodicity,
        metrics.pitch, metrics.ppg)
    return torch.tensor(
        [value for metric in pairs for value in (metric.total, metric.count)], dtype=torch.float64)


def _validation_items(count):
    from oracle import make_golden
    return [
        (make_golden.metric_inputs(50 + i, 20 + 7 * i, 8), make_golden.metric_inputs(70 + i, 20 + 7 * i, 8))
        for i in range(count)]


def _validation_worker(rank, world, port, count, queue):
    from oracle import metrics as oracle_metrics
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    parallel.initialize('gloo')
    mine = oracle_metrics.Metrics()
    for index, (predicted, target) in enumerate(_validation_items(count)):
        if parallel.owns(index, rank, world):
            mine.update(*predicted, *target)
    table = _oracle_sums(mine)[None].repeat(7, 1)        # the (conditions, slots) table of evaluate
    parallel.all_reduce_sum(table)
    queue.put((rank, table))
    parallel.barrier()
    torch.distributed.destroy_process_group()


def test_round_robin_ownership():
    for world in (1, 2, 3, 8):
        owners = [[r for r in range(world) if parallel.owns(i, r, world)] for i in range(20)]
        assert owners == [[i % world] for i in range(20)]
    with pytest.raises(ValueError):
        parallel.owns(0, 2, 2)


def test_two_rank_validation_sums_equal_the_single_process_metrics():
    """evaluate() under data parallelism: per-rank sums of the round-robin items, all-reduced,
    give the scalars promonet.evaluate.Metrics gives over all items in one process"""
    from oracle import metrics as oracle_metrics
    from promonet_b200.evaluate import metrics as product_metrics
    count = 5
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_validation_worker, args=(rank, 2, port, count, queue))
        for rank in range(2)]
    for worker in workers:
        worker.start()
    results = dict(queue.get(timeout=120) for _ in workers)
    for worker in workers:
        worker.join(timeout=60)
        assert worker.exitcode == 0
    whole = oracle_metrics.Metrics()
    for predicted, target in _validation_items(count):
        whole.update(*predicted, *target)
    expected = whole()
    for rank in range(2):
        assert torch.equal(results[rank], results[0])
        scalars = product_metrics.finish(results[rank][3].tolist())
        assert list(scalars) == ['pitch', 'periodicity', 'ppg', 'loudness', 'loudness-loud', 'loudness-quiet']
        for name, value in expected.items():
            assert scalars[name] == pytest.approx(value, rel=1e-12), name


class _Closable:
    """Stands in for a Trainer: shutdown must close it before the process group goes away"""

    def __init__(self):
        self.closed_while_initialized = None

    def close(self):
        self.closed_while_initialized = torch.distributed.is_initialized()


def _shutdown_worker(rank, world, port, queue):
    os.environ.update(
        RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
        MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    parallel.initialize('gloo')
    trainer = _Closable()
    parallel.shutdown(trainer, grace=60.)
    queue.put((rank, trainer.closed_while_initialized, torch.distributed.is_initialized()))


def test_shutdown_closes_trainers_then_destroys_the_process_group():
    """parallel.shutdown: close, barrier, destroy_process_group (the multi-GPU benchmark and the
    data-parallel tests end through it); a no-op without a process group"""
    parallel.shutdown()
    context = mp.get_context('spawn')
    queue = context.Queue()
    port = _free_port()
    workers = [
        context.Process(target=_shutdown_worker, args=(rank, 2, port, queue)) for rank in range(2)]
    for worker in workers:
        worker.start()
    results = sorted(queue.get(timeout=120) for _ in workers)
    for worker in workers:
        worker.join(timeout=60)
        assert worker.exitcode == 0
    assert results == [(0, True, False), (1, True, False)]
