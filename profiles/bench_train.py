"""Training step benchmark: BASELINE.json configs[3] (one training step, config/promonet.py:
Generator + MPD/CMB discriminators + mel loss, data-parallel, synthetic VCTK-shape batches).

    python profiles/bench_train.py [--batch 8] [--steps 5] [--warmup 2]
    torchrun --nproc-per-node N profiles/bench_train.py ...      (DDP over NCCL)

One step = promonet_b200.train.Trainer.step (train/core.py:183-369): generator forward,
discriminator step, generator step, both AdamW updates, gradient all-reduce when N > 1.
Per-GPU batch 8 x 16 384 samples (global 64 at 8 GPUs, config/defaults.py:367-370).
Prints one JSON line: items/s, ms per step (max over ranks), per-kernel device times,
TFLOP/s against the fp32 FMA pipe, and the CPU oracle (oracle/train.py) on 1 item.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200 import _lib, parallel  # noqa: E402
from promonet_b200.model import init  # noqa: E402
from promonet_b200.train.core import Trainer  # noqa: E402
from oracle import train as oracle_train  # noqa: E402

KERNELS = (
    'conv_fprop_tc_kernel', 'conv_dgrad_tc_kernel', 'conv_wgrad_tc_kernel', 'pack_weight_taps_kernel',
    'fold_weights_kernel', 'pack_weights_kernel',
    'conv_fprop_kernel', 'conv_dgrad_kernel', 'conv_wgrad_kernel', 'conv_transpose1d_kernel',
    'weight_norm_fold_kernel', 'weight_norm_backward_kernel', 'weight_norm_backward_table_kernel', 'transpose_weight_kernel',
    'stft_train_kernel', 'stft_train_backward_kernel', 'mel_loss_kernel', 'mel_kernel',
    'l1_mean_kernel', 'mse_to_target_kernel', 'axpby_kernel', 'adamw_kernel', 'adamw_peer_kernel',
    'reflect_pad_kernel', 'reflect_pad_backward_kernel', 'copy_columns_kernel',
    'channel_sum_kernel', 'row_sum_kernel', 'features_kernel', 'embedding_backward_kernel',
    'pitch_bins_kernel', 'global_features_kernel')
# conv FLOPs (2 x MAC) per item of 16 384 samples
G_FORWARD = 39.3e9
D_FORWARD_PAIR = 40.5e9            # real + fake
# G fwd + (D fwd + D bwd: dgrad + wgrad) + (D fwd + D dgrad on the fake half) + G bwd (dgrad + wgrad)
FLOP_PER_ITEM = G_FORWARD + 3 * D_FORWARD_PAIR + (D_FORWARD_PAIR + D_FORWARD_PAIR / 2) + 2 * G_FORWARD


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--batch', type=int, default=8)
    parser.add_argument('--frames', type=int, default=64)
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--warmup', type=int, default=2)
    parser.add_argument('--no-cpu', action='store_true')
    parser.add_argument('--math', default='tf32', choices=['tf32', 'fp32'])
    parser.add_argument('--nccl', action='store_true', help='N > 1: NCCL all-reduce + AdamW instead of the fused peer-memory optimizer step')
    parser.add_argument('--eager', action='store_true', help='launch every kernel from Python instead of replaying CUDA graphs')
    args = parser.parse_args()
    rank, local_rank, world = parallel.environment()
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    parallel.initialize('nccl', device)
    trainer = Trainer(
        init.hifigan_state(1234), init.discriminator_state(1234), device, math=args.math,
        peer_optimizer=not args.nccl)
    trainer.broadcast_parameters()
    batch = [t.to(device).contiguous() for t in oracle_train.batch(args.batch, args.frames, 1234 + rank)]
    run = trainer.step if args.eager else trainer.step_graphed
    for _ in range(args.warmup):
        run(*batch)
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    parallel.barrier()
    before = _lib.launch_count()
    start.record()
    for _ in range(args.steps):
        losses = run(*batch)
    stop.record()
    parallel.barrier()
    launches = (_lib.launch_count() - before) / args.steps
    ms = parallel.max_over_ranks(start.elapsed_time(stop), device) / args.steps
    _lib.profile(True)
    trainer.step(*batch)
    torch.cuda.synchronize()
    kernels = {}
    for name in KERNELS:
        total, count = _lib.profile_read(name)
        if count:
            kernels[name] = {'ms': round(total, 3), 'launches': count}
    _lib.profile(False)
    if rank == 0:
        items = world * args.batch
        result = {
            'metric': 'training items/sec (16 384-sample chunks)', 'value': items / (ms * 1e-3),
            'unit': 'items/s', 'n_gpus': world, 'ms_per_step': ms, 'batch_per_gpu': args.batch,
            'global_batch': items, 'frames': args.frames,
            'dtype': 'f32 (tf32 tensor-core products, fp32 accumulate)' if args.math == 'tf32' else 'f32',
            'gpu_launches_per_step': launches if args.eager else None,
            'launch': 'eager' if args.eager else 'three CUDA graphs per step, optimizer steps between them',
            'exchange': None if world == 1 else (
                'NCCL all-reduce + AdamW kernel' if args.nccl else
                'pmn_adamw_peer: reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory'),
            'losses': dict(zip(('discriminator', 'mel', 'feature_matching', 'adversarial', 'generator'),
                               [float(v) for v in losses.cpu()])),
            'tflops': args.batch * FLOP_PER_ITEM * (args.frames / 64) / (ms * 1e-3) / 1e12,
            'flop_per_item': FLOP_PER_ITEM,
            'kernels': kernels,
            'kernel_ms_total': round(sum(k['ms'] for k in kernels.values()), 3)}
        if not args.no_cpu and world == 1:
            torch.set_num_threads(os.cpu_count())
            cpu_batch = oracle_train.batch(1, args.frames, 1234)
            g = oracle_train.leaf_state(init.hifigan_state(1234))
            d = oracle_train.leaf_state(init.discriminator_state(1234))
            optimizers = oracle_train.make_optimizers(g, d)
            begin = time.perf_counter()
            oracle_train.step(g, d, cpu_batch, optimizers)
            seconds = time.perf_counter() - begin
            result['cpu_baseline'] = {
                'value': 1 / seconds, 'unit': 'items/s', 'cores': os.cpu_count(), 'kind': 'port',
                'sample': '1 item through oracle/train.py (torch autograd fp32 + AdamW), one step'}
        print(json.dumps(result))
    if world > 1:
        from promonet_b200 import parallel
        parallel.shutdown(trainer)


if __name__ == '__main__':
    main()
