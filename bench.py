"""Headline benchmark: 22.05 kHz audio samples/s synthesized (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one Generator.forward (promonet/model/generator.py:116-135, HiFi-GAN,
config/promonet.py, random-init seed 1234, fp32) over one batch of 32 synthetic
5 s utterances per GPU (430 frames -> 110 080 samples each): BASELINE.json
configs[1].  N > 1 is launched by torchrun, one rank per GPU; utterances are
independent so ranks share nothing (weak scaling, no data-path collective).

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs
resident in HBM; `e2e` = the same through the host entry point
(pmn_generator_forward_host: pinned host inputs -> H2D -> forward -> D2H audio);
`roofline` = the dominant kernel timed with CUDA events inside this run;
`cpu_baseline` = the CPU oracle (a torch fp32 restatement of the reference
modules, bit-checked against the reference in tests/) on the host cores.

`--impl reference` times that CPU path alone on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# torchrun exports OMP_NUM_THREADS=1; the CPU baseline is entitled to every host core
# and OpenMP sizes its pool at the first parallel region, so set it before torch loads
os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)

BATCH = 32                      # utterances per GPU (north_star)
FRAMES = 430                    # 5 s = 110 250 samples -> 430 frames
HOPSIZE = 256
SAMPLES = FRAMES * HOPSIZE      # 110 080
CPU_SAMPLE_BATCH = 2            # utterances per CPU-baseline step
METRIC = 'audio samples/sec synthesized (22.05 kHz)'
UNIT = 'samples/s'
# conv FLOPs (2 x MAC) per output sample, by kernel
FLOP_PER_SAMPLE_CONV1D = (264.167165952e9 - 8.117e9 - 0.049e9) / SAMPLES
FLOP_PER_SAMPLE_RESBLOCKS = FLOP_PER_SAMPLE_CONV1D - 0.348e9 / SAMPLES
FLOP_PER_SAMPLE_TOTAL = 264.167165952e9 / SAMPLES


def parse():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=10)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--math', default='bf16x3', choices=['fp32', 'bf16x3'])
    parser.add_argument('--no-cpu-baseline', action='store_true')
    return parser.parse_args()


def peaks():
    file = ROOT / 'MEASURED_PEAKS.json'
    if file.exists():
        data = json.loads(file.read_text())
        return {
            'hbm_gbs': data['hbm_gbs'], 'bf16_tflops': data['bf16_tflops'],
            'bf16_tflops_sustained': data.get('bf16_tflops_sustained', data['bf16_tflops']),
            'source': 'measured'}
    # /opt/skills/guides/B200_PROFILING.md fallback
    return {'hbm_gbs': 6650., 'bf16_tflops': 1590., 'bf16_tflops_sustained': 1400.,
            'source': 'fallback'}


###############################################################################
# Clock sampling (nvidia-smi during the timed region)
###############################################################################


class Clocks:

    QUERY = (
        'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
        'clocks_event_reasons.hw_thermal_slowdown,'
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.process = None

    def __enter__(self):
        try:
            self.process = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.process = None
        return self

    def _read(self):
        for line in self.process.stdout:
            self.lines.append(line)

    def __exit__(self, *args):
        if self.process is not None:
            self.process.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, sm_max, reasons = [], [], set()
        for line in self.lines:
            fields = [f.strip() for f in line.split(',')]
            if len(fields) != 6:
                continue
            try:
                sm.append(float(fields[0]))
                sm_max.append(float(fields[1]))
            except ValueError:
                continue
            for name, value in zip(self.NAMES, fields[2:]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {
            'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(sm_max),
            'reasons': sorted(reasons), 'samples': len(sm)}


###############################################################################
# CPU path (the oracle port of the reference modules)
###############################################################################


def cpu_step(state, args):
    import torch
    from oracle import hifigan
    with torch.inference_mode():
        return hifigan.generator(state, *args)


def cpu_baseline(steps, warmup):
    """samples/s of the CPU oracle on all host cores, bounded sample per step"""
    import torch
    from oracle import inputs
    from promonet_b200.model import init
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = init.hifigan_state(1234)
    args = inputs.synthesis(CPU_SAMPLE_BATCH, FRAMES, seed=1234)
    for _ in range(warmup):
        cpu_step(state, args)
    start = time.perf_counter()
    for _ in range(steps):
        cpu_step(state, args)
    elapsed = time.perf_counter() - start
    return {
        'value': CPU_SAMPLE_BATCH * SAMPLES * steps / elapsed,
        'unit': UNIT,
        'cores': cores,
        'kind': 'port',
        'sample': (
            f'{CPU_SAMPLE_BATCH} of the {BATCH} utterances x {FRAMES} frames per step, '
            f'{steps} steps after {warmup} warm-up, torch {torch.__version__} CPU fp32, '
            'oracle/hifigan.py (restatement of promonet/model/{generator,hifigan}.py, '
            'pinned to the reference by tests/golden)'),
    }, elapsed / steps * 1e3


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port)"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    baseline, ms = cpu_baseline(max(1, min(args.steps, 5)), min(args.warmup, 1))
    print(json.dumps({
        'impl': 'reference',
        'metric': METRIC, 'value': baseline['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': baseline,
        'e2e': {'value': baseline['value'], 'unit': UNIT,
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def measured_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the
    committed ncu capture of this same workload (profiles/r1_conv1d_tc_traffic.json)"""
    file = ROOT / 'profiles' / 'r1_conv1d_tc_traffic.json'
    if kernel != 'conv1d_tc_kernel' or not file.exists():
        return None
    return json.loads(file.read_text())['traffic_bytes_per_launch']


def workload_config(gpus):
    return {
        'workload': (
            f'HiFi-GAN Generator.forward, config/promonet.py, random init seed 1234, '
            f'{BATCH} utterances x 5 s ({FRAMES} frames -> {SAMPLES} samples) per GPU, '
            '8-band loudness, fp32 parity'),
        'batch_per_gpu': BATCH, 'frames': FRAMES, 'samples_per_utterance': SAMPLES,
        'global_batch': BATCH * gpus,
        'parallelism': f'utterance-sharded x{gpus}, no collectives',
        'l2': 'per-step working set (>2 GB of activations per GPU) exceeds the 126 MB L2; no flush needed',
    }


###############################################################################
# B200 path
###############################################################################


def run_b200(args):
    import torch
    import promonet_b200
    from promonet_b200 import _lib
    from oracle import inputs  # synthetic inputs only; the oracle is not on this path

    from promonet_b200 import parallel
    rank, local_rank, world = parallel.environment()
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    parallel.initialize('nccl', device)
    if distributed:
        import torch.distributed as dist

    math = _lib.MATH_BF16X3_TC if args.math == 'bf16x3' else _lib.MATH_FP32_SIMT
    state = promonet_b200.model.init.hifigan_state(promonet_b200.RANDOM_SEED)
    model = promonet_b200.model.Generator(device=device, state=state, math=math)
    host = [t.pin_memory() for t in inputs.synthesis(BATCH, FRAMES, seed=1234 + rank)]
    dev = [t.to(device) for t in host]
    audio_host = torch.empty(BATCH, 1, SAMPLES, pin_memory=True)

    barrier = parallel.barrier

    def timed(step, steps):
        """Device time of `steps` calls, max over ranks (ms)"""
        start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
        barrier()
        start.record()
        for _ in range(steps):
            step()
        stop.record()
        barrier()
        return parallel.max_over_ranks(start.elapsed_time(stop), device)

    resident = lambda: model(*dev)
    end_to_end = lambda: model.forward_host(*host, out=audio_host)

    for _ in range(max(args.warmup, 3)):
        resident()
    launches_before = _lib.launch_count()
    with Clocks(local_rank) as clocks:
        ms = timed(resident, args.steps)
    launches = _lib.launch_count() - launches_before

    # End to end: host buffers in, host audio out, copies inside the timed region
    for _ in range(2):
        end_to_end()
    start = time.perf_counter()
    barrier()
    for _ in range(args.steps):
        end_to_end()
    barrier()
    e2e_seconds = parallel.max_over_ranks(time.perf_counter() - start, device)

    # Roofline of the dominant kernel: same steps again with per-launch CUDA events
    dominant = 'conv1d_kernel' if math == _lib.MATH_FP32_SIMT else 'conv1d_tc_kernel'
    _lib.profile(True)
    timed(resident, args.steps)
    kernel_ms, kernel_launches = _lib.profile_read(dominant)
    shares = {}
    for name in ('conv1d_kernel', 'conv1d_tc_kernel', 'conv_transpose1d_kernel',
                 'conv_transpose1d_tc_kernel',
                 'planes_from_f32_kernel', 'zero_plane_pads_kernel',
                 'head_kernel', 'features_kernel', 'speaker_bias_kernel'):
        total, count = _lib.profile_read(name)
        if count:
            shares[name] = {'ms_per_step': total / args.steps, 'launches_per_step': count / args.steps}
    _lib.profile(False)

    if rank == 0:
        peak = peaks()
        total_samples = world * BATCH * SAMPLES * args.steps
        flop_per_sample = (
            FLOP_PER_SAMPLE_CONV1D if math == _lib.MATH_FP32_SIMT else FLOP_PER_SAMPLE_RESBLOCKS)
        flops = BATCH * SAMPLES * flop_per_sample * args.steps
        achieved = flops / (kernel_ms * 1e-3) / 1e12 if kernel_ms else None
        result = {
            'metric': METRIC,
            'value': total_samples / (ms * 1e-3),
            'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32' if math == _lib.MATH_FP32_SIMT else 'f32 (bf16x3 tensor-core products, fp32 accumulate)',
            'data': 'synthetic',
            'config': workload_config(world),
            'clocks': clocks.summary(),
            'gpu_launches': launches,
            'e2e': {
                'value': total_samples / e2e_seconds, 'unit': UNIT,
                'h2d_bytes_per_step': sum(t.numel() * t.element_size() for t in host),
                'd2h_bytes_per_step': audio_host.numel() * 4,
                'api': 'promonet_b200.model.Generator.forward_host -> pmn_generator_forward_host'},
            'roofline': {
                'bound': 'tensor', 'kernel': dominant,
                'achieved': achieved, 'peak': peak['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / peak['bf16_tflops_sustained'] if achieved else None,
                'peak_source': f"{peak['source']} sustained bf16 cuBLAS (kernel timed inside a long step)",
                'traffic': measured_traffic(dominant),
                'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)',
                'launches_per_step': kernel_launches / args.steps,
                'avg_launch_ms': kernel_ms / kernel_launches if kernel_launches else None,
                'flop_per_launch': flops / kernel_launches if kernel_launches else None,
                'how': 'CUDA events around every launch (pmn_profile_*), second pass of the same steps',
                'whole_step_tflops': world * BATCH * SAMPLES * FLOP_PER_SAMPLE_TOTAL * args.steps / (ms * 1e-3) / 1e12,
                'hbm_frac_layer_boundary': (total_samples / world / (ms * 1e-3)) * 15830 / (peak['hbm_gbs'] * 1e9),
                'kernels': shares},
        }
        if not args.no_cpu_baseline and world == 1:
            result['cpu_baseline'], _ = cpu_baseline(2, 1)
        print(json.dumps(result))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
