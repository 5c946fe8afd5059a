"""Benchmark of the ProMoNet hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline (`value`): 22.05 kHz audio samples/s synthesized.  A step = one
Generator.forward (promonet/model/generator.py:116-135, HiFi-GAN,
config/promonet.py, random-init seed 1234, fp32 parity) over one batch of 32
synthetic 5 s utterances per GPU (430 frames -> 110 080 samples each):
BASELINE.json configs[1].  N > 1 is launched by torchrun, one rank per GPU;
utterances are independent so ranks share nothing (weak scaling, no data-path
collective).

ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in
HBM; `e2e` = the same through the host entry point (pinned host inputs -> H2D ->
forward -> D2H audio, copies inside the timed region); `roofline` = the dominant
kernel timed with CUDA events inside this run; `cpu_baseline` = the reference's
own modules on the host cores (oracle/_ref, the unmodified package behind stub
imports; the oracle port when that copy is absent); `parity` = this run's own
output against that CPU result on the sampled utterances.

`secondary` carries the other BASELINE.json configs, timed in the same run at the
same N (each with its own roofline / cpu_baseline / parity):
  preprocess  configs[2]: promonet.preprocess.from_audio chain, 32 x 10 s per GPU
              -> frames/s (the second half of BASELINE.json's metric)
  train       configs[3]: one training step, 8 items x 16 384 samples per GPU,
              gradient exchange over NVLink when N > 1 -> items/s
  fargan      configs[4]: FARGAN generator, 32 x 5 s per GPU (and 4 per GPU)
  plumbing    configs[0]: promonet.synthesize.from_features, 1 x 1 s
`gpu_eager_baseline` = the reference modules run by PyTorch eager ON the same
B200 (cuDNN; strict fp32, TF32, bf16 autocast): the bar to beat on that box.

`--impl reference` times the reference's CPU implementation of the headline
path alone: K steps of a bounded sample (2 utterances) after W warm-up steps.
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
BYTES_PER_SAMPLE = 15830        # layer-boundary bytes, SURVEY 8d
PARITY_BAR = 1e-4               # max|a - b| / max|b| per utterance (north_star)

PRE_BATCH = 32                  # configs[2]: 256 utterances x 10 s over 8 GPUs
PRE_SAMPLES = 220500
PRE_FRAMES = PRE_SAMPLES // HOPSIZE
PRE_FLOP_PER_FRAME = 394.6e6    # FCNF0++
PRE_BYTES_PER_FRAME = 15.8e3
TRAIN_BATCH = 8                 # configs[3]: global batch 64 over 8 GPUs
TRAIN_FRAMES = 64
G_FORWARD = 39.3e9              # conv FLOPs per 16 384-sample item
D_FORWARD_PAIR = 40.5e9
TRAIN_FLOP_PER_ITEM = (
    G_FORWARD + 3 * D_FORWARD_PAIR + (D_FORWARD_PAIR + D_FORWARD_PAIR / 2) + 2 * G_FORWARD)
FARGAN_FLOP_PER_SAMPLE = 73.8e3


def parse():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=10)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--math', default='bf16x3', choices=['fp32', 'bf16x3'])
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-secondary', action='store_true')
    parser.add_argument('--only', default=None, help='comma list of secondary workloads to run')
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
# CPU legs: the reference (or its oracle port) on the host cores.  The only
# places this file touches oracle/: timed as the baseline, and its output reused
# as the parity check of the B200 result on the same sampled inputs.
###############################################################################


def relative_error(actual, expected):
    """max|a - b| / max|b| (the parity measure of north_star)"""
    return float((actual.double() - expected.double()).abs().max() / expected.double().abs().max())


def reference_generator():
    """(callable(args) -> audio, kind, description) for the headline path on CPU"""
    import torch
    from promonet_b200.model import init
    try:
        from oracle import reference
        promonet = reference.load()
    except Exception as error:       # the vendored tree is optional; the port always exists
        promonet, why = None, repr(error)[:120]
    if promonet is not None:
        model = reference.generator(promonet, 1234)
        return (
            lambda args: reference.forward(model, *args), 'reference',
            f'promonet.model.Generator (unmodified reference package at {reference.root()}, '
            'imported behind stub third-party modules), seed 1234, eval, no autocast')
    from oracle import hifigan
    state = init.hifigan_state(1234)
    return (
        lambda args: hifigan.generator(state, *args), 'port',
        'oracle/hifigan.py (restatement of promonet/model/{generator,hifigan}.py, pinned to '
        'the reference by tests/golden; the reference tree did not travel to this box)')


def cpu_synthesis(steps, warmup, seed=1234):
    """samples/s of the reference Generator.forward on all host cores over a bounded
    sample; returns (cpu_baseline dict, ms per step, last output, its inputs)"""
    import torch
    from promonet_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    forward, kind, what = reference_generator()
    args = synthetic.synthesis(BATCH, FRAMES, seed=seed)
    args = [t[:CPU_SAMPLE_BATCH] for t in args]
    with torch.inference_mode():
        for _ in range(warmup):
            forward(args)
        start = time.perf_counter()
        for _ in range(steps):
            audio = forward(args)
        elapsed = time.perf_counter() - start
    return {
        'value': CPU_SAMPLE_BATCH * SAMPLES * steps / elapsed,
        'unit': UNIT,
        'cores': cores,
        'kind': kind,
        'sample': (
            f'the first {CPU_SAMPLE_BATCH} of the {BATCH} utterances x {FRAMES} frames per step, '
            f'{steps} steps after {warmup} warm-up, torch {torch.__version__} CPU fp32, {what}'),
    }, elapsed / steps * 1e3, audio, args


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the headline path,
    exactly --steps steps of the bounded sample after --warmup warm-up steps"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    baseline, ms, _, _ = cpu_synthesis(max(1, args.steps), max(0, args.warmup))
    config = workload_config(args.gpus)
    config['sample'] = (
        f'each step = {CPU_SAMPLE_BATCH} of the {BATCH} utterances (bounded CPU sample); '
        'value = samples synthesized / time, so it is comparable to the full-batch B200 line')
    print(json.dumps({
        'impl': 'reference',
        'metric': METRIC, 'value': baseline['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': max(1, args.steps), 'warmup': max(0, args.warmup),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config,
        'cpu_baseline': baseline,
        'e2e': {'value': baseline['value'], 'unit': UNIT,
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def measured_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the
    newest committed ncu capture of this same workload (profiles/r*_conv1d_tc_traffic.json)"""
    if kernel != 'conv1d_tc_kernel':
        return None, None
    files = sorted((ROOT / 'profiles').glob('r*_conv1d_tc_traffic.json'))
    if not files:
        return None, None
    return json.loads(files[-1].read_text())['traffic_bytes_per_launch'], files[-1].name


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


class Context:
    """Rank, device and the timing helpers every workload shares"""

    def __init__(self, args):
        import torch
        from promonet_b200 import parallel
        self.args = args
        self.torch = torch
        self.parallel = parallel
        self.rank, self.local_rank, self.world = parallel.environment()
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device('cuda', self.local_rank)
        parallel.initialize('nccl', self.device)
        self.cpu = not args.no_cpu_baseline and self.world == 1 and self.rank == 0
        self.peak = peaks()

    def timed(self, step, steps):
        """Device time of `steps` calls on the current stream, max over ranks (ms)"""
        torch = self.torch
        start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
        self.parallel.barrier()
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            step()
        stop.record()
        torch.cuda.synchronize()
        self.parallel.barrier()
        return self.parallel.max_over_ranks(start.elapsed_time(stop), self.device)

    def wall(self, step, steps):
        """Wall-clock seconds of `steps` host-synchronous calls, max over ranks; the
        barriers sit outside the timed region"""
        self.parallel.barrier()
        self.torch.cuda.synchronize()
        start = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        seconds = time.perf_counter() - start
        self.parallel.barrier()
        return self.parallel.max_over_ranks(seconds, self.device)

    def kernels(self, step, names, steps=1):
        """Per-kernel device time of `steps` more calls (CUDA events around every launch)"""
        from promonet_b200 import _lib
        _lib.profile(True)
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        table = {}
        for name in names:
            total, count = _lib.profile_read(name)
            if count:
                table[name] = {
                    'ms_per_step': round(total / steps, 4), 'launches_per_step': count / steps}
        _lib.profile(False)
        return table


def synthesis(ctx):
    """configs[1], the headline"""
    import promonet_b200
    from promonet_b200 import _lib, synthetic
    torch, args = ctx.torch, ctx.args
    math = _lib.MATH_BF16X3_TC if args.math == 'bf16x3' else _lib.MATH_FP32_SIMT
    state = promonet_b200.model.init.hifigan_state(promonet_b200.RANDOM_SEED)
    model = promonet_b200.model.Generator(device=ctx.device, state=state, math=math)
    seed = 1234 + ctx.rank
    host = [t.pin_memory() for t in synthetic.synthesis(BATCH, FRAMES, seed=seed)]
    dev = [t.to(ctx.device) for t in host]
    audio_host = torch.empty(BATCH, 1, SAMPLES, pin_memory=True)
    warmup = max(args.warmup, 3)

    resident = lambda: model(*dev)
    end_to_end = lambda: model.forward_host(*host, out=audio_host)

    for _ in range(warmup):
        resident()
    launches_before = _lib.launch_count()
    with Clocks(ctx.local_rank) as clocks:
        ms = ctx.timed(resident, args.steps)
    launches = _lib.launch_count() - launches_before

    # End to end: host buffers in, host audio out, every step's copies inside the timed
    # region.  stream_host pipelines them (the D2H of step i and the H2D of step i + 1 overlap
    # the synthesis of step i + 1); forward_host is the same call made synchronously
    def streamed():
        for audio in model.stream_host(host for _ in range(args.steps)):
            last = audio
        return last

    for _ in range(2):
        end_to_end()
    for _ in model.stream_host(host for _ in range(3)):
        pass
    e2e_seconds = ctx.wall(streamed, 1)
    e2e_synchronous_seconds = ctx.wall(end_to_end, args.steps)

    # Roofline of the dominant kernel: same steps again with per-launch CUDA events
    dominant = 'conv1d_kernel' if math == _lib.MATH_FP32_SIMT else 'conv1d_tc_kernel'
    names = ('conv1d_kernel', 'conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel', 'conv_transpose1d_kernel',
             'conv_transpose1d_tc_kernel', 'planes_from_f32_kernel', 'zero_plane_pads_kernel',
             'head_kernel', 'features_kernel', 'speaker_bias_kernel')
    shares = ctx.kernels(resident, names, args.steps)
    # the 72 residual-block convolutions run on three kernels (conv1d_tc_kernel: time on the M side;
    # conv1d_tcw_kernel: weights on the M side, the narrow layers; conv_pair_tc_kernel: a fused pair)
    resblock_kernels = ('conv1d_tc_kernel', 'conv1d_tcw_kernel', 'conv_pair_tc_kernel')
    tensor_ms = sum(shares.get(n, {'ms_per_step': 0.})['ms_per_step'] for n in resblock_kernels)
    tensor_launches = sum(shares.get(n, {'launches_per_step': 0})['launches_per_step'] for n in resblock_kernels)
    if math == _lib.MATH_FP32_SIMT:
        tensor_ms = shares[dominant]['ms_per_step']
        tensor_launches = shares[dominant]['launches_per_step']

    total_samples = ctx.world * BATCH * SAMPLES * args.steps
    flop_per_sample = (
        FLOP_PER_SAMPLE_CONV1D if math == _lib.MATH_FP32_SIMT else FLOP_PER_SAMPLE_RESBLOCKS)
    flops = BATCH * SAMPLES * flop_per_sample          # per step, residual-block convolutions
    achieved = flops / (tensor_ms * 1e-3) / 1e12 if tensor_ms else None
    peak = ctx.peak
    traffic, traffic_file = measured_traffic(dominant)
    result = {
        'metric': METRIC,
        'value': total_samples / (ms * 1e-3),
        'unit': UNIT,
        'n_gpus': ctx.world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if math == _lib.MATH_FP32_SIMT else (
            'f32 (bf16x3 tensor-core products; fp16 + 2 x e4m3 correction products in the C = 128 '
            'residual blocks; fp32 accumulate)' if model.f8 else
            'f32 (bf16x3 tensor-core products, fp32 accumulate)'),
        'data': 'synthetic',
        'config': workload_config(ctx.world),
        'clocks': clocks.summary(),
        'gpu_launches': launches,
        'e2e': {
            'value': total_samples / e2e_seconds, 'unit': UNIT,
            'h2d_bytes_per_step': sum(t.numel() * t.element_size() for t in host),
            'd2h_bytes_per_step': audio_host.numel() * 4,
            'api': 'promonet_b200.model.Generator.stream_host (pinned host batches in, pinned host audio '
                   'out; copies on their own streams overlap the next step\'s kernels)',
            'synchronous_value': total_samples / e2e_synchronous_seconds,
            'synchronous_api': 'promonet_b200.model.Generator.forward_host -> pmn_generator_forward_host'},
        'roofline': {
            'bound': 'tensor',
            'kernel': dominant + (' + conv1d_tcw_kernel + conv_pair_tc_kernel (the residual-block convolutions)'
                                  if 'conv_pair_tc_kernel' in shares or 'conv1d_tcw_kernel' in shares else ''),
            'achieved': achieved, 'peak': peak['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
            'frac': achieved / peak['bf16_tflops_sustained'] if achieved else None,
            'peak_source': f"{peak['source']} sustained bf16 cuBLAS (kernel timed inside a long step)",
            'traffic': traffic, 'traffic_source': traffic_file,
            'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)',
            'launches_per_step': tensor_launches,
            'avg_launch_ms': tensor_ms / tensor_launches if tensor_launches else None,
            'flop_per_launch': flops / tensor_launches if tensor_launches else None,
            'how': 'CUDA events around every launch (pmn_profile_*), second pass of the same steps; '
                   'algorithmic FLOP = 2 x MAC of the 72 residual-block convolutions (SURVEY Appendix A)',
            'whole_step_tflops': BATCH * SAMPLES * FLOP_PER_SAMPLE_TOTAL * args.steps / (ms * 1e-3) / 1e12,
            'hbm_frac_layer_boundary': (BATCH * SAMPLES * args.steps / (ms * 1e-3)) * BYTES_PER_SAMPLE / (peak['hbm_gbs'] * 1e9),
            'kernels': shares},
    }
    if ctx.cpu:
        # the CPU leg's output on the first utterances of this rank's batch is the checker
        baseline, _, expected, _ = cpu_synthesis(2, 1, seed=seed)
        actual = model(*dev)[:CPU_SAMPLE_BATCH].cpu()
        errors = [relative_error(actual[i], expected[i]) for i in range(CPU_SAMPLE_BATCH)]
        result['cpu_baseline'] = baseline
        result['parity'] = {
            'max_rel_err': max(errors), 'utterances_checked': CPU_SAMPLE_BATCH, 'bar': PARITY_BAR,
            'shape': [BATCH, FRAMES], 'against': baseline['kind'],
            'measure': 'max|a - b| / max|b| per utterance, this run\'s B=32 output vs the CPU leg'}
        if max(errors) > PARITY_BAR:
            raise SystemExit(f'parity failure: {errors} exceeds {PARITY_BAR}')
    else:
        result['parity'] = {'max_rel_err': None, 'utterances_checked': 0,
                            'note': 'the CPU checker runs on rank 0 at N = 1 only'}
    del model
    return result


def plumbing(ctx):
    """configs[0]: promonet.synthesize.from_features, 1 utterance x 1 s (86 frames)"""
    import promonet_b200
    from promonet_b200 import synthetic
    torch = ctx.torch
    loudness, pitch, periodicity, ppg, _, _, _ = synthetic.synthesis(1, 86, seed=1234)
    call = lambda: promonet_b200.synthesize.from_features(
        loudness[0], pitch, periodicity, ppg, speaker=0, gpu=ctx.local_rank)
    audio = call()
    seconds = ctx.wall(call, 5) / 5
    result = {
        'metric': METRIC, 'unit': UNIT, 'value': ctx.world * 86 * HOPSIZE / seconds,
        'ms_per_call': seconds * 1e3, 'shape': list(audio.shape),
        'config': {'workload': 'promonet.synthesize.from_features, 1 utterance x 1 s (86 frames), '
                               'host tensors in, host-visible audio out, per GPU'}}
    if ctx.cpu:
        forward, kind, _ = reference_generator()
        batch = [
            loudness, pitch, periodicity, ppg,
            torch.zeros(1, dtype=torch.long), torch.ones(1), torch.ones(1)]
        with torch.inference_mode():
            forward(batch)
            begin = time.perf_counter()
            expected = forward(batch)
            cpu_seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': 86 * HOPSIZE / cpu_seconds, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': kind,
            'sample': 'the same utterance (speaker 0, ratios 1) through Generator.forward'}
        result['parity'] = {
            'max_rel_err': relative_error(audio.cpu().reshape(-1), expected.reshape(-1)),
            'utterances_checked': 1, 'bar': PARITY_BAR}
    return result


def preprocess(ctx):
    """configs[2]: the from_audio chain on 32 utterances x 10 s per GPU"""
    import promonet_b200
    from promonet_b200 import synthetic
    torch, args = ctx.torch, ctx.args
    steps = max(1, min(args.steps, 10))
    host = synthetic.audio(PRE_BATCH, PRE_SAMPLES, seed=1234 + ctx.rank).pin_memory()
    audio = host.to(ctx.device)
    features = ['loudness', 'pitch', 'periodicity', 'mels']
    step = lambda: promonet_b200.preprocess.from_audio_batch(audio, features=features)

    def end_to_end():
        outputs = promonet_b200.preprocess.from_audio_batch(
            host.to(ctx.device, non_blocking=True), features=features)
        return [o.cpu() for o in outputs]

    for _ in range(3):
        outputs = step()
    from promonet_b200 import _lib
    before = _lib.launch_count()
    ms = ctx.timed(step, steps) / steps
    launches = (_lib.launch_count() - before) / steps
    end_to_end()
    e2e_seconds = ctx.wall(end_to_end, steps) / steps
    names = (
        'stft_kernel', 'loudness_finish_kernel', 'mel_kernel', 'resample_kernel', 'frames_kernel',
        'conv1d_kernel', 'conv1d_tc_kernel', 'im2col_planes_kernel', 'pool_norm_kernel',
        'pool_norm_planes_kernel', 'zero_plane_pads_kernel', 'posterior_kernel', 'band_fill_kernel',
        'viterbi_kernel', 'viterbi_cluster_kernel', 'viterbi_toeplitz_kernel', 'pitch_kernel', 'padded_kernel',
        'column_sums_kernel', 'frame_stats_kernel', 'shared_norm_planes_kernel', 'frame_major_stats_kernel',
    'frame_major_norm_planes_kernel')
    kernels = ctx.kernels(step, names)
    frames = PRE_BATCH * PRE_FRAMES
    cnn_ms = sum(kernels.get(n, {'ms_per_step': 0.})['ms_per_step'] for n in ('conv1d_kernel', 'conv1d_tc_kernel'))
    peak = ctx.peak
    achieved = frames * PRE_FLOP_PER_FRAME / (cnn_ms * 1e-3) / 1e12 if cnn_ms else None
    result = {
        'metric': 'preprocess frames/sec', 'unit': 'frames/s',
        'value': ctx.world * frames / (ms * 1e-3), 'ms_per_step': ms, 'steps': steps,
        'gpu_launches_per_step': launches,
        'config': {
            'workload': f'promonet.preprocess.from_audio chain (A-weighted loudness, log-mel, FCNF0++ '
                        f'pitch + periodicity, Viterbi) on {PRE_BATCH} utterances x 10 s '
                        f'({PRE_SAMPLES} samples -> {PRE_FRAMES} frames) per GPU; FCNF0++ random init '
                        'seed 1234 (penn weights are not available offline)',
            'batch_per_gpu': PRE_BATCH, 'frames_per_utterance': PRE_FRAMES,
            'global_batch': PRE_BATCH * ctx.world,
            'parallelism': f'utterance-sharded x{ctx.world}, no collectives'},
        'e2e': {'value': ctx.world * frames / e2e_seconds, 'unit': 'frames/s',
                'h2d_bytes_per_step': host.numel() * 4,
                'd2h_bytes_per_step': sum(o.numel() * o.element_size() for o in outputs),
                'api': 'promonet_b200.preprocess.from_audio_batch (pinned host audio in, host features out)'},
        'roofline': {
            'bound': 'tensor', 'kernel': 'conv1d_tc_kernel (FCNF0++ convolutions)',
            'achieved': achieved, 'peak': peak['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
            'frac': achieved / peak['bf16_tflops_sustained'] if achieved else None,
            'flop_per_frame': PRE_FLOP_PER_FRAME,
            'chain_hbm_frac': frames / (ms * 1e-3) * PRE_BYTES_PER_FRAME / (peak['hbm_gbs'] * 1e9),
            'chain_bytes_per_frame': PRE_BYTES_PER_FRAME, 'kernels': kernels},
    }
    if ctx.cpu:
        from oracle import dsp
        from oracle import penn as oracle_penn
        from oracle import viterbi as oracle_viterbi
        torch.set_num_threads(os.cpu_count())
        state = oracle_penn.init_state(1234)
        one = host[:1].clone()
        begin = time.perf_counter()
        loudness = dsp.loudness(one, 8)
        mels = dsp.linear_to_mel(dsp.magnitude(one))
        pitch, periodicity, aux = oracle_penn.from_audio(state, one)
        seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': PRE_FRAMES / seconds, 'unit': 'frames/s', 'cores': os.cpu_count(),
            'kind': 'port',
            'sample': 'utterance 0 of the batch through oracle/dsp.py + oracle/penn.py + oracle/viterbi.c '
                      '(restatements: librosa / penn / torbi are absent from the reference tree)'}
        # parity of this run's own batch-32 output, utterance 0
        model = promonet_b200.preprocess.core._pitch_model(ctx.device)
        _, _, logits, bins = model(audio[:1], return_intermediates=True)
        with torch.no_grad():
            posterior = torch.softmax(logits[0].cpu(), dim=-1).numpy()[None]
        expected_bins = oracle_viterbi.decode(
            posterior, None, oracle_penn.transition_matrix(256 / 22050).numpy(),
            oracle_penn.initial_distribution().numpy())
        result['parity'] = {
            'loudness': relative_error(outputs[0][0].cpu(), loudness),
            'mels': relative_error(outputs[3][0].cpu(), mels),
            'periodicity_max_abs': float((outputs[2][0].cpu() - periodicity[0]).abs().max()),
            'viterbi_path_agreement_on_device_posterior': float(
                (bins[0].cpu().numpy() == expected_bins[0]).mean()),
            'pitch_path_agreement_vs_cpu_chain': float(
                (bins[0].cpu() == aux['bins']).float().mean()),
            'utterances_checked': 1, 'bar': PARITY_BAR,
            'pinning': 'parity unpinned for the librosa / penn / torbi arithmetic (restated oracle)'}
    return result


def fargan(ctx):
    """configs[4]: FARGAN generator, 32 x 5 s per GPU and 32 in total over 8 GPUs (4 per GPU)"""
    import promonet_b200
    from promonet_b200 import synthetic
    torch, args = ctx.torch, ctx.args
    steps = max(1, min(args.steps, 10))
    state = promonet_b200.model.init.fargan_state(1234)
    model = promonet_b200.model.FarganGenerator(device=ctx.device, state=state)
    host = synthetic.synthesis(BATCH, FRAMES, seed=1234 + ctx.rank)
    dev = [t.to(ctx.device) for t in host]
    small = [t[:4].contiguous() for t in dev]
    step = lambda: model(*dev)
    for _ in range(3):
        audio = step()
    ms = ctx.timed(step, steps) / steps
    model(*small)
    ms_small = ctx.timed(lambda: model(*small), steps) / steps
    kernels = ctx.kernels(step, ('fargan_kernel', 'conv1d_kernel', 'features_kernel', 'cond_input_kernel'))
    samples = BATCH * SAMPLES
    result = {
        'metric': METRIC + ', FARGAN', 'unit': UNIT,
        'value': ctx.world * samples / (ms * 1e-3), 'ms_per_step': ms, 'steps': steps,
        'small_batch': {'batch_per_gpu': 4, 'ms_per_step': ms_small,
                        'value': ctx.world * 4 * SAMPLES / (ms_small * 1e-3)},
        'config': {
            'workload': f'FARGAN Generator.forward, config/fargan.py, random init seed 1234, '
                        f'{BATCH} utterances x 5 s per GPU (small_batch: 4 per GPU = 32 over 8 GPUs)',
            'batch_per_gpu': BATCH, 'frames': FRAMES,
            'parallelism': f'utterance-sharded x{ctx.world}, no collectives'},
        'roofline': {
            'bound': 'latency (1 720 sequential subframes); reported against the tensor peak for scale',
            'achieved': samples * FARGAN_FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12, 'unit': 'TFLOP/s',
            'us_per_subframe': kernels.get('fargan_kernel', {'ms_per_step': 0.})['ms_per_step'] * 1e3 / (FRAMES * 4),
            'kernels': kernels},
    }
    if ctx.cpu:
        from oracle import fargan as oracle_fargan
        torch.set_num_threads(os.cpu_count())
        sample = [t[:2] for t in host]
        with torch.no_grad():
            begin = time.perf_counter()
            expected = oracle_fargan.generator(state, *sample)
            seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': 2 * SAMPLES / seconds, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
            'sample': 'the first 2 utterances through oracle/fargan.py (pinned to the reference by '
                      'tests/golden/fargan.npz)'}
        errors = [relative_error(audio[i].cpu(), expected[i]) for i in range(2)]
        result['parity'] = {'max_rel_err': max(errors), 'utterances_checked': 2, 'bar': PARITY_BAR}
    del model
    return result


def train(ctx):
    """configs[3]: one training step, 8 items x 16 384 samples per GPU"""
    import promonet_b200
    from promonet_b200 import _lib, synthetic
    from promonet_b200.model import init
    from promonet_b200.train.core import Trainer
    torch, args = ctx.torch, ctx.args
    steps = max(1, min(args.steps, 10))
    trainer = Trainer(init.hifigan_state(1234), init.discriminator_state(1234), ctx.device)
    trainer.broadcast_parameters()
    *inputs, audio = synthetic.training(TRAIN_BATCH, TRAIN_FRAMES, 1234 + ctx.rank)
    audio = audio.to(ctx.device)
    spectrograms = promonet_b200.preprocess.spectrogram.from_audio(audio)
    batch = [t.to(ctx.device).contiguous() for t in inputs] + [spectrograms.contiguous(), audio.contiguous()]
    step = lambda: trainer.step_graphed(*batch)
    for _ in range(3):
        losses = step()
    first = [float(v) for v in losses.cpu()]
    ms = ctx.timed(step, steps) / steps
    names = (
        'conv_fprop_tc_kernel', 'conv_dgrad_tc_kernel', 'conv_wgrad_tc_kernel', 'fold_weights_kernel',
        'pack_weights_kernel', 'weight_norm_backward_kernel', 'weight_norm_backward_table_kernel', 'stft_train_kernel',
        'stft_train_backward_kernel', 'mel_loss_kernel', 'l1_mean_kernel', 'mse_to_target_kernel',
        'adamw_kernel', 'adamw_peer_kernel', 'reflect_pad_kernel', 'copy_columns_kernel')
    kernels = ctx.kernels(lambda: trainer.step(*batch), names)
    items = ctx.world * TRAIN_BATCH
    tflops = TRAIN_BATCH * TRAIN_FLOP_PER_ITEM / (ms * 1e-3) / 1e12
    peak = ctx.peak
    exchange = None
    if ctx.world > 1:
        exchange = (
            'pmn_adamw_peer: reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory'
            if trainer.generator.params.peers is not None else 'NCCL all-reduce + AdamW kernel')
    result = {
        'metric': 'training items/sec (16 384-sample chunks)', 'unit': 'items/s',
        'value': items / (ms * 1e-3), 'ms_per_step': ms, 'steps': steps,
        'config': {
            'workload': 'one step of promonet/train/core.py:183-369 (HiFi-GAN generator, 5 multi-period + '
                        'complex multi-band discriminators, LSGAN + feature matching + mel loss, 2 x AdamW), '
                        f'{TRAIN_BATCH} items x 16 384 samples per GPU, three CUDA graphs per step',
            'batch_per_gpu': TRAIN_BATCH, 'global_batch': items, 'frames': TRAIN_FRAMES,
            'parallelism': f'data parallel x{ctx.world}', 'exchange': exchange},
        'dtype': 'f32 storage, tf32 tensor-core products, fp32 accumulate',
        'losses_after_warmup': dict(zip(
            ('discriminator', 'mel', 'feature_matching', 'adversarial', 'generator'), first)),
        'roofline': {
            'bound': 'tensor', 'kernel': 'conv_gemm_tc_kernel / conv_wgrad_tc_kernel',
            'achieved': tflops, 'peak': peak['bf16_tflops_sustained'] / 2, 'unit': 'TFLOP/s',
            'frac': tflops / (peak['bf16_tflops_sustained'] / 2),
            'peak_source': 'half the sustained bf16 rate (kind::tf32)',
            'flop_per_item': TRAIN_FLOP_PER_ITEM, 'kernels': kernels},
    }
    if ctx.cpu:
        from oracle import train as oracle_train
        torch.set_num_threads(os.cpu_count())
        cpu_batch = oracle_train.batch(1, TRAIN_FRAMES, 1234)
        g = oracle_train.leaf_state(init.hifigan_state(1234))
        d = oracle_train.leaf_state(init.discriminator_state(1234))
        optimizers = oracle_train.make_optimizers(g, d)
        begin = time.perf_counter()
        oracle_train.step(g, d, cpu_batch, optimizers)
        seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': 1 / seconds, 'unit': 'items/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': '1 item through oracle/train.py (torch autograd fp32 + AdamW; pinned to the '
                      'reference step by tests/golden/train.npz), one step'}
    ctx.trainers.append(trainer)
    return result


def gpu_eager_baseline(ctx):
    """The reference's modules run eagerly by PyTorch on this B200: strict
    fp32 (the parity yardstick), TF32 convolutions (PyTorch's default) and bf16 autocast"""
    torch = ctx.torch
    from oracle import reference
    from promonet_b200 import synthetic
    promonet = reference.load()
    if promonet is None:
        return {'unavailable': 'the reference tree (oracle/_ref) did not travel to this box'}
    model = reference.generator(promonet, 1234).to(ctx.device)
    batch = [t.to(ctx.device) for t in synthetic.synthesis(BATCH, FRAMES, seed=1234)]
    torch.backends.cudnn.benchmark = True
    result = {'what': f'promonet.model.Generator.forward on cuda, {BATCH} x {FRAMES} frames, '
                      f'torch {torch.__version__} eager, cuDNN {torch.backends.cudnn.version()}',
              'unit': UNIT, 'modes': {}}
    exact = None
    for name, tf32, autocast in (('fp32', False, False), ('tf32', True, False), ('bf16_autocast', True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def forward():
            with torch.inference_mode(), torch.autocast('cuda', torch.bfloat16, enabled=autocast):
                return reference.forward(model, *batch)
        try:
            for _ in range(2):
                forward()
            torch.cuda.synchronize()
            start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
            start.record()
            for _ in range(3):
                audio = forward()
            stop.record()
            torch.cuda.synchronize()
            ms = start.elapsed_time(stop) / 3
        except Exception as error:      # a mode that does not run is reported, not fatal
            result['modes'][name] = {'error': repr(error)[:200]}
            continue
        audio = audio.float()
        if exact is None:
            exact = audio
        result['modes'][name] = {
            'ms_per_step': ms, 'value': BATCH * SAMPLES / (ms * 1e-3),
            'error_vs_fp32': relative_error(audio, exact)}
    del model, batch
    torch.cuda.empty_cache()
    # the training step (configs[3]) the same way: the reference's loop body, its modules,
    # losses and optimizers; fp16 autocast + GradScaler is how the reference trains
    from promonet_b200 import synthetic
    *inputs, audio = synthetic.training(TRAIN_BATCH, TRAIN_FRAMES, 1234)
    audio = audio.to(ctx.device)
    with torch.no_grad():
        spectrograms = promonet.preprocess.spectrogram.from_audio(audio)
    batch = [t.to(ctx.device) for t in inputs] + [spectrograms, audio]
    result['train'] = {
        'what': f'promonet/train/core.py:183-369 composed from the reference modules on cuda, '
                f'{TRAIN_BATCH} items x 16 384 samples', 'unit': 'items/s', 'modes': {}}
    for name, tf32, autocast in (('fp32', False, False), ('tf32', True, False), ('fp16_autocast', True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            step = reference.TrainingStep(promonet, ctx.device, autocast)
            for _ in range(2):
                step(*batch)
            torch.cuda.synchronize()
            start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
            start.record()
            for _ in range(3):
                loss = step(*batch)
            stop.record()
            torch.cuda.synchronize()
            ms = start.elapsed_time(stop) / 3
            result['train']['modes'][name] = {
                'ms_per_step': ms, 'value': TRAIN_BATCH / (ms * 1e-3), 'generator_loss': float(loss)}
            del step
        except Exception as error:
            result['train']['modes'][name] = {'error': repr(error)[:200]}
        torch.cuda.empty_cache()
    # FARGAN (configs[4]) the same way, in a process of its own: the reference freezes its
    # configuration at import.  A Python loop of 4 subframes x frames, ~60 launches each.
    try:
        import subprocess
        done = subprocess.run(
            [sys.executable, '-m', 'oracle.reference', '--fargan-eager', str(BATCH), str(FRAMES)],
            cwd=str(ROOT), capture_output=True, text=True, timeout=300)
        lines = [line for line in done.stdout.splitlines() if line.startswith('{')]
        result['fargan'] = json.loads(lines[-1]) if lines else {'error': done.stderr[-300:]}
    except Exception as error:
        result['fargan'] = {'error': repr(error)[:200]}
    return result


def run_b200(args):
    ctx = Context(args)
    ctx.trainers = []
    result = synthesis(ctx)
    wanted = None if args.only is None else set(args.only.split(','))
    secondary = {}
    if not args.no_secondary:
        for name, workload in (
                ('preprocess', preprocess), ('fargan', fargan), ('plumbing', plumbing), ('train', train)):
            if wanted is not None and name not in wanted:
                continue
            try:
                secondary[name] = workload(ctx)
            except SystemExit:
                raise
            except Exception as error:      # a secondary line must not take the headline down
                import traceback
                secondary[name] = {'error': repr(error)[:300], 'trace': traceback.format_exc()[-600:]}
            ctx.torch.cuda.empty_cache()
        if ctx.cpu and (wanted is None or 'eager' in wanted):
            try:
                result['gpu_eager_baseline'] = gpu_eager_baseline(ctx)
            except Exception as error:
                result['gpu_eager_baseline'] = {'error': repr(error)[:300]}
    result['secondary'] = secondary
    if ctx.rank == 0:
        print(json.dumps(result), flush=True)
    if ctx.world > 1:
        # closes the trainers' peer-memory mappings first (a teardown with live mappings has been
        # seen to wait on the peers; the helper ends the process if it still does)
        from promonet_b200 import parallel
        parallel.shutdown(*ctx.trainers)


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
