"""In-training validation throughput: promonet_b200.train.evaluate.

    python profiles/bench_evaluate.py [--items 8] [--seconds 5] [--steps 3]

One step = evaluate() over `items` synthetic validation items of `seconds` each: per item
seven synthesized versions (reconstruction, 2 pitch shifts, 2 time stretches, 2 loudness
scalings: promonet/train/core.py:568-799), each re-analysed (loudness, pitch, periodicity)
and scored.  Prints one JSON line: items/s, launches per item, the device time of the
validation-only kernels, and the CPU oracle (oracle generator -> oracle features -> oracle
metrics, one version of one item per condition) on the host cores.
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import promonet_b200  # noqa: E402
from promonet_b200 import _lib  # noqa: E402
from promonet_b200.model import init  # noqa: E402
from promonet_b200.train import evaluate  # noqa: E402
from oracle import dsp, hifigan, inputs  # noqa: E402
from oracle import metrics as oracle_metrics  # noqa: E402
from oracle import penn as oracle_penn  # noqa: E402

KERNELS = ('metrics_update_kernel', 'edit_contour_kernel', 'grid_sample_kernel')


def loader(items, frames):
    result = []
    for index in range(items):
        loudness, pitch, periodicity, ppg, speakers, _, _ = inputs.synthesis(1, frames, seed=index)
        result.append((
            None, loudness, pitch, periodicity, ppg, speakers, None, None, torch.zeros(1),
            inputs.audio(1, frames * 256, seed=index)[:, None], None))
    return result


def cpu_item(state, pitch_state, batch, every_version=False):
    """The reference's loop for one item (batch 1 per version) on the CPU oracle; by default
    only the reconstruction (the seven versions together are 7.1 x its frames)"""
    _, loudness, pitch, periodicity, ppg, speakers, _, _, _, _, _ = batch
    versions = [(loudness, pitch, periodicity, ppg)]
    for ratio in promonet_b200.EVALUATION_RATIOS if every_version else ():
        versions.append((loudness, ratio * pitch, periodicity, ppg))
        versions.append((loudness + 10 * math.log2(ratio), pitch, periodicity, ppg))
        stretched = oracle_metrics.edit_from_features(
            loudness[0], pitch, periodicity, ppg[0], time_stretch_ratio=ratio)
        versions.append((stretched[0][None], stretched[1], stretched[2], stretched[3][None]))
    for features in versions:
        metrics = oracle_metrics.Metrics()
        with torch.no_grad():
            audio = hifigan.generator(state, *features, speakers, torch.ones(1), torch.ones(1))[0]
        predicted_pitch, predicted_periodicity, _ = oracle_penn.from_audio(pitch_state, audio)
        metrics.update(
            *features[:3], None, dsp.loudness(audio, 8), predicted_pitch, predicted_periodicity, None)
        metrics()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--items', type=int, default=8)
    parser.add_argument('--seconds', type=float, default=5.)
    parser.add_argument('--steps', type=int, default=3)
    parser.add_argument('--no-cpu', action='store_true')
    args = parser.parse_args()
    frames = int(args.seconds * 22050) // 256
    state = init.hifigan_state(1234)
    generator = promonet_b200.model.Generator(state=state)
    batches = loader(args.items, frames)
    step = lambda: evaluate(None, 1, generator, batches)
    step()
    torch.cuda.synchronize()
    launches = _lib.launch_count()
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / args.steps
    launches = (_lib.launch_count() - launches) / (args.steps * args.items)
    _lib.profile(True)
    step()
    torch.cuda.synchronize()
    kernels = {}
    for name in KERNELS:
        total, count = _lib.profile_read(name)
        if count:
            kernels[name] = {'ms': round(total, 4), 'launches': count}
    _lib.profile(False)
    result = {
        'metric': 'validation items/sec (7 synthesized + re-analysed versions per item)',
        'value': args.items / (ms * 1e-3), 'unit': 'items/s', 'ms_per_item': ms / args.items,
        'items': args.items, 'frames_per_item': frames, 'gpu_launches_per_item': launches,
        'validation_kernels': kernels}
    if not args.no_cpu:
        torch.set_num_threads(os.cpu_count())
        pitch_state = oracle_penn.init_state(1234)
        begin = time.perf_counter()
        cpu_item(state, pitch_state, batches[0])
        seconds = time.perf_counter() - begin
        # frames of the seven versions relative to the reconstruction: 5 + 1 / .717 + 1 / 1.414
        work = 5. + sum(1. / ratio for ratio in promonet_b200.EVALUATION_RATIOS)
        result['cpu_baseline'] = {
            'value': 1. / (seconds * work), 'unit': 'items/s', 'cores': os.cpu_count(),
            'kind': 'port',
            'sample': f'the reconstruction of 1 item at batch 1 ({seconds:.1f} s: oracle/hifigan.py -> '
                      f'oracle/dsp.py + oracle/penn.py -> oracle/metrics.py), scaled by the {work:.2f} x '
                      'frames of the seven versions the reference loop synthesizes per item'}
    print(json.dumps(result))


if __name__ == '__main__':
    main()
