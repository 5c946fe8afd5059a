"""Secondary metric of BASELINE.json: preprocess frames/s (frame = 256-sample hop).

    python profiles/bench_preprocess.py [--batch 32] [--seconds 10] [--steps 3]

One step = promonet_b200.preprocess.from_audio_batch over `batch` synthetic
utterances (loudness + log-mel + pitch + periodicity: STFT features, FCNF0++
network, Viterbi).  Prints one JSON line with the per-kernel device times and the
CPU oracle (oracle/dsp.py + oracle/penn.py) on a bounded sample.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import promonet_b200  # noqa: E402
from promonet_b200 import _lib  # noqa: E402
from oracle import dsp, inputs  # noqa: E402
from oracle import penn as oracle_penn  # noqa: E402

KERNELS = (
    'stft_kernel', 'loudness_finish_kernel', 'resample_kernel', 'frames_kernel', 'conv1d_kernel',
    'conv1d_tc_kernel', 'im2col_planes_kernel', 'pool_norm_kernel', 'pool_norm_planes_kernel',
    'zero_plane_pads_kernel', 'posterior_kernel', 'band_fill_kernel', 'viterbi_kernel',
    'viterbi_cluster_kernel', 'pitch_kernel', 'padded_kernel', 'column_sums_kernel',
    'frame_stats_kernel', 'shared_norm_planes_kernel', 'frame_major_stats_kernel',
    'frame_major_norm_planes_kernel')


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--batch', type=int, default=32)
    parser.add_argument('--seconds', type=float, default=10.)
    parser.add_argument('--steps', type=int, default=3)
    parser.add_argument('--no-cpu', action='store_true')
    args = parser.parse_args()
    samples = int(args.seconds * 22050)
    audio = inputs.audio(args.batch, samples).cuda()
    features = ['loudness', 'pitch', 'periodicity', 'mels']
    step = lambda: promonet_b200.preprocess.from_audio_batch(audio, features=features)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / args.steps
    frames = args.batch * (samples // 256)
    _lib.profile(True)
    step()
    torch.cuda.synchronize()
    kernels = {}
    for name in KERNELS:
        total, count = _lib.profile_read(name)
        if count:
            kernels[name] = {'ms': round(total, 3), 'launches': count}
    _lib.profile(False)
    result = {
        'metric': 'preprocess frames/sec', 'value': frames / (ms * 1e-3), 'unit': 'frames/s',
        'ms_per_step': ms, 'batch': args.batch, 'frames_per_utterance': samples // 256,
        'features': features, 'kernels': kernels,
        'flop_per_frame_cnn': 394.6e6,
        'cnn_tflops': frames * 394.6e6 / (
            (kernels.get('conv1d_kernel', {'ms': 0.})['ms'] +
             kernels.get('conv1d_tc_kernel', {'ms': 0.})['ms']) * 1e-3) / 1e12}
    if not args.no_cpu:
        torch.set_num_threads(os.cpu_count())
        state = oracle_penn.init_state(1234)
        one = audio[:1].cpu()
        begin = time.perf_counter()
        dsp.loudness(one, 8)
        dsp.linear_to_mel(dsp.magnitude(one))
        oracle_penn.from_audio(state, one)
        seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': (samples // 256) / seconds, 'unit': 'frames/s', 'cores': os.cpu_count(),
            'kind': 'port', 'sample': '1 utterance through oracle/dsp.py + oracle/penn.py'}
    print(json.dumps(result))


if __name__ == '__main__':
    main()
