"""config 5 of BASELINE.json: FARGAN generator, batch-32 x 5 s synthesis.

    python profiles/bench_fargan.py [--batch 32] [--frames 430] [--steps 3]
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
from oracle import fargan as oracle_fargan  # noqa: E402
from oracle import inputs  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--batch', type=int, default=32)
    parser.add_argument('--frames', type=int, default=430)
    parser.add_argument('--steps', type=int, default=3)
    parser.add_argument('--no-cpu', action='store_true')
    args = parser.parse_args()
    state = promonet_b200.model.init.fargan_state(1234)
    model = promonet_b200.model.FarganGenerator(state=state)
    host = inputs.synthesis(args.batch, args.frames, seed=1234)
    dev = [t.cuda() for t in host]
    for _ in range(2):
        model(*dev)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    for _ in range(args.steps):
        model(*dev)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / args.steps
    samples = args.batch * args.frames * 256
    _lib.profile(True)
    model(*dev)
    torch.cuda.synchronize()
    kernels = {}
    for name in ('fargan_kernel', 'conv1d_kernel', 'features_kernel', 'cond_input_kernel'):
        total, count = _lib.profile_read(name)
        if count:
            kernels[name] = {'ms': round(total, 3), 'launches': count}
    _lib.profile(False)
    subframes = args.frames * 4
    result = {
        'metric': 'audio samples/sec synthesized (22.05 kHz), FARGAN', 'value': samples / (ms * 1e-3),
        'unit': 'samples/s', 'ms_per_step': ms, 'batch': args.batch, 'frames': args.frames,
        'us_per_subframe': kernels['fargan_kernel']['ms'] * 1e3 / subframes,
        'gflops': samples * 73.8e3 / (ms * 1e-3) / 1e9, 'kernels': kernels}
    if not args.no_cpu:
        torch.set_num_threads(os.cpu_count())
        small = [t[:2] for t in host]
        with torch.no_grad():
            begin = time.perf_counter()
            oracle_fargan.generator(state, *small)
            seconds = time.perf_counter() - begin
        result['cpu_baseline'] = {
            'value': 2 * args.frames * 256 / seconds, 'unit': 'samples/s', 'cores': os.cpu_count(),
            'kind': 'port', 'sample': '2 utterances through oracle/fargan.py'}
    print(json.dumps(result))


if __name__ == '__main__':
    main()
