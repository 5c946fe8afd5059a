"""Synthesis step (32 x 430 frames, shipped defaults) with each variant of the fused pair kernel
(pmn_debug_pair_tc: converter / final warp counts, mid-image buffers); results do not depend on it.

    python profiles/pair_variants.py [--steps 10]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import promonet_b200  # noqa: E402
from promonet_b200 import _lib, synthetic  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--steps', type=int, default=10)
    args = parser.parse_args()
    lib = _lib.library()
    state = promonet_b200.model.init.hifigan_state(1234)
    model = promonet_b200.model.Generator(state=state)
    inputs = [t.cuda() for t in synthetic.synthesis(32, 430)]
    reference = None
    for round_index in range(2):
        for variant in (-1, 0, 1, 2, 3):
            lib.pmn_debug_pair_tc(None, variant)
            for _ in range(3):
                audio = model(*inputs)
            torch.cuda.synchronize()
            start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
            start.record()
            for _ in range(args.steps):
                audio = model(*inputs)
            stop.record()
            torch.cuda.synchronize()
            if reference is None:
                reference = audio.clone()
            print(f'round {round_index} variant {variant:2d}: {start.elapsed_time(stop) / args.steps:.3f} ms per step, '
                  f'same bits as the default: {bool(torch.equal(audio, reference))}', flush=True)
    lib.pmn_debug_pair_tc(None, -1)


if __name__ == '__main__':
    main()
