"""Which residual blocks should take the fused pair kernel?  Times the benchmark's synthesis
step (32 x 430 frames) with each block of stages 1-3 fused on its own, all fused and none
fused, and prints the mask of the blocks that are faster fused.

    python profiles/pair_selection.py [--steps 5]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import promonet_b200  # noqa: E402
from promonet_b200 import _lib, synthetic  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--batch', type=int, default=32)
    parser.add_argument('--frames', type=int, default=430)
    args = parser.parse_args()
    state = promonet_b200.model.init.hifigan_state(1234)
    inputs = [t.cuda() for t in synthetic.synthesis(args.batch, args.frames)]

    def measure(mask):
        model = promonet_b200.model.Generator(state=state, pair_mask=mask)
        for _ in range(3):
            audio = model(*inputs)
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
        start.record()
        for _ in range(args.steps):
            model(*inputs)
        stop.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(stop) / args.steps
        _lib.profile(True)
        model(*inputs)
        torch.cuda.synchronize()
        kernels = {}
        for name in ('conv1d_tc_kernel', 'conv_pair_tc_kernel', 'planes_from_f32_kernel', 'zero_plane_pads_kernel'):
            total, count = _lib.profile_read(name)
            kernels[name] = (round(total, 3), count)
        _lib.profile(False)
        return ms, kernels, audio

    base_ms, base_kernels, base_audio = measure(0)
    print(json.dumps({'mask': '0x000', 'ms': base_ms, 'kernels': base_kernels}))
    best = 0
    for stage in (1, 2, 3):
        for block, kernel in enumerate((3, 7, 11)):
            bit = 1 << (3 * stage + block)
            ms, kernels, audio = measure(bit)
            same = bool(torch.equal(audio, base_audio))
            print(json.dumps({
                'mask': hex(bit), 'channels': 256 >> stage, 'kernel': kernel, 'ms': ms,
                'gain_ms': base_ms - ms, 'bit_identical': same, 'kernels': kernels}))
            if ms < base_ms:
                best |= bit
    for mask in sorted({0xFF8, best}):
        ms, kernels, audio = measure(mask)
        print(json.dumps({'mask': hex(mask), 'ms': ms, 'gain_ms': base_ms - ms,
                          'bit_identical': bool(torch.equal(audio, base_audio)), 'kernels': kernels}))
    print(json.dumps({'best_mask': hex(best)}))


if __name__ == '__main__':
    main()
