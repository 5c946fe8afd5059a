"""Where the tensor-core conv spends its time: CUDA-event time per launch and the
in-kernel cycle counters of each warp role (producer / MMA issuer / epilogue).

    python profiles/tc_breakdown.py [channels] [f8]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200 import _lib  # noqa: E402
from promonet_b200.tc_probe import run_tc_conv  # noqa: E402


def main():
    lib = _lib.library()
    counters = torch.zeros(148, 10, 4, dtype=torch.int64, device='cuda')
    shapes = ((256, 3440), (128, 27520), (64, 55040), (32, 110080))
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if s[0] == int(sys.argv[1])]
    f8 = len(sys.argv) > 2 and sys.argv[2] == 'f8'
    for channels, t_len in shapes:
        for kernel in (3, 7, 11) if f8 else (3, 11):
            for mode in ('c1', 'c2', 'c2acc'):
                counters.zero_()
                lib.pmn_debug_tc_counters(counters.data_ptr())
                ms = run_tc_conv(32, channels, t_len, kernel, mode, repeats=3, f8=f8)
                lib.pmn_debug_tc_counters(None)
                c = counters.double().cpu()
                prod, mma, epi = c[:, 0], c[:, 1], c[:, 2:].mean(1)
                flops = 2 * 32 * channels * channels * kernel * t_len
                print(
                    f'C={channels:3d} k={kernel:2d} {mode:5s} {ms:7.3f} ms '
                    f'{flops / ms / 1e9:7.1f} TFLOP/s | kcycles/CTA: '
                    f'producer {prod[:, 0].mean() / 1e3:6.0f} (x_empty {prod[:, 1].mean() / 1e3:5.0f} '
                    f'w_empty {prod[:, 2].mean() / 1e3:5.0f}) | '
                    f'mma {mma[:, 0].mean() / 1e3:6.0f} (x_full {mma[:, 1].mean() / 1e3:5.0f} '
                    f'w_full {mma[:, 2].mean() / 1e3:5.0f} acc_empty {mma[:, 3].mean() / 1e3:5.0f}) | '
                    f'epilogue {epi[:, 0].mean() / 1e3:6.0f} (acc_full {epi[:, 1].mean() / 1e3:5.0f})')


if __name__ == '__main__':
    main()
