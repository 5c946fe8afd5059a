"""Where the fused pair kernel spends its time, at the benchmark's stage shapes: kernel time
(CUDA events around the launch) for every variant next to the two conv1d_tc launches it
replaces, and the in-kernel cycle counters of each warp role.

    python profiles/pair_breakdown.py [--channels 64] [--variants 0,1,2,3]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200 import _lib  # noqa: E402

SHAPES = {128: 27520, 64: 55040, 32: 110080}


def kernel_ms(name, function, repeats=3):
    function()
    torch.cuda.synchronize()
    _lib.profile(True)
    for _ in range(repeats):
        function()
    torch.cuda.synchronize()
    total, count = _lib.profile_read(name)
    _lib.profile(False)
    return total / repeats


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--channels', default='128,64,32')
    parser.add_argument('--variants', default='0,1,2,3')
    parser.add_argument('--kernels', default='3,7,11')
    parser.add_argument('--dilations', default='1,5')
    args = parser.parse_args()
    lib = _lib.library()
    counters = torch.zeros(148, 10, 4, dtype=torch.int64, device='cuda')
    batch = 32
    for channels in (int(c) for c in args.channels.split(',')):
        t_len = SHAPES[channels]
        torch.manual_seed(0)
        x = torch.randn(batch, channels, t_len, device='cuda')
        out = torch.empty_like(x)
        planes = torch.empty_like(x)
        for k in (int(v) for v in args.kernels.split(',')):
            w1 = torch.randn(channels, channels, k, device='cuda') / (channels * k) ** .5
            w2 = torch.randn(channels, channels, k, device='cuda') / (channels * k) ** .5
            b1, b2 = torch.randn(channels, device='cuda'), torch.randn(channels, device='cuda')
            size = lib.pmn_conv1d_tc_workspace_bytes(batch, channels, t_len, k)
            workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
            for dilation in (int(v) for v in args.dilations.split(',')):
                def two_launches():
                    _lib.check(lib.pmn_conv1d_tc(
                        x.data_ptr(), w1.data_ptr(), b1.data_ptr(), None, None, planes.data_ptr(), None, 0, 1.,
                        batch, channels, t_len, k, dilation, 0.1, 0.1, workspace.data_ptr(), size, _lib.stream()))
                    _lib.check(lib.pmn_conv1d_tc(
                        planes.data_ptr(), w2.data_ptr(), b2.data_ptr(), x.data_ptr(), out.data_ptr(), None, None, 0, 1.,
                        batch, channels, t_len, k, 1, 1., 1., workspace.data_ptr(), size, _lib.stream()))
                base = kernel_ms('conv1d_tc_kernel', two_launches)
                # the second launch of the real pipeline also writes planes (+4 B per element)
                flops = 2 * 2 * batch * channels * channels * k * t_len
                print(f'C={channels:3d} k={k:2d} d={dilation}: two conv1d_tc launches {base:6.3f} ms '
                      f'({flops / base / 1e9:6.1f} TFLOP/s)')
                for variant in (int(v) for v in args.variants.split(',')):
                    counters.zero_()
                    lib.pmn_debug_pair_tc(counters.data_ptr(), variant)

                    def fused():
                        _lib.check(lib.pmn_conv_pair_tc(
                            x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                            out.data_ptr(), None, 0, 1., batch, channels, t_len, k, dilation, 0.1,
                            workspace.data_ptr(), size, _lib.stream()))
                    ms = kernel_ms('conv_pair_tc_kernel', fused)
                    lib.pmn_debug_pair_tc(None, -1)
                    c = counters.double().cpu().mean(0) / 1e3      # kcycles per CTA, mean over CTAs
                    print(
                        f'    variant {variant}: {ms:6.3f} ms ({flops / ms / 1e9:6.1f} TFLOP/s, x{base / ms:4.2f}) | kcycles '
                        f'total {c[1, 0]:6.0f} | mma waits: x_full {c[1, 1]:5.0f} w_full {c[1, 2]:5.0f} '
                        f'acc1_empty {c[1, 3]:5.0f} mid_full {c[5, 0]:5.0f} acc2_empty {c[5, 1]:5.0f} | '
                        f'converter wait x_empty {c[2, 1]:5.0f} of {c[2, 0]:6.0f} | '
                        f'mid waits acc1_full {c[3, 1]:5.0f} mid_empty {c[3, 2]:5.0f} of {c[3, 0]:6.0f} | '
                        f'final wait acc2_full {c[4, 1]:5.0f} of {c[4, 0]:6.0f} | producer wait w_empty {c[0, 1]:5.0f}')


if __name__ == '__main__':
    main()
