"""Per-layer throughput of the training convolution kernels on the shapes of one training
step (8 items x 16 384 samples per GPU): tcgen05 tf32 forward / data gradient
(train_conv_tc.cu), and weight gradient (train_conv_tc.cu).

    python profiles/bench_conv_tc.py [--wgrad] [--filter substring]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200.train import ops  # noqa: E402

B = 8
# name, batch, c_in, c_out, size_in, kernel, stride, dilation, padding
SHAPES = []
for channels, t in ((256, 512), (128, 4096), (64, 8192), (32, 16384)):
    for k in (3, 11):
        for d in (1, 5):
            SHAPES.append((f'G c{channels} k{k} d{d}', B, channels, channels, (t, 1), (k, 1), 1, (d, 1),
                           (d * (k - 1) // 2, 0)))
for p in (2, 11):
    h = -(-16384 // p)
    sizes = [h]
    for _ in range(4):
        sizes.append((sizes[-1] + 4 - 5) // 3 + 1)
    chans = (1, 32, 128, 512, 1024, 1024)
    for i in range(1, 5):
        SHAPES.append((f'MPD p{p} conv{i}', 2 * B, chans[i], chans[i + 1], (sizes[i], p), (5, 1),
                       (3, 1) if i < 4 else 1, 1, (2, 0)))
for width in (51, 129):
    SHAPES.append((f'CMB w{width} conv1', 2 * B, 32, 32, (64, width), (3, 9), (1, 2), 1, (1, 4)))
SHAPES.append(('CMB w33 conv4', 2 * B, 32, 32, (64, 17), (3, 3), 1, 1, (1, 1)))


def timed(fn, repeats=5):
    fn()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    for _ in range(repeats):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / repeats


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--filter', default='')
    parser.add_argument('--wgrad', action='store_true')
    parser.add_argument('--cycles', action='store_true', help='cycle breakdown of CTA 0 (forward)')
    parser.add_argument('--plain', action='store_true', help='no fused activation / residual / mask')
    args = parser.parse_args()
    print(f'{"layer":22s} {"GFLOP":>8s} {"fprop ms":>9s} {"TF/s":>7s} {"dgrad ms":>9s} {"TF/s":>7s}'
          + (f' {"wgrad ms":>9s} {"TF/s":>7s}' if args.wgrad else ''))
    totals = [0., 0., 0., 0.]
    for name, batch, c_in, c_out, size, kernel, stride, dilation, padding in SHAPES:
        if args.filter not in name:
            continue
        geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
        taps = kernel[0] * kernel[1]
        x = torch.randn(batch, c_in, *size, device='cuda')
        w = torch.randn(c_out, c_in, *kernel, device='cuda') / (c_in * taps) ** .5
        y = torch.empty(batch, c_out, geom.h_out, geom.w_out, device='cuda')
        dy = torch.randn_like(y)
        dx = torch.empty_like(x)
        packed = ops.pack_weight_taps(
            w, torch.empty(ops.packed_floats(c_out, c_in, taps), device='cuda'), c_out, c_in, taps, False)
        packed_t = ops.pack_weight_taps(
            w, torch.empty(ops.packed_floats(c_in, c_out, taps), device='cuda'), c_out, c_in, taps, True)
        flop = 2. * batch * geom.h_out * geom.w_out * c_out * c_in * taps
        fused = {} if args.plain else dict(a_act=ops.ACT_LRELU, a_slope=.1, residual=dy)
        forward = timed(lambda: ops.conv_gemm_tc(geom, False, x, packed, y, **fused))
        fused = {} if args.plain else dict(mask_src=x, mask_slope=.1)
        backward = timed(lambda: ops.conv_gemm_tc(geom, True, dy, packed_t, dx, **fused))
        if args.cycles:
            from promonet_b200 import _lib
            counters = torch.zeros(8, dtype=torch.int64, device='cuda')
            _lib.library().pmn_debug_train_tc_counters(counters.data_ptr())
            ops.conv_gemm_tc(geom, False, x, packed, y, a_act=ops.ACT_LRELU, a_slope=.1, residual=dy)
            torch.cuda.synchronize()
            _lib.library().pmn_debug_train_tc_counters(None)
            c = counters.tolist()
            steps = max(c[6], 1)
            print(f'    CTA 0 cycles: producers done {c[0]} (waiting for a stage {c[1]}, storing {c[3]}), '
                  f'MMA thread waited {c[5]}, accumulators done {c[7]}, epilogue done {c[4]}; {steps} K steps')
        line = (f'{name:22s} {flop / 1e9:8.2f} {forward:9.3f} {flop / forward / 1e9:7.1f} '
                f'{backward:9.3f} {flop / backward / 1e9:7.1f}')
        totals[0] += flop; totals[1] += forward; totals[2] += backward
        if args.wgrad:
            gw = torch.zeros_like(w)
            gb = torch.zeros(c_out, device='cuda')
            weight = timed(lambda: ops.conv_wgrad_tc(geom, dy, x, gw, gb, x_act=ops.ACT_LRELU, x_slope=.1))
            line += f' {weight:9.3f} {flop / weight / 1e9:7.1f}'
            totals[3] += weight
        print(line)
    print(f'{"total":22s} {totals[0] / 1e9:8.2f} {totals[1]:9.3f} {totals[0] / totals[1] / 1e9:7.1f} '
          f'{totals[2]:9.3f} {totals[0] / totals[2] / 1e9:7.1f}'
          + (f' {totals[3]:9.3f} {totals[0] / totals[3] / 1e9:7.1f}' if args.wgrad else ''))


if __name__ == '__main__':
    main()
