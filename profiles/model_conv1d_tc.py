"""A two-roof model of the synthesis step's residual-block convolutions (conv1d_tc_kernel),
compared with the measured per-layer times, and what the two kernels planned in DESIGN.md
section 7 would buy.  Runs anywhere (no GPU):

    python profiles/model_conv1d_tc.py > profiles/r1_model_conv1d_tc.txt

Model of one launch (B = 32 utterances, T time steps, C channels, k taps, fp32 parity = three
bf16 products per MAC):
  tensor roof  tiles of 128 time steps, 148 CTAs; per tap and per 16-channel K step the kernel
               issues x_hi . [w_hi | w_lo] (N = 2 C columns, split in two when 2 C > 256) and
               x_lo . w_hi (N = C); an M = 128 tcgen05.mma costs max(N / 2, 64) cycles: N / 2 is
               the tensor pipe (8192 dense bf16 FLOP / cycle / SM), 64 the fetch of its 4 KB A
               tile from shared memory (profiles/r1_tc_breakdown_*.txt);
               cycles are converted to time at the rate cuBLAS sustains on this GPU
               (MEASURED_PEAKS.json: bf16_tflops_sustained / 2250 nominal), i.e. "at the roof"
               means "as good as a cuBLAS GEMM of the same MMA count"
  HBM roof     c1: bf16 hi / lo planes in (4 B) and out (4 B) per element; c2: planes in (4 B),
               fp32 residual in (4 B), fp32 out (4 B), planes out (4 B); at the measured copy
               bandwidth (MEASURED_PEAKS.json)
Measured: profiles/r1_tc_breakdown_after_epilogue_fix.txt (CUDA events, one layer at a time).
"""
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BATCH, SMS, CLOCK_GHZ = 32, 148, 1.86
STAGES = ((256, 3440), (128, 27520), (64, 55040), (32, 110080))      # channels, time steps (5 s)
KERNELS, DILATIONS = (3, 7, 11), (1, 3, 5)
# A fused c1 -> c2 pair that reads the fp32 residual stream once (operand planes made in shared
# memory by the kernel, the residual taken from the same tile) and writes it once: 4 + 4 bytes per
# element.  Today: c1 planes in + out (8), c2 planes in + fp32 residual in + fp32 and planes out (16).
PAIR_BYTES = 8


def peaks():
    file = ROOT / 'MEASURED_PEAKS.json'
    data = json.loads(file.read_text()) if file.exists() else {}
    return data.get('hbm_gbs', 6549.), data.get('bf16_tflops_sustained', 1384.)


def hbm_gbs():
    return peaks()[0]


def tensor_efficiency():
    """fraction of the nominal 2.25 PFLOP/s (8192 FLOP / cycle / SM x 148 x 1.86 GHz) cuBLAS sustains"""
    return peaks()[1] / 2250.


def mma_cycles(columns):
    return max(columns / 2, 64)


def tensor_ms(channels, taps, t_len, columns_scale=1, rows=None, scheme='bf16x3'):
    """columns_scale = 2: the space-to-depth form (twice the channels over half the steps);
    scheme 'fp16+mxfp8x2': per 16 channels one fp16 MMA (N = C) and two halves of an fp8 MMA
    (K = 32, twice the rate: N / 2 cycles per 32 channels, same 64-cycle A fetch)"""
    c = channels * columns_scale
    if scheme == 'bf16x3':
        wide = [2 * c] if 2 * c <= 256 else [c, c]
        per_step = sum(mma_cycles(n) for n in wide) + mma_cycles(c)
    else:
        per_step = mma_cycles(c) + 2 * mma_cycles(c) / 2
    tiles = -(-(BATCH * (t_len // columns_scale if rows is None else rows)) // 128)
    cycles = -(-tiles // SMS) * taps * (c // 16) * per_step
    return cycles / (CLOCK_GHZ * 1e6) / tensor_efficiency()


def hbm_ms(channels, t_len, bytes_per_element):
    return BATCH * t_len * channels * bytes_per_element / (hbm_gbs() * 1e6)


def measured():
    table = {}
    file = ROOT / 'profiles' / 'r1_tc_breakdown_after_epilogue_fix.txt'
    for line in file.read_text().splitlines():
        match = re.match(r'C=\s*(\d+) k=\s*(\d+) (\w+)\s+([\d.]+) ms', line)
        if match:
            table[(int(match[1]), int(match[2]), match[3])] = float(match[4])
    return table


def main():
    times = measured()
    print(__doc__.split('\n\n')[0] + '\n')
    print(f'HBM {hbm_gbs():.0f} GB/s; tensor pipe at {tensor_efficiency():.2f} of nominal '
          f'({peaks()[1]:.0f} TFLOP/s sustained bf16); times in ms per launch\n')
    print(' C    k  layer | tensor roof  HBM roof   model | measured  measured / model')
    for channels, t_len in STAGES:
        for kernel in (3, 11):
            for layer, bytes_per_element in (('c1', 8), ('c2', 16)):
                tensor = tensor_ms(channels, kernel, t_len)
                hbm = hbm_ms(channels, t_len, bytes_per_element)
                model = max(tensor, hbm)
                real = times.get((channels, kernel, layer))
                print(f'{channels:3d}  {kernel:3d}  {layer:5s} | {tensor:10.3f} {hbm:9.3f} {model:7.3f} | '
                      f'{real:8.3f} {real / model:10.2f}' + ('   HBM-bound' if hbm > tensor else ''))
    # The whole step: 3 dilations x (c1, c2) per kernel size per stage; k = 7 interpolated from the
    # model ratio of its neighbours
    print('\nWhole step (72 launches), model roofs and what the planned kernels change:')
    header = (' C  | now: model  measured | fused c1+c2 pairs (8 B / element)  | + space-to-depth for d = 1'
              ' | + fp16 main, 2 x mxfp8 corrections')
    print(header)
    totals = [0., 0., 0., 0., 0.]
    for channels, t_len in STAGES:
        now = real = fused = s2d = two = 0.
        for kernel in KERNELS:
            near = 3 if kernel == 3 else 11
            for dilation in DILATIONS:
                pair_tensor = 2 * tensor_ms(channels, kernel, t_len)
                c1 = max(tensor_ms(channels, kernel, t_len), hbm_ms(channels, t_len, 8))
                c2 = max(tensor_ms(channels, kernel, t_len), hbm_ms(channels, t_len, 16))
                now += c1 + c2
                ratio = (times[(channels, near, 'c1')] + times[(channels, near, 'c2')]) / (
                    max(tensor_ms(channels, near, t_len), hbm_ms(channels, t_len, 8)) +
                    max(tensor_ms(channels, near, t_len), hbm_ms(channels, t_len, 16)))
                real += (c1 + c2) * ratio
                # fused pair: a tile keeps 128 - (k - 1) of its rows, traffic PAIR_BYTES / element
                halo = 128 / (128 - (kernel - 1))
                pair = max(pair_tensor * halo, hbm_ms(channels, t_len, PAIR_BYTES))
                # fuse only where it wins
                fused += min(pair, c1 + c2)
                # space-to-depth (c2 always, c1 when d = 1) for the narrow stages
                best = min(pair, c1 + c2)
                if channels <= 64:
                    taps = -(-(kernel + 2) // 2)
                    deep = tensor_ms(channels, taps, t_len, columns_scale=2)
                    plain = tensor_ms(channels, kernel, t_len)
                    first = deep if dilation == 1 else plain
                    best = min(best, max((first + deep) * halo, hbm_ms(channels, t_len, PAIR_BYTES)))
                s2d += best
                # 2 MMA units per MAC instead of 3, on the fused pair
                cheaper = 2 * tensor_ms(channels, kernel, t_len, scheme='fp16+mxfp8x2')
                two += min(best, max(cheaper * halo, hbm_ms(channels, t_len, PAIR_BYTES)))
        print(f'{channels:3d} | {now:10.2f} {real:9.2f} | {fused:32.2f} | {s2d:27.2f} | {two:33.2f}')
        for i, value in enumerate((now, real, fused, s2d, two)):
            totals[i] += value
    print('sum | {:10.2f} {:9.2f} | {:32.2f} | {:27.2f} | {:33.2f}'.format(*totals))
    flops = 3 * 8.117e12           # three bf16 products per MAC of the 8.117 TFLOP of residual blocks
    print(f'\n(no A-fetch floor, no HBM roof: {flops / (peaks()[1] * 1e12) * 1e3:.1f} ms per step for the '
          'residual blocks at the sustained bf16 rate)')


if __name__ == '__main__':
    main()
