"""The narrow residual-block convolutions (C = 32 / 64 at the synthesis benchmark's shapes, batch 32)
on conv1d_tc_kernel (time on the M side) against conv1d_tcw_kernel (weights on the M side):

    python profiles/narrow_layers.py            # runs itself with PMN_TCW=0 and PMN_TCW=1
"""
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

SHAPES = [(32, 110080), (64, 55040)]


def measure():
    from promonet_b200.tc_probe import run_tc_conv
    for channels, t_len in SHAPES:
        for kernel in (3, 7, 11):
            for dilation, mode in ((1, 'c1'), (5, 'c1'), (1, 'c2'), (1, 'c2acc')):
                ms = run_tc_conv(32, channels, t_len, kernel, mode, dilation, repeats=3)
                flop = 2. * 32 * t_len * channels * channels * kernel
                print(f'{channels} {kernel} {dilation} {mode} {ms:.4f} {flop / ms / 1e9:.1f}', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--measure':
        measure()
        sys.exit(0)
    results = {}
    for flag in ('0', '2'):
        env = dict(os.environ, PMN_TCW=flag)
        out = subprocess.run([sys.executable, __file__, '--measure'], env=env, capture_output=True, text=True)
        if out.returncode:
            print(out.stderr)
            sys.exit(1)
        for line in out.stdout.splitlines():
            c, k, d, mode, ms, tflops = line.split()
            results.setdefault((int(c), int(k), int(d), mode), {})[flag] = (float(ms), float(tflops))
    print(f'{"C":>3s} {"k":>2s} {"d":>2s} {"mode":>6s} | {"tc ms":>8s} {"TFLOP/s":>8s} | {"tcw ms":>8s} {"TFLOP/s":>8s} | ratio')
    for (c, k, d, mode), r in results.items():
        print(f'{c:3d} {k:2d} {d:2d} {mode:>6s} | {r["0"][0]:8.4f} {r["0"][1]:8.1f} | {r["2"][0]:8.4f} {r["2"][1]:8.1f} | '
              f'{r["0"][0] / r["2"][0]:.2f}')
