"""Does the tap row offset alignment of the A operand matter?  Same conv (k = 7, c1
epilogue) with dilation 1 (offsets 0..6 rows, unaligned) and dilation 8 (offsets
0, 8, .., 48: every 8-row core matrix 128 B aligned)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200.tc_probe import run_tc_conv  # noqa: E402

for channels, t_len in ((256, 3440), (128, 27520), (64, 55040), (32, 110080)):
    for dilation in (1, 8, 1, 8):
        ms = run_tc_conv(32, channels, t_len, 7, 'c1', dilation=dilation, repeats=3)
        flops = 2 * 32 * channels * channels * 7 * t_len
        print(f'C={channels:3d} k=7 dilation={dilation} {ms:7.3f} ms {flops / ms / 1e9:7.1f} TFLOP/s')
