import sys
sys.path.insert(0, '/root/repo')
import torch
from promonet_b200.train import ops
from profiles.bench_conv_tc import timed
for channels, k, t in ((32, 3, 16384), (32, 3, 16384 + 384), (32, 3, 16384 + 32), (128, 11, 4096), (128, 11, 4096 + 128), (128, 11, 4096 + 32)):
    B = 8
    geom = ops.geometry(B, channels, channels, (t, 1), (k, 1), 1, 1, ((k - 1) // 2, 0))
    x = torch.randn(B, channels, t, device='cuda'); w = torch.randn(channels, channels, k, device='cuda')
    y = torch.empty_like(x); r = torch.randn_like(x)
    packed = ops.pack_weight_taps(w, torch.empty(ops.packed_floats(channels, channels, k), device='cuda'), channels, channels, k, False)
    ms = timed(lambda: ops.conv_gemm_tc(geom, False, x, packed, y, a_act=ops.ACT_LRELU, a_slope=.1, residual=r), 10)
    flop = 2. * B * t * channels * channels * k
    print(channels, k, t, f'{ms*1e3:.1f} us', f'{flop/ms/1e9:.1f} TF/s')
