"""One tcgen05 conv launch per generator stage (B = 32, 5 s), for ncu:

    ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel \
        -o gpurun_out/conv1d_tc python profiles/profile_tc.py

Shapes: (C, T) of the four stages of config/promonet.py; k = 11, dilation 1.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200 import _lib  # noqa: E402

lib = _lib.library()
kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 11
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for channels, t_len in ((256, 3440), (128, 27520), (64, 55040), (32, 110080)):
    x = torch.randn(32, channels, t_len, device='cuda')
    w = torch.randn(channels, channels, kernel, device='cuda') / (channels * kernel) ** .5
    bias = torch.randn(channels, device='cuda')
    out = torch.empty_like(x)
    planes = torch.empty_like(x)
    size = lib.pmn_conv1d_tc_workspace_bytes(32, channels, t_len, kernel)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    for _ in range(repeats):
        _lib.check(lib.pmn_conv1d_tc(
            x.data_ptr(), w.data_ptr(), bias.data_ptr(), x.data_ptr(), out.data_ptr(),
            planes.data_ptr(), None, 0, 1., 32, channels, t_len, kernel, 1, 0.1, 0.1,
            workspace.data_ptr(), size, _lib.stream()))
    torch.cuda.synchronize()
    del x, w, out, planes, workspace
