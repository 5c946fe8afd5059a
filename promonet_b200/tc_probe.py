"""Drive single tensor-core conv launches (profiling scripts and tests)"""
import torch

from promonet_b200 import _lib


def run_tc_conv(batch, channels, t_len, kernel, mode='c2', dilation=1, repeats=1, f8=False):
    """Best CUDA-event ms of the conv kernel alone over `repeats` launches of
    pmn_conv1d_tc (timed through pmn_profile_*)

    mode: 'c1' planes out only; 'c2' residual + fp32 out + planes; 'c2acc'
    residual + MRF accumulate (the three epilogues of Block.forward); f8: pmn_conv1d_tc_f8"""
    lib = _lib.library()
    x = torch.randn(batch, channels, t_len, device='cuda')
    w = torch.randn(channels, channels, kernel, device='cuda') / (channels * kernel) ** .5
    bias = torch.randn(channels, device='cuda')
    out = torch.empty_like(x) if mode == 'c2' else None
    planes = torch.empty_like(x) if mode in ('c1', 'c2') else None
    residual = x if mode != 'c1' else None
    accum = torch.zeros_like(x) if mode == 'c2acc' else None
    size = lib.pmn_conv1d_tc_workspace_bytes(batch, channels, t_len, kernel)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    best = float('inf')
    for _ in range(repeats):
        _lib.profile(True)
        _lib.check((lib.pmn_conv1d_tc_f8 if f8 else lib.pmn_conv1d_tc)(
            x.data_ptr(), w.data_ptr(), bias.data_ptr(), _lib.ptr(residual), _lib.ptr(out),
            _lib.ptr(planes), _lib.ptr(accum), 2 if accum is not None else 0, 1 / 3,
            batch, channels, t_len, kernel, dilation, 0.1, 0.1,
            workspace.data_ptr(), size, _lib.stream()))
        torch.cuda.synchronize()
        ms, count = _lib.profile_read('conv1d_tc_kernel')
        if not count:   # the narrow layers' kernel (conv1d_tcw.cu) took the launch
            ms, count = _lib.profile_read('conv1d_tcw_kernel')
        best = min(best, ms)
    _lib.profile(False)
    return best
