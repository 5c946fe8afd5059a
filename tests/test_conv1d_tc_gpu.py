"""tcgen05 (bf16 x 3) dilated Conv1d against an fp64 torch convolution.
Tolerance 1e-4 relative to max|reference| (north_star); measured ~1e-5."""
import pytest
import torch

from conftest import relative_error

pytestmark = pytest.mark.gpu


def conv1d_tc(x, w, bias=None, residual=None, dilation=1, in_slope=1., out_slope=1.,
              want_planes=False, accum=None, accum_mode=0, accum_scale=1., want_out=True, f8=False):
    from promonet_b200 import _lib
    lib = _lib.library()
    batch, channels, t_len = x.shape
    k = w.shape[-1]
    xd, wd = x.cuda().contiguous(), w.cuda().contiguous()
    keep = [t.cuda().contiguous() if t is not None else None for t in (bias, residual)]
    out = torch.empty_like(xd) if want_out else None
    planes = torch.empty_like(xd) if want_planes else None
    size = lib.pmn_conv1d_tc_workspace_bytes(batch, channels, t_len, k)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    entry = lib.pmn_conv1d_tc_f8 if f8 else lib.pmn_conv1d_tc
    _lib.check(entry(
        xd.data_ptr(), wd.data_ptr(), *[_lib.ptr(t) for t in keep], _lib.ptr(out),
        _lib.ptr(planes), _lib.ptr(accum), accum_mode, accum_scale,
        batch, channels, t_len, k, dilation, in_slope, out_slope,
        workspace.data_ptr(), size, _lib.stream()))
    torch.cuda.synchronize()
    return out, planes


def reference(x, w, bias, dilation, in_slope):
    k = w.shape[-1]
    return torch.nn.functional.conv1d(
        torch.nn.functional.leaky_relu(x.double(), in_slope), w.double(),
        None if bias is None else bias.double(),
        padding=dilation * (k - 1) // 2, dilation=dilation)


@pytest.mark.parametrize('channels,k,dilation,t_len,batch', [
    (32, 3, 1, 512, 1), (32, 11, 5, 1500, 2), (64, 7, 3, 700, 3), (64, 3, 1, 256, 1),
    (128, 11, 1, 300, 2), (128, 11, 5, 1111, 1), (128, 7, 1, 256, 1),
    (256, 3, 5, 129, 2), (256, 11, 3, 1000, 1), (256, 7, 1, 128, 1),
    # the narrow layers' kernel (weights on the M side, 224-step tiles): shorter than the halo, one
    # tile exactly, one step into the next tile, the widest tap shifts, many tiles per CTA
    (64, 11, 5, 30, 2), (64, 11, 1, 224, 1), (32, 11, 5, 225, 3), (64, 7, 5, 449, 2),
    (32, 11, 1, 100, 1), (64, 3, 5, 2000, 1), (32, 11, 3, 224 * 160 + 7, 1)])
def test_conv1d_tc_matches_fp64(channels, k, dilation, t_len, batch):
    torch.manual_seed(channels + k + dilation)
    x = torch.randn(batch, channels, t_len)
    w = torch.randn(channels, channels, k) / (channels * k) ** .5
    bias = torch.randn(channels)
    expected = reference(x, w, bias, dilation, 0.1)
    out, _ = conv1d_tc(x, w, bias, dilation=dilation, in_slope=0.1)
    error = relative_error(out, expected)
    assert error < 1e-4, error


@pytest.mark.parametrize('channels,k,dilation,t_len,batch', [
    (128, 11, 1, 300, 2), (128, 11, 5, 1111, 1), (128, 7, 3, 256, 1), (128, 3, 1, 700, 2),
    (256, 3, 5, 129, 2), (256, 11, 3, 1000, 1), (256, 7, 1, 128, 1), (256, 11, 5, 256 * 150 + 3, 1)])
@pytest.mark.parametrize('weight_scale', [1., 1e-2])
def test_conv1d_tc_f8_matches_fp64(channels, k, dilation, t_len, batch, weight_scale):
    """"fp16 + 2 x fp8" operands (pmn_conv1d_tc_f8): x w = fp16 x fp16 + two e4m3 correction products
    in one accumulator.  Same 1e-4 bar; measured 1-2e-5 on one layer (bf16 x 3: ~5e-6).  The weight
    scale exercises the power-of-two weight shift"""
    torch.manual_seed(channels + k + dilation)
    x = torch.randn(batch, channels, t_len)
    w = weight_scale * torch.randn(channels, channels, k) / (channels * k) ** .5
    bias = weight_scale * torch.randn(channels)
    expected = reference(x, w, bias, dilation, 0.1)
    out, _ = conv1d_tc(x, w, bias, dilation=dilation, in_slope=0.1, f8=True)
    error = relative_error(out, expected)
    assert error < 5e-5, error


@pytest.mark.parametrize('channels', [128, 256])
def test_conv1d_tc_f8_epilogue(channels):
    """Residual, MRF accumulate and the "fp16 + 2 x fp8" operand written by the epilogue"""
    torch.manual_seed(3)
    x = torch.randn(2, channels, 600)
    w = torch.randn(channels, channels, 7) / (channels * 7) ** .5
    bias, residual = torch.randn(channels), torch.randn(2, channels, 600)
    y = reference(x, w, bias, 3, 0.1) + residual.double()
    accum = torch.ones(2, channels, 600, device='cuda')
    out, planes = conv1d_tc(
        x, w, bias, residual, dilation=3, in_slope=0.1, out_slope=0.1, want_planes=True,
        accum=accum, accum_mode=2, accum_scale=1 / 3, f8=True)
    assert relative_error(out, y) < 5e-5
    assert relative_error(accum, 1 + y / 3) < 5e-5
    # fp16 part + e4m3 low part: 11 + 3 mantissa bits of lrelu(y) per element
    expected = torch.nn.functional.leaky_relu(out.double().cpu(), 0.1)
    worst = ((planes.double().cpu() - expected).abs() / expected.abs().clamp(min=1e-2)).max()
    assert float(worst) < 2 ** -13, float(worst)


def test_conv1d_tc_f8_saturates_instead_of_overflowing():
    """|x| beyond the e4m3 / fp16 ranges of the operand must stay finite (saturating converts)"""
    torch.manual_seed(4)
    x = torch.randn(1, 128, 300)
    x[0, 5, 100] = 3000.
    x[0, 9, 7] = -100.
    w = torch.randn(128, 128, 3) / (128 * 3) ** .5
    out, _ = conv1d_tc(x, w, dilation=1, f8=True)
    assert bool(torch.isfinite(out).all())


def test_conv1d_tc_many_tiles_per_cta():
    """More tiles than SMs: exercises ring phase wrap-around and TMEM double buffering"""
    torch.manual_seed(1)
    x = torch.randn(8, 128, 256 * 40)
    w = torch.randn(128, 128, 3) / (128 * 3) ** .5
    expected = reference(x, w, None, 1, 1.)
    out, _ = conv1d_tc(x, w, dilation=1)
    assert relative_error(out, expected) < 1e-4


@pytest.mark.parametrize('channels', [64, 32, 128])
def test_conv1d_tc_epilogue(channels):
    """64: conv1d_tcw_kernel; 32 (k = 7 with a residual): conv1d_tc_kernel, CONCAT; 128: plain"""
    torch.manual_seed(2)
    x = torch.randn(2, channels, 600)
    w = torch.randn(channels, channels, 7) / (channels * 7) ** .5
    bias, residual = torch.randn(channels), torch.randn(2, channels, 600)
    y = reference(x, w, bias, 3, 0.1) + residual.double()
    accum = torch.ones(2, channels, 600, device='cuda')
    out, planes = conv1d_tc(
        x, w, bias, residual, dilation=3, in_slope=0.1, out_slope=0.1, want_planes=True,
        accum=accum, accum_mode=2, accum_scale=1 / 3)
    assert relative_error(out, y) < 1e-4
    assert relative_error(planes, torch.nn.functional.leaky_relu(y, 0.1)) < 1e-4
    assert relative_error(accum, 1. + y / 3) < 1e-4
    accum = torch.empty(2, channels, 600, device='cuda')
    out, _ = conv1d_tc(x, w, bias, dilation=3, in_slope=0.1, accum=accum, accum_mode=1,
                       accum_scale=0.5, want_out=False)
    assert out is None
    assert relative_error(accum, 0.5 * reference(x, w, bias, 3, 0.1)) < 1e-4


@pytest.mark.parametrize('tag,kernel', [('c32k3', 3), ('c64k11', 11)])
def test_block_tc_matches_reference_golden(golden, tag, kernel):
    """Block.forward hifigan.py:198-210 chained through the hi/lo planes"""
    from oracle import hifigan
    g = golden('block')
    cur = g[f'{tag}_x']
    for i, dilation in enumerate((1, 3, 5)):
        def folded(name):
            return hifigan.fold_weight_norm(
                g[f'{tag}_{name}.{i}.weight_g'], g[f'{tag}_{name}.{i}.weight_v'])
        xt, _ = conv1d_tc(cur, folded('convs1'), g[f'{tag}_convs1.{i}.bias'],
                          dilation=dilation, in_slope=0.1)
        cur, _ = conv1d_tc(xt, folded('convs2'), g[f'{tag}_convs2.{i}.bias'],
                           residual=cur, in_slope=0.1)
        cur = cur.cpu()
    assert relative_error(cur, g[f'{tag}_y']) < 1e-4


@pytest.mark.parametrize('c_in,k,s,t,batch', [
    (512, 16, 8, 86, 2), (256, 16, 8, 430, 1), (128, 4, 2, 1000, 2), (64, 4, 2, 2049, 1),
    (512, 16, 8, 5, 1)])
def test_conv_transpose1d_tc_matches_fp64(c_in, k, s, t, batch):
    """MultiReceptiveFieldFusion upsampler hifigan.py:97-106 on the tensor cores"""
    from promonet_b200 import _lib
    lib = _lib.library()
    torch.manual_seed(c_in + t)
    c_out = c_in // 2
    x = torch.randn(batch, c_in, t)
    w = torch.randn(c_in, c_out, k) / c_in ** .5
    b = torch.randn(c_out)
    expected = torch.nn.functional.conv_transpose1d(
        torch.nn.functional.leaky_relu(x.double(), 0.1), w.double(), b.double(),
        stride=s, padding=(k - s) // 2)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = torch.empty(batch, c_out, s * t, device='cuda')
    size = lib.pmn_conv_transpose1d_tc_workspace_bytes(batch, c_in, t, s)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    _lib.check(lib.pmn_conv_transpose1d_tc(
        xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(),
        batch, c_in, c_out, t, k, s, 0.1, workspace.data_ptr(), size, _lib.stream()))
    error = relative_error(out, expected)
    assert error < 1e-4, error


@pytest.mark.parametrize('c_in,c_out,k,t_len,batch', [
    (256, 32, 32, 481 * 3, 1), (32, 32, 32, 225 * 5, 1), (32, 128, 32, 97 * 7, 2),
    (128, 256, 32, 66 * 9, 1), (256, 32, 5, 300, 2), (32, 128, 1, 130, 1)])
def test_conv1d_tc_valid_relu_rectangular(c_in, c_out, k, t_len, batch):
    """penn FCNF0++ block shapes: valid convolution + ReLU, C_in != C_out"""
    from promonet_b200 import _lib
    lib = _lib.library()
    torch.manual_seed(c_in + c_out + k)
    x = torch.randn(batch, c_in, t_len)
    w = torch.randn(c_out, c_in, k) / (c_in * k) ** .5
    bias = torch.randn(c_out)
    expected = torch.relu(torch.nn.functional.conv1d(x.double(), w.double(), bias.double()))
    xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
    out = torch.empty(batch, c_out, t_len - k + 1, device='cuda')
    size = lib.pmn_conv1d_tc_workspace_bytes(batch, c_in, t_len, k)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    _lib.check(lib.pmn_conv1d_tc_general(
        xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), None, out.data_ptr(), None, None, 0, 1.,
        batch, c_in, c_out, t_len, k, 1, 1, 1, 1., 1., workspace.data_ptr(), size, _lib.stream()))
    error = relative_error(out, expected)
    assert error < 1e-4, error
