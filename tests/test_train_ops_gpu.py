"""Training-step operators (C ABI, include/promonet_b200.h "Training step") against
plain PyTorch fp32/fp64 references of the same op on the CPU.  Tolerance 1e-4
relative to max|reference| unless stated (fp32 accumulation order differs)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import relative_error

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-4


@pytest.fixture(scope='module')
def ops():
    from promonet_b200.train import ops
    return ops


# (c_in, c_out, size_in, kernel, stride, dilation, padding): the shapes of the hot path
CONVS = [
    # HiFi-GAN dilated Conv1d as (k, 1) over (T, 1): hifigan.py:167-183
    (32, 32, (700, 1), (11, 1), 1, (5, 1), (25, 0)),
    (64, 64, (300, 1), (7, 1), 1, (3, 1), (9, 0)),
    (113, 512, (64, 1), (7, 1), 1, 1, (3, 0)),
    (32, 1, (900, 1), (7, 1), 1, 1, (3, 0)),
    # MPD Conv2d (5, 1) stride (3, 1): discriminator.py:67-72
    (1, 32, (301, 3), (5, 1), (3, 1), 1, (2, 0)),
    (32, 128, (101, 5), (5, 1), (3, 1), 1, (2, 0)),
    (128, 96, (34, 2), (5, 1), 1, 1, (2, 0)),
    (96, 1, (34, 7), (3, 1), 1, 1, (1, 0)),
    # CMB Conv2d (3, 9), stride (1, 2): discriminator.py:160-170
    (1, 32, (16, 77), (3, 9), (1, 1), 1, (1, 4)),
    (32, 32, (16, 51), (3, 9), (1, 2), 1, (1, 4)),
    (32, 1, (16, 40), (3, 3), 1, 1, (1, 1)),
    # strided Conv1d: the data gradient of the upsampling ConvTranspose1d
    (16, 24, (96, 1), (16, 1), (8, 1), 1, (4, 0)),
]


@pytest.mark.parametrize('c_in,c_out,size,kernel,stride,dilation,padding', CONVS)
def test_conv_forward_dgrad_wgrad(ops, c_in, c_out, size, kernel, stride, dilation, padding):
    torch.manual_seed(0)
    batch = 3
    x = torch.randn(batch, c_in, *size)
    w = torch.randn(c_out, c_in, *kernel) / (c_in * kernel[0] * kernel[1]) ** .5
    bias = torch.randn(c_out)
    residual_slope, out_slope = .1, .2

    # reference: y = lrelu_out(conv(lrelu_in(x)) + bias), loss = <y, dy>
    xr = x.double().requires_grad_()
    wr = w.double().requires_grad_()
    br = bias.double().requires_grad_()
    y = F.leaky_relu(
        F.conv2d(F.leaky_relu(xr, residual_slope), wr, br, stride, padding, dilation), out_slope)
    dy = torch.randn(y.shape)
    (y * dy.double()).sum().backward()

    geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
    assert (geom.h_out, geom.w_out) == tuple(y.shape[2:])
    xd, wd, bd, dyd = (t.cuda().contiguous() for t in (x, w, bias, dy))
    out = torch.empty(y.shape, device='cuda')
    ops.conv_gemm(geom, False, xd, wd, out, a_act=ops.ACT_LRELU, a_slope=residual_slope,
                  bias=bd, out_act=ops.OUT_LRELU, out_slope=out_slope)
    assert relative_error(out, y) < TOLERANCE

    taps = kernel[0] * kernel[1]
    wt = ops.transpose_weight(wd, torch.empty(c_in, c_out, *kernel, device='cuda'), c_out, c_in, taps)
    assert torch.equal(wt.cpu(), w.transpose(0, 1).contiguous())
    dx = torch.empty(x.shape, device='cuda')
    ops.conv_gemm(geom, True, dyd, wt, dx, a_companion=out, a_act=ops.ACT_LRELU_MASK,
                  a_slope=out_slope, mask_src=xd, mask_slope=residual_slope)
    assert relative_error(dx, xr.grad) < TOLERANCE

    gw = torch.zeros(w.shape, device='cuda')
    gb = torch.zeros(c_out, device='cuda')
    ops.conv_wgrad(geom, dyd, xd, gw, gb, dy_companion=out, dy_act=ops.ACT_LRELU_MASK,
                   dy_slope=out_slope, x_act=ops.ACT_LRELU, x_slope=residual_slope)
    assert relative_error(gw, wr.grad) < TOLERANCE
    assert relative_error(gb, br.grad) < TOLERANCE


TF32_TOLERANCE = 3e-3   # tf32 operands: 2^-11 relative rounding per operand, fp32 accumulation


@pytest.mark.parametrize('c_in,c_out,size,kernel,stride,dilation,padding', CONVS)
def test_conv_tensor_core_forward_and_dgrad(ops, c_in, c_out, size, kernel, stride, dilation, padding):
    """The tcgen05 (tf32) implicit GEMM against the same fp64 reference"""
    torch.manual_seed(0)
    batch = 3
    x = torch.randn(batch, c_in, *size)
    w = torch.randn(c_out, c_in, *kernel) / (c_in * kernel[0] * kernel[1]) ** .5
    bias = torch.randn(c_out)
    in_slope, out_slope = .1, .2
    xr = x.double().requires_grad_()
    y = F.leaky_relu(
        F.conv2d(F.leaky_relu(xr, in_slope), w.double(), bias.double(), stride, padding, dilation),
        out_slope)
    dy = torch.randn(y.shape)
    (y * dy.double()).sum().backward()
    residual = torch.randn(y.shape)

    geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
    taps = kernel[0] * kernel[1]
    xd, wd, bd, dyd = (t.cuda().contiguous() for t in (x, w, bias, dy))
    packed = ops.pack_weight_taps(
        wd, torch.empty(ops.packed_floats(c_out, c_in, taps), device='cuda'), c_out, c_in, taps, False)
    # [row tile][tap][channel block][k / 4][row in tile][4], rounded to tf32 (10-bit mantissa)
    bn = 128 if c_out > 64 else (64 if c_out > 32 else 32)
    tiles, blocks = -(-c_out // bn), ops.channel_pad(c_in) // 32
    expected = torch.zeros(tiles * bn, taps, blocks * 32)
    expected[:c_out, :, :c_in] = w.flatten(2).permute(0, 2, 1)
    expected = expected.view(tiles, bn, taps, blocks, 8, 4).permute(0, 2, 3, 4, 1, 5).reshape(-1)
    assert float((packed.cpu() - expected).abs().max()) <= 2 ** -11 * float(expected.abs().max())
    out = torch.full(y.shape, float('nan'), device='cuda')
    ops.conv_gemm_tc(geom, False, xd, packed, out, a_act=ops.ACT_LRELU, a_slope=in_slope,
                     bias=bd, out_act=ops.OUT_LRELU, out_slope=out_slope)
    assert relative_error(out, y) < TF32_TOLERANCE
    # epilogue: residual, alpha, accumulate
    out2 = out.clone()
    ops.conv_gemm_tc(geom, False, xd, packed, out2, a_act=ops.ACT_LRELU, a_slope=in_slope,
                     bias=bd, residual=residual.cuda(), alpha=.5, accumulate=True)
    pre = F.conv2d(F.leaky_relu(x.double(), in_slope), w.double(), bias.double(), stride, padding, dilation)
    assert relative_error(out2, y.detach() + .5 * (pre + residual.double())) < TF32_TOLERANCE

    packed_t = ops.pack_weight_taps(
        wd, torch.empty(ops.packed_floats(c_in, c_out, taps), device='cuda'), c_out, c_in, taps, True)
    dx = torch.full(x.shape, float('nan'), device='cuda')
    exact = torch.empty(y.shape, device='cuda')
    ops.conv_gemm(geom, False, xd, wd, exact, a_act=ops.ACT_LRELU, a_slope=in_slope,
                  bias=bd, out_act=ops.OUT_LRELU, out_slope=out_slope)
    ops.conv_gemm_tc(geom, True, dyd, packed_t, dx, a_companion=exact, a_act=ops.ACT_LRELU_MASK,
                     a_slope=out_slope, mask_src=xd, mask_slope=in_slope)
    assert relative_error(dx, xr.grad) < TF32_TOLERANCE


EXACT = [
    # c_in, c_out, size, kernel, stride, dilation, padding, batch: every pipeline shape case
    (32, 32, (2048, 1), (3, 1), 1, (1, 1), (1, 0), 2),       # 3 K steps (odd), BN 32
    (32, 32, (2048, 1), (11, 1), 1, (5, 1), (25, 0), 2),     # 11 K steps, halo 25
    (64, 64, (1024, 1), (7, 1), 1, (3, 1), (9, 0), 2),       # BN 64, 14 K steps
    (256, 256, (64, 1), (3, 1), 1, (1, 1), (1, 0), 2),       # one M tile, two N tiles, 24 K steps
    (113, 512, (8, 1), (7, 1), 1, 1, (3, 0), 2),             # padded channels, 16 rows only
    (128, 512, (38, 2), (5, 1), (3, 1), 1, (2, 0), 4),       # strided, M = 104 < one tile
    (64, 512, (1700, 2), (5, 1), 1, 1, (2, 0), 4),           # 107 M tiles x 2: the 256-wide CTAs
    (32, 32, (8, 77), (3, 9), (1, 2), 1, (1, 4), 4),         # 27 taps
    (258, 512, (1, 1), (1, 1), 1, 1, 0, 2),                  # speaker projection: 2 rows
]


@pytest.mark.parametrize('c_in,c_out,size,kernel,stride,dilation,padding,batch', EXACT)
def test_conv_tensor_core_is_exact_on_small_integers(
        ops, c_in, c_out, size, kernel, stride, dilation, padding, batch):
    """Integer operands are exact in tf32 and their sums exact in fp32, so the tcgen05 path
    must reproduce the fp64 reference bit for bit: any pipeline race or indexing slip shows"""
    torch.manual_seed(12)
    x = torch.randint(-3, 4, (batch, c_in, *size)).float()
    w = torch.randint(-2, 3, (c_out, c_in, *kernel)).float()
    y = F.conv2d(x.double(), w.double(), None, stride, padding, dilation)
    dy = torch.randint(-3, 4, y.shape).float()
    dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), stride, padding, dilation)
    geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
    taps = kernel[0] * kernel[1]
    xd, wd, dyd = x.cuda(), w.cuda(), dy.cuda()
    packed = ops.pack_weight_taps(
        wd, torch.empty(ops.packed_floats(c_out, c_in, taps), device='cuda'), c_out, c_in, taps, False)
    packed_t = ops.pack_weight_taps(
        wd, torch.empty(ops.packed_floats(c_in, c_out, taps), device='cuda'), c_out, c_in, taps, True)
    for _ in range(3):
        out = torch.full(y.shape, float('nan'), device='cuda')
        ops.conv_gemm_tc(geom, False, xd, packed, out)
        assert torch.equal(out.cpu().double(), y)
        grad = torch.full(x.shape, float('nan'), device='cuda')
        ops.conv_gemm_tc(geom, True, dyd, packed_t, grad)
        assert torch.equal(grad.cpu().double(), dx)


@pytest.mark.parametrize('c_in,c_out,size,kernel,stride,dilation,padding,batch', EXACT)
def test_wgrad_tensor_core_is_exact_on_small_integers(
        ops, c_in, c_out, size, kernel, stride, dilation, padding, batch):
    torch.manual_seed(13)
    x = torch.randint(-2, 3, (batch, c_in, *size)).float()
    geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
    dy = torch.randint(-2, 3, (batch, c_out, geom.h_out, geom.w_out)).float()
    gw = torch.nn.grad.conv2d_weight(
        x.double(), (c_out, c_in, *kernel), dy.double(), stride, padding, dilation)
    gb = dy.double().sum((0, 2, 3))
    for _ in range(2):
        out = torch.zeros(c_out, c_in, *kernel, device='cuda')
        bias = torch.zeros(c_out, device='cuda')
        ops.conv_wgrad_tc(geom, dy.cuda(), x.cuda(), out, bias)
        assert torch.equal(out.cpu().double(), gw)
        assert torch.equal(bias.cpu().double(), gb)


@pytest.mark.parametrize('c_in,c_out,size,kernel,stride,dilation,padding', CONVS)
def test_wgrad_tensor_core_with_activations(ops, c_in, c_out, size, kernel, stride, dilation, padding):
    """The activation pairs of the training step against fp64 autograd"""
    torch.manual_seed(14)
    batch, slope = 3, .1
    x = torch.randn(batch, c_in, *size)
    w = torch.randn(c_out, c_in, *kernel) / (c_in * kernel[0] * kernel[1]) ** .5
    geom = ops.geometry(batch, c_in, c_out, size, kernel, stride, dilation, padding)
    # (NONE, LRELU): y = conv(lrelu(x)); (LRELU_MASK, NONE): y = lrelu(conv(x))
    for dy_act, x_act in ((ops.ACT_NONE, ops.ACT_LRELU), (ops.ACT_LRELU_MASK, ops.ACT_NONE)):
        wr = w.double().requires_grad_()
        br = torch.zeros(c_out, dtype=torch.double, requires_grad=True)
        source = F.leaky_relu(x.double(), slope) if x_act == ops.ACT_LRELU else x.double()
        y = F.conv2d(source, wr, br, stride, padding, dilation)
        if dy_act == ops.ACT_LRELU_MASK:
            y = F.leaky_relu(y, slope)
        dy = torch.randn(y.shape)
        (y * dy.double()).sum().backward()
        gw = torch.zeros(w.shape, device='cuda')
        gb = torch.zeros(c_out, device='cuda')
        ops.conv_wgrad_tc(
            geom, dy.cuda(), x.cuda(), gw, gb,
            dy_companion=y.detach().float().cuda() if dy_act == ops.ACT_LRELU_MASK else None,
            dy_act=dy_act, dy_slope=slope, x_act=x_act, x_slope=slope)
        assert relative_error(gw, wr.grad) < TF32_TOLERANCE
        assert relative_error(gb, br.grad) < TF32_TOLERANCE


def test_conv_tensor_core_large_shapes(ops):
    """Several M and N tiles, K = 5 x 1024 (MPD conv4, discriminator.py:72)"""
    torch.manual_seed(11)
    x = torch.randn(4, 1024, 51, 3)
    w = torch.randn(1024, 1024, 5, 1) / (1024 * 5) ** .5
    bias = torch.randn(1024)
    y = F.conv2d(x.double(), w.double(), bias.double(), 1, (2, 0))
    geom = ops.geometry(4, 1024, 1024, (51, 3), (5, 1), 1, 1, (2, 0))
    packed = ops.pack_weight_taps(
        w.cuda(), torch.empty(ops.packed_floats(1024, 1024, 5), device='cuda'), 1024, 1024, 5, False)
    out = torch.empty(y.shape, device='cuda')
    ops.conv_gemm_tc(geom, False, x.cuda(), packed, out, bias=bias.cuda())
    assert relative_error(out, y) < TF32_TOLERANCE


def test_conv_epilogue_residual_alpha_accumulate_tanh(ops):
    torch.manual_seed(1)
    x = torch.randn(2, 24, 130, 1)
    w = torch.randn(24, 24, 3, 1) / 8
    bias2 = torch.randn(2, 24)
    residual = torch.randn(2, 24, 130, 1)
    previous = torch.randn(2, 24, 130, 1)
    geom = ops.geometry(2, 24, 24, (130, 1), (3, 1), 1, 1, (1, 0))
    out = previous.cuda().clone()
    ops.conv_gemm(geom, False, x.cuda(), w.cuda(), out, bias2=bias2.cuda(),
                  residual=residual.cuda(), alpha=1 / 3, accumulate=True)
    expected = previous + (F.conv2d(x, w, None, 1, (1, 0)) + bias2[:, :, None, None] + residual) / 3
    assert relative_error(out, expected) < TOLERANCE

    out = torch.empty(2, 24, 130, 1, device='cuda')
    ops.conv_gemm(geom, False, x.cuda(), w.cuda(), out, out_act=ops.OUT_TANH)
    y = torch.tanh(F.conv2d(x, w, None, 1, (1, 0)))
    assert relative_error(out, y) < TOLERANCE
    # tanh mask on the gradient operand
    dy = torch.randn(y.shape)
    gw = torch.zeros(w.shape, device='cuda')
    ops.conv_wgrad(geom, dy.cuda(), x.cuda(), gw, None, dy_companion=out, dy_act=ops.ACT_TANH_MASK)
    wr = w.double().requires_grad_()
    (torch.tanh(F.conv2d(x.double(), wr, None, 1, (1, 0))) * dy.double()).sum().backward()
    assert relative_error(gw, wr.grad) < TOLERANCE


@pytest.mark.parametrize('c_in,c_out,k,stride,t', [(64, 32, 16, 8, 40), (32, 16, 4, 2, 300)])
def test_conv_transpose_forward_and_gradients(ops, c_in, c_out, k, stride, t):
    """LeakyReLU + ConvTranspose1d (hifigan.py:97-106): forward through the transposed
    gather, data gradient as a strided convolution, weight gradient with exchanged roles"""
    torch.manual_seed(2)
    batch, pad, slope = 2, (k - stride) // 2, .1
    x = torch.randn(batch, c_in, t)
    w = torch.randn(c_in, c_out, k) / (c_in * k / stride) ** .5
    bias = torch.randn(c_out)
    xr, wr, br = (v.double().requires_grad_() for v in (x, w, bias))
    y = F.conv_transpose1d(F.leaky_relu(xr, slope), wr, br, stride, pad)
    dy = torch.randn(y.shape)
    (y * dy.double()).sum().backward()

    # the convolution this transposes: C_out channels over s*T -> C_in channels over T
    geom = ops.geometry(batch, c_out, c_in, (t * stride, 1), (k, 1), (stride, 1), 1, (pad, 0),
                        size_out=(t, 1))
    xd, wd, bd, dyd = (v.cuda().contiguous() for v in (x, w, bias, dy))
    wt = ops.transpose_weight(wd, torch.empty(c_out, c_in, k, device='cuda'), c_in, c_out, k)
    out = torch.empty(y.shape, device='cuda')
    ops.conv_gemm(geom, True, xd, wt, out, a_act=ops.ACT_LRELU, a_slope=slope, bias=bd)
    assert relative_error(out, y) < TOLERANCE
    fast = ops.conv_transpose1d(xd, wd, bd, stride, slope)
    assert relative_error(fast, y) < TOLERANCE

    dx = torch.empty(x.shape, device='cuda')
    ops.conv_gemm(geom, False, dyd, wd, dx, mask_src=xd, mask_slope=slope)
    assert relative_error(dx, xr.grad) < TOLERANCE

    gw = torch.zeros(w.shape, device='cuda')
    ops.conv_wgrad(geom, xd, dyd, gw, None, dy_act=ops.ACT_LRELU, dy_slope=slope)
    assert relative_error(gw, wr.grad) < TOLERANCE
    gb = torch.empty(c_out, device='cuda')
    ops.row_sum(dyd.transpose(0, 1).contiguous(), gb, c_out, batch * t * stride)
    assert relative_error(gb, br.grad) < TOLERANCE


def test_weight_norm_backward(ops):
    torch.manual_seed(3)
    v = torch.randn(48, 40 * 5)
    g = torch.rand(48) + .5
    gw = torch.randn(48, 40 * 5)
    vr, gr = v.double().requires_grad_(), g.double().requires_grad_()
    w = gr[:, None] * vr / vr.norm(dim=1, keepdim=True)
    (w * gw.double()).sum().backward()
    folded = ops.weight_norm_fold(v.cuda(), g.cuda(), torch.empty(48, 200, device='cuda'), 48, 200)
    assert relative_error(folded, w) < 1e-6
    gv, gg = torch.empty(48, 200, device='cuda'), torch.empty(48, device='cuda')
    ops.weight_norm_backward(v.cuda(), g.cuda(), gw.cuda(), gv, gg, 48, 200)
    assert relative_error(gv, vr.grad) < 1e-5
    assert relative_error(gg, gr.grad) < 1e-5


@pytest.mark.parametrize('t,left,right', [(16384, 0, 1), (16384, 0, 9), (700, 384, 384), (50, 0, 0)])
def test_reflect_pad_and_adjoint(ops, t, left, right):
    torch.manual_seed(4)
    x = torch.randn(3, 2, t)
    xr = x.double().requires_grad_()
    y = F.pad(xr, (left, right), 'reflect')
    out = ops.reflect_pad(x.cuda(), left, right)
    assert torch.equal(out.cpu(), y.detach().float())
    dy = torch.randn(y.shape)
    (y * dy.double()).sum().backward()
    gx = ops.reflect_pad_backward(dy.cuda(), torch.empty(x.shape, device='cuda'), left, right)
    assert relative_error(gx, xr.grad) < 1e-6
    gx = ops.reflect_pad_backward(dy.cuda(), torch.ones(x.shape, device='cuda'), left, right, True)
    assert relative_error(gx, xr.grad + 1) < 1e-6


def test_losses(ops):
    torch.manual_seed(5)
    x = torch.randn(4, 1000)
    xr = x.double().requires_grad_()
    loss = torch.zeros(1, device='cuda')
    grad = torch.empty(x.shape, device='cuda')
    ops.mse_to_target(x.cuda(), 1., 1., loss, grad)
    expected = torch.mean((1. - xr) ** 2.)
    expected.backward()
    assert relative_error(loss, expected.reshape(1)) < 1e-5
    assert relative_error(grad, xr.grad) < 1e-5
    ops.mse_to_target(x.cuda(), 0., 2., loss)   # accumulates
    assert relative_error(loss, (expected + 2 * torch.mean(xr ** 2.)).reshape(1)) < 1e-5

    fake, real = torch.randn(3, 7, 500), torch.randn(3, 7, 500)
    fr = fake.double().requires_grad_()
    expected = 2.5 * torch.mean(torch.abs(real.double() - fr))
    expected.backward()
    loss.zero_()
    gfake = torch.ones(fake.shape, device='cuda')
    ops.l1_mean(fake.cuda(), real.cuda(), 2.5, loss, gfake, accumulate=True)
    assert relative_error(loss, expected.reshape(1)) < 1e-5
    assert relative_error(gfake, fr.grad + 1) < 1e-6


def test_adamw_matches_torch(ops):
    torch.manual_seed(6)
    param = torch.nn.Parameter(torch.randn(5000))
    # config/defaults.py:390-394
    optimizer = torch.optim.AdamW([param], lr=2e-4, betas=(.8, .99), eps=1e-9)
    p = param.detach().clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(5000)
        param.grad = grad.clone()
        optimizer.step()
        ops.adamw(p, (grad * 4).cuda(), m, v, 2e-4, (.8, .99), 1e-9, .01, step, grad_scale=.25)
        assert relative_error(p, param) < 1e-6


def test_axpby_row_sum(ops):
    torch.manual_seed(7)
    x, y = torch.randn(1000), torch.randn(1000)
    out = ops.axpby(2., x.cuda(), .5, y.cuda())
    assert relative_error(out, 2 * x + .5 * y) < 1e-6
    out = ops.axpby(2., x.cuda(), 0., torch.full((1000,), float('nan'), device='cuda'))
    assert relative_error(out, 2 * x) < 1e-6
    m = torch.randn(37, 300)
    assert relative_error(ops.row_sum(m.cuda(), torch.empty(37, device='cuda'), 37, 300), m.sum(1)) < 1e-5


def test_feature_assembly_backward_pieces(ops):
    from promonet_b200.model import init
    torch.manual_seed(8)
    edges = init.pitch_distribution()
    pitch = 40. + 600. * torch.rand(2, 50)
    bins = ops.pitch_bins(pitch.cuda(), edges.cuda(), 50., 550.)
    expected = torch.clip(torch.searchsorted(edges, torch.clip(pitch, 50., 550.)), 0, 255)
    assert torch.equal(bins.cpu(), expected)
    gout = torch.randn(2, 113, 50)
    gtable = torch.zeros(256, 64, device='cuda')
    ops.embedding_backward(gout.cuda(), bins, gtable, channel_offset=40)
    table = torch.zeros(256, 64, dtype=torch.double, requires_grad=True)
    (table[expected].permute(0, 2, 1) * gout[:, 40:104].double()).sum().backward()
    assert relative_error(gtable, table.grad) < 1e-5

    embedding = torch.randn(109, 256)
    speakers = torch.tensor([3, 108])
    sbr, lr = torch.rand(2), torch.rand(2)
    g = ops.global_features(embedding.cuda(), speakers.cuda(), sbr.cuda(), lr.cuda())
    assert torch.equal(g.cpu(), torch.cat([embedding[speakers], sbr[:, None], lr[:, None]], 1))


@pytest.mark.parametrize('window,eps,layout', [('hann', 1e-6, 0), ('rect', 0., 1)])
def test_stft_magnitude_and_backward(ops, window, eps, layout):
    """preprocess/spectrogram.py:36-52 (hann) and discriminator.py:175-195 (no window)"""
    torch.manual_seed(9)
    audio = .3 * torch.randn(2, 4096)
    ar = audio.double().requires_grad_()
    padded = F.pad(ar[:, None], (384, 384), mode='reflect')[:, 0]
    spec = torch.stft(
        padded, 1024, 256, 1024, torch.hann_window(1024, dtype=torch.double) if window == 'hann' else None,
        center=False, return_complex=True)
    real = torch.view_as_real(spec)
    mag = torch.sqrt(real.pow(2).sum(-1) + eps)
    if layout == 1:
        mag = mag.permute(0, 2, 1)
    g = torch.randn(mag.shape)
    (mag * g.double()).sum().backward()
    magnitude, spectrum = ops.stft_magnitude(audio.cuda(), window, eps, layout)
    assert relative_error(magnitude, mag) < TOLERANCE
    gaudio = ops.stft_magnitude_backward(
        g.cuda().contiguous(), spectrum, torch.empty(audio.shape, device='cuda'), window, eps, layout)
    assert relative_error(gaudio, ar.grad) < TOLERANCE


def test_mel_loss(ops):
    """train/core.py:277-305 against the oracle's mel basis"""
    from oracle import dsp
    torch.manual_seed(10)
    basis = torch.from_numpy(dsp.mel_basis(22050, 1024, 80)).double()
    magnitude = torch.rand(2, 513, 20) + .05
    target = torch.log(basis.float() @ (torch.rand(2, 513, 20) + .05))
    mr = magnitude.double().requires_grad_()
    expected = 45. * F.l1_loss(target.double(), torch.log(basis @ mr))
    expected.backward()
    loss = torch.zeros(1, device='cuda')
    gmag = torch.empty(magnitude.shape, device='cuda')
    ops.mel_loss(magnitude.cuda(), target.cuda(), 45., loss, gmag)
    assert relative_error(loss, expected.reshape(1)) < 1e-5
    assert relative_error(gmag, mr.grad) < 1e-4
    mels = ops.linear_to_mel(magnitude.cuda())
    assert relative_error(mels, torch.log(basis @ magnitude.double())) < 1e-5


@pytest.mark.parametrize('n_fft', [80, 320, 2560])
def test_spectral_convergence_pieces(ops, n_fft):
    """promonet/train/loss.py:61-121 for one resolution: the STFT as a 1 x 1 convolution over
    frames read in place, the loss ratio and its gradient back to the signal"""
    from oracle import train as oracle_train
    torch.manual_seed(15)
    batch, samples, hop = 2, 4096, n_fft // 4
    y = .3 * torch.randn(batch, 1, samples)
    x = (y + .1 * torch.randn(batch, 1, samples)).double().requires_grad_()
    window = torch.hann_window(n_fft, dtype=torch.double)
    magnitudes = [
        torch.sqrt(torch.clamp(torch.abs(torch.stft(
            s.squeeze(1), n_fft, hop, n_fft, window, return_complex=True)), min=1e-7))
        for s in (x, y.double())]
    expected = torch.norm(magnitudes[1] - magnitudes[0], p=1) / torch.norm(magnitudes[1], p=1)
    expected.backward()

    basis = ops.dft_basis(n_fft, 'cuda')
    rows, frames = basis.shape[0], 1 + samples // hop
    k = torch.arange(n_fft // 2 + 1, dtype=torch.double)[:, None] * torch.arange(n_fft, dtype=torch.double)
    reference = torch.cat([window * torch.cos(2 * torch.pi * k / n_fft), -window * torch.sin(2 * torch.pi * k / n_fft)])
    assert relative_error(basis, reference) < 1e-6
    both = torch.cat([y, x.detach().float()]).cuda().view(2 * batch, samples)
    padded = ops.reflect_pad(both, n_fft // 2, n_fft // 2)
    geom = ops.geometry(2 * batch, n_fft, rows, (frames, 1), (1, 1), strides=(1, hop, samples + n_fft))
    spec = ops.conv_gemm(geom, False, padded, basis, torch.empty(2 * batch, rows, frames, device='cuda'))
    stft = torch.stft(torch.cat([y, x.detach().float()]).squeeze(1).double(), n_fft, hop, n_fft, window,
                      return_complex=True)
    assert relative_error(spec, torch.cat([stft.real, stft.imag], 1)) < TOLERANCE
    sums, loss = torch.zeros(2, device='cuda'), torch.zeros(1, device='cuda')
    gspec = torch.empty(batch, rows, frames, device='cuda')
    ops.spectral_convergence(spec, batch, 1., sums, loss, gspec)
    assert relative_error(loss, expected.detach().reshape(1)) < TOLERANCE
    basis_t = ops.transpose_weight(basis, torch.empty_like(basis), rows, n_fft, 1)
    gframes = ops.conv_gemm(
        ops.geometry(batch, rows, n_fft, (frames, 1), (1, 1)), False, gspec, basis_t,
        torch.empty(batch, n_fft, frames, device='cuda'))
    gpadded = ops.frame_overlap_add(gframes, torch.zeros(batch, samples + n_fft, device='cuda'), hop)
    gx = ops.reflect_pad_backward(
        gpadded, torch.empty(batch, samples, device='cuda'), n_fft // 2, n_fft // 2)
    assert relative_error(gx, x.grad.squeeze(1)) < 5e-4


def test_grouped_weight_preparation(ops):
    """Grouped Conv1d (discriminator.py:218-224) as a block-diagonal dense convolution"""
    torch.manual_seed(16)
    groups, c_in, c_out, k = 4, 16, 64, 5
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(c_in, c_out, k, 2, groups=groups, padding=2))
    v, g = conv.weight_v.detach(), conv.weight_g.detach()
    x = torch.randn(3, c_in, 50)
    expected = conv(x).detach()
    dense = torch.empty(c_out, c_in, k, device='cuda')
    folded = torch.empty(c_out, c_in // groups, k, device='cuda')
    packed = torch.empty(ops.packed_floats(c_out, c_in, k), device='cuda')
    table = ops.weight_table([{
        'v': v.cuda().contiguous(), 'g': g.cuda().contiguous(), 'w': folded, 'packed': packed,
        'dense': dense, 'dim0': c_out, 'dim1': c_in, 'taps': k, 'groups': groups}], 'cuda')
    keep = table  # the descriptors point into these tensors
    ops.prepare_weights(table, 1, c_out)
    reference = torch.zeros(c_out, c_in, k)
    per_in, per_out = c_in // groups, c_out // groups
    for group in range(groups):
        reference[group * per_out:(group + 1) * per_out, group * per_in:(group + 1) * per_in] = \
            conv.weight.detach()[group * per_out:(group + 1) * per_out]
    assert relative_error(dense, reference) < 1e-6
    geom = ops.geometry(3, c_in, c_out, (50, 1), (k, 1), (2, 1), 1, (2, 0))
    out = torch.empty(3, c_out, geom.h_out, 1, device='cuda')
    ops.conv_gemm_tc(geom, False, x.cuda().view(3, c_in, 50, 1), packed, out, bias=conv.bias.detach().cuda())
    assert relative_error(out.view(expected.shape), expected) < TF32_TOLERANCE
    extracted = ops.extract_grouped(
        reference.cuda(), torch.empty(c_out, per_in, k, device='cuda'), c_out, c_in, k, groups)
    assert torch.equal(extracted.cpu(), conv.weight.detach())
