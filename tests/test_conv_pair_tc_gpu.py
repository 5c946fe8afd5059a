"""The fused residual pair y = x + c2(lrelu(c1(lrelu(x)))) (conv_pair_tc.cu; Block.forward,
promonet/model/hifigan.py:198-210) through the C ABI: against an fp64 torch evaluation
(1e-4 relative, north_star) and against the two-launch tensor-core path it replaces (same
operand rounding, same products, same epilogue order).  Inside the generator, where the
two-launch path hands its bf16 hi/lo planes over untouched, the two are bit-identical
(last test); through the unit entry points the planes make a round trip through fp32
(hi + lo, split again), which may move a tie between hi and lo, so there the bar is 1e-5."""
import pytest
import torch

from conftest import relative_error
from test_conv1d_tc_gpu import conv1d_tc, reference

pytestmark = pytest.mark.gpu


def conv_pair(x, w1, b1, w2, b2, dilation, accum=None, accum_mode=0, accum_scale=1., want_out=True):
    from promonet_b200 import _lib
    lib = _lib.library()
    batch, channels, t_len = x.shape
    k = w1.shape[-1]
    tensors = [t.cuda().contiguous() if t is not None else None for t in (x, w1, b1, w2, b2)]
    out = torch.full_like(tensors[0], float('nan')) if want_out else None
    size = lib.pmn_conv_pair_tc_workspace_bytes(channels, k)
    workspace = torch.empty(size, dtype=torch.uint8, device='cuda')
    _lib.check(lib.pmn_conv_pair_tc(
        *[_lib.ptr(t) for t in tensors], _lib.ptr(out), _lib.ptr(accum), accum_mode, accum_scale,
        batch, channels, t_len, k, dilation, 0.1, workspace.data_ptr(), size, _lib.stream()))
    torch.cuda.synchronize()
    return out


def pair_inputs(channels, k, t_len, batch, seed):
    torch.manual_seed(seed)
    x = torch.randn(batch, channels, t_len)
    scale = (channels * k) ** -.5
    return (x, torch.randn(channels, channels, k) * scale, torch.randn(channels),
            torch.randn(channels, channels, k) * scale, torch.randn(channels))


def two_launches(x, w1, b1, w2, b2, dilation):
    """The path the kernel replaces: c1 writes lrelu'd operand planes, c2 adds the residual"""
    _, planes = conv1d_tc(x, w1, b1, dilation=dilation, in_slope=0.1, out_slope=0.1,
                          want_planes=True, want_out=False)
    out, _ = conv1d_tc(planes, w2, b2, residual=x, dilation=1, in_slope=1.)
    return out


def fp64(x, w1, b1, w2, b2, dilation):
    mid = reference(x, w1, b1, dilation, 0.1)
    return x.double() + reference(mid, w2, b2, 1, 0.1)


@pytest.mark.parametrize('channels,k,dilation,t_len,batch', [
    (32, 3, 1, 512, 1), (32, 11, 5, 1500, 2), (32, 7, 3, 247, 3), (32, 3, 5, 100, 1),
    (64, 7, 3, 700, 3), (64, 3, 1, 256, 1), (64, 11, 5, 118, 2), (64, 11, 1, 119, 1),
    (128, 11, 1, 300, 2), (128, 11, 5, 1111, 1), (128, 7, 1, 256, 1), (128, 3, 3, 5, 2),
    (128, 1, 1, 130, 1)])
def test_conv_pair_matches_fp64_and_the_two_launch_path(channels, k, dilation, t_len, batch):
    args = pair_inputs(channels, k, t_len, batch, channels + k + dilation + t_len)
    out = conv_pair(*args, dilation)
    assert bool(torch.isfinite(out).all())          # every output element was written
    assert relative_error(out, fp64(*args, dilation)) < 1e-4
    assert relative_error(out, two_launches(*args, dilation)) < 1e-5


@pytest.mark.parametrize('channels', [32, 64, 128])
def test_conv_pair_many_tiles_per_cta(channels):
    """More tiles than SMs x 2: ring phases, TMEM double buffering, the single mid buffer"""
    args = pair_inputs(channels, 3, 256 * 40 + 17, 8, channels)
    out = conv_pair(*args, 3)
    assert relative_error(out, two_launches(*args, 3)) < 1e-5
    again = conv_pair(*args, 3)
    assert torch.equal(out, again)


def test_conv_pair_accumulate_modes():
    """MRF mean (hifigan.py:141-145): accum = y / 3, then accum += y / 3, no fp32 output"""
    args = pair_inputs(64, 7, 600, 2, 5)
    y = fp64(*args, 5)
    accum = torch.full((2, 64, 600), float('nan'), device='cuda')
    assert conv_pair(*args, 5, accum=accum, accum_mode=1, accum_scale=1 / 3, want_out=False) is None
    assert relative_error(accum, y / 3) < 1e-4
    out = conv_pair(*args, 5, accum=accum, accum_mode=2, accum_scale=1 / 3)
    assert relative_error(accum, 2 * y / 3) < 1e-4
    assert relative_error(out, y) < 1e-4


def test_conv_pair_rejects_bad_arguments():
    from promonet_b200 import _lib
    args = pair_inputs(256, 3, 64, 1, 0)
    with pytest.raises(_lib.Error):
        conv_pair(*args, 1)                           # C = 256 does not fit one SM
    args = pair_inputs(64, 3, 64, 1, 0)
    x = args[0].cuda()
    lib = _lib.library()
    workspace = torch.empty(lib.pmn_conv_pair_tc_workspace_bytes(64, 3), dtype=torch.uint8, device='cuda')
    tensors = [t.cuda().contiguous() for t in args[1:]]
    status = lib.pmn_conv_pair_tc(
        x.data_ptr(), *[t.data_ptr() for t in tensors], x.data_ptr(), None, 0, 1.,
        1, 64, 64, 3, 1, 0.1, workspace.data_ptr(), workspace.numel(), _lib.stream())
    assert status != 0                                # in place is refused (halo rows)


@pytest.mark.parametrize('tag,kernel', [('c32k3', 3), ('c64k11', 11)])
def test_block_of_fused_pairs_matches_reference_golden(golden, tag, kernel):
    """Block.forward hifigan.py:198-210, outputs of the reference module itself"""
    from oracle import hifigan
    g = golden('block')
    cur = g[f'{tag}_x']
    for i, dilation in enumerate((1, 3, 5)):
        def folded(name):
            return hifigan.fold_weight_norm(
                g[f'{tag}_{name}.{i}.weight_g'], g[f'{tag}_{name}.{i}.weight_v'])
        cur = conv_pair(cur, folded('convs1'), g[f'{tag}_convs1.{i}.bias'],
                        folded('convs2'), g[f'{tag}_convs2.{i}.bias'], dilation).cpu()
    assert relative_error(cur, g[f'{tag}_y']) < 1e-4


def test_generator_output_does_not_depend_on_the_pair_mask():
    """Fused and two-launch residual blocks give the same result end to end: the same bits where the
    two launches run on conv1d_tc_kernel (stage 1, C = 128: the pair kernel repeats its arithmetic
    exactly), fp32 rounding apart where they run on conv1d_tcw_kernel (C = 64 and the long kernels
    of C = 32, whose tap sum has another order).  bf16 x 3 everywhere (f8=False): with the fp8 form
    at C = 128 the unfused blocks of that stage use other operands than the pair kernel, and the
    outputs then agree within the parity bar"""
    import promonet_b200
    from oracle import inputs
    state = promonet_b200.model.init.hifigan_state(promonet_b200.RANDOM_SEED)
    args = [a.cuda() for a in inputs.synthesis(3, 37, seed=9)]
    outputs = {
        mask: promonet_b200.model.Generator(state=state, pair_mask=mask, f8=False)(*args)
        for mask in (0, 0x008, 0x038, 0xFF8, 0x248, 0x1C0)}
    for mask in (0x008, 0x038):
        assert torch.equal(outputs[0], outputs[mask])
    for mask in (0xFF8, 0x248, 0x1C0):
        assert relative_error(outputs[mask], outputs[0]) < 1e-5
    with_f8 = {
        mask: promonet_b200.model.Generator(state=state, pair_mask=mask, f8=True)(*args)
        for mask in (0, 0x008, 0x038)}
    for mask, output in with_f8.items():
        assert relative_error(output, outputs[0]) < 5e-5, mask
