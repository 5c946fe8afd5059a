"""Numerics experiment (CPU, no GPU): the training convolutions round their operands to tf32
(tcgen05.mma.kind::tf32, half the bf16 rate, 4-byte operands in shared memory).  fp16 operands
(kind::f16) have the same 11 significant bits, run at twice the rate and halve the bytes the
gather stores and the MMA fetches - but only 5 exponent bits.  Does one training step keep its
gradients?  The whole Trainer.step runs on the CPU over tests/emulated_ops.py (the plain-torch
double of the kernel launches) with the operands of every convolution (forward, data gradient,
weight gradient) rounded as the tensor-core kernels would round them, and every parameter
gradient is compared with fp64 autograd of the oracle.

    python profiles/debug/training_operand_numerics.py

The reference itself trains under fp16 autocast + GradScaler (train/core.py:118,220,262).
"""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
import emulated_ops  # noqa: E402
from conftest import relative_error  # noqa: E402
from oracle import train as oracle_train  # noqa: E402
from promonet_b200.model import init  # noqa: E402


def tf32(value):
    bits = value.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def fp16(value):
    return value.half().float()


def bf16(value):
    return value.bfloat16().float()


def rounded_ops(rounding):
    """conv_gemm / conv_wgrad of the double with both operands rounded after their fused
    activation (where the kernels round them)"""
    import torch.nn.functional as F
    conv2d, conv2d_input, conv2d_weight = F.conv2d, torch.nn.grad.conv2d_input, torch.nn.grad.conv2d_weight

    class Patched:
        def __enter__(self):
            F.conv2d = lambda x, w, *a, **k: conv2d(rounding(x), rounding(w), *a, **k)
            torch.nn.grad.conv2d_input = lambda size, w, g, *a, **k: conv2d_input(
                size, rounding(w), rounding(g), *a, **k)
            torch.nn.grad.conv2d_weight = lambda x, size, g, *a, **k: conv2d_weight(
                rounding(x), size, rounding(g), *a, **k)

        def __exit__(self, *args):
            F.conv2d = conv2d
            torch.nn.grad.conv2d_input, torch.nn.grad.conv2d_weight = conv2d_input, conv2d_weight
    return Patched()


def main():
    from promonet_b200.train import Trainer
    patch = pytest.MonkeyPatch()
    emulated_ops.install(patch)
    states = init.hifigan_state(1234), init.discriminator_state(1234)
    batch = oracle_train.batch(2, 8, seed=21)
    g_state = oracle_train.leaf_state(states[0], torch.float64)
    d_state = oracle_train.leaf_state(states[1], torch.float64)
    _, g_grads, d_grads, _ = oracle_train.step(
        g_state, d_state, [t.double() if t.is_floating_point() else t for t in batch])
    print('relative error of the parameter gradients against fp64 autograd: median / 90 % / max, '
          'tensors above 3e-2')
    for name, rounding in (('fp32', lambda v: v), ('tf32', tf32), ('fp16', fp16), ('bf16', bf16)):
        trainer = Trainer(*states, device='cpu', math='fp32')
        with rounded_ops(rounding):
            trainer.step(*[t.contiguous() for t in batch], update=False)
        for kind, module, expected in (
                ('generator', trainer.generator, g_grads),
                ('discriminator', trainer.discriminators, d_grads)):
            gradients = module.params.gradients()
            errors = sorted(relative_error(gradients[n], g) for n, g in expected.items())
            print(f'{name:5s} {kind:13s} {errors[len(errors) // 2]:.1e} / '
                  f'{errors[len(errors) * 9 // 10]:.1e} / {errors[-1]:.1e}, '
                  f'{sum(e > 3e-2 for e in errors)} of {len(errors)}')
    patch.undo()


if __name__ == '__main__':
    main()
