"""Per-tensor difference between the tf32 (tcgen05) and fp32 (FMA) training steps"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from oracle import train as ot
from promonet_b200.model import init
from promonet_b200.train.core import Trainer

states = init.hifigan_state(1234), init.discriminator_state(1234)
batch = [t.cuda().contiguous() for t in ot.batch(2, 8, seed=21)]
grads = {}
for math in ('fp32', 'tf32'):
    trainer = Trainer(*states, math=math)
    losses = trainer.step(*batch, update=False)
    print(math, losses.tolist())
    grads[math] = (
        {k: v.clone() for k, v in trainer.generator.params.gradients().items()},
        {k: v.clone() for k, v in trainer.discriminators.params.gradients().items()},
        trainer.generated.clone())
print('generated', float((grads['tf32'][2] - grads['fp32'][2]).abs().max() / grads['fp32'][2].abs().max()))
for kind in (1, 0):
    for name, ref in grads['fp32'][kind].items():
        error = float((grads['tf32'][kind][name] - ref).abs().max() / ref.abs().max())
        flag = ' <<<' if error > 3e-2 else ''
        if kind == 0 or error > 1e-2:
            print(f'{error:.2e} {name}{flag}')
