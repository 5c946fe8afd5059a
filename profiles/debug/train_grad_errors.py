"""Gradient error of the GPU training step and of torch fp32 autograd, both against
fp64 autograd on the oracle (which tensors are numerically touchy, and how touchy)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch
from oracle import train as ot
from promonet_b200.model import init
from promonet_b200.train.core import Trainer

states = init.hifigan_state(1234), init.discriminator_state(1234)
batch = ot.batch(2, 8, seed=21)
res = {}
for dt in (torch.float32, torch.float64):
    g, d = ot.leaf_state(states[0], dt), ot.leaf_state(states[1], dt)
    b = [t.to(dt) if t.is_floating_point() else t for t in batch]
    res[dt] = ot.step(g, d, b)
trainer = Trainer(*states)
trainer.step(*[t.cuda().contiguous() for t in batch], update=False)
ours = (trainer.generator.params.gradients(), trainer.discriminators.params.gradients())
rows = []
for kind in (0, 1):
    for k, ref in res[torch.float64][kind + 1].items():
        scale = ref.abs().max()
        e_ours = float((ours[kind][k].double().cpu() - ref).abs().max() / scale)
        e_32 = float((res[torch.float32][kind + 1][k].double() - ref).abs().max() / scale)
        rows.append((e_ours, e_32, float(scale), k))
rows.sort(reverse=True)
for r in rows[:25]:
    print('ours %.2e  torch32 %.2e  scale %.2e  %s' % r)
