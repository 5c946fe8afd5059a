"""Two eager training steps (BASELINE configs[3] shapes: 8 items x 16 384 samples) for ncu:
the first warms up (~790 launches), the second is the one to look at.
    ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 800 --csv \
        --log-file gpurun_out/train_launches.csv python profiles/train_step_once.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import train as oracle_train  # noqa: E402  (synthetic batch only)
from promonet_b200.model import init  # noqa: E402
from promonet_b200.train.core import Trainer  # noqa: E402

trainer = Trainer(init.hifigan_state(1234), init.discriminator_state(1234))
batch = [t.cuda().contiguous() for t in oracle_train.batch(8, 64, 1234)]
for _ in range(2):
    losses = trainer.step(*batch)
torch.cuda.synchronize()
print(losses.tolist())
