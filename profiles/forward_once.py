"""Two generator forwards (B = 32 x 5 s, tensor-core math) for ncu:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none -s <launches of forward 1> --csv --log-file x.csv python profiles/forward_once.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import promonet_b200  # noqa: E402
from oracle import inputs  # noqa: E402

state = promonet_b200.model.init.hifigan_state(1234)
model = promonet_b200.model.Generator(state=state)
args = [t.cuda() for t in inputs.synthesis(32, 430, seed=1234)]
before = promonet_b200._lib.launch_count()
model(*args)
torch.cuda.synchronize()
print('launches per forward', promonet_b200._lib.launch_count() - before)
model(*args)
torch.cuda.synchronize()
