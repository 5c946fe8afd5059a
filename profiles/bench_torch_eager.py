"""The bar on the same box: the reference's PyTorch modules run eagerly ON THE B200 (cuDNN).

    python profiles/bench_torch_eager.py [--batch 32] [--frames 430] [--steps 3]

SURVEY 8d asks for it next to the CPU baseline: `oracle/hifigan.py` is a functional restatement of
promonet/model/{generator,hifigan}.py (pinned to the reference, tests/golden), so running it with
its state dict on cuda:0 is what `promonet.model.Generator` does there: one cuDNN convolution, one
LeakyReLU, one add and one weight-norm fold per layer.  Three precision modes: strict fp32 (the
parity yardstick: torch.backends.cudnn.allow_tf32 = False), TF32 (PyTorch's default for
convolutions) and bf16 autocast; for the last two the error against strict fp32 is printed too
(the bar is 1e-4, BASELINE.json).  Not a product path: nothing here is imported by promonet_b200.
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import hifigan, inputs  # noqa: E402
from promonet_b200.model import init  # noqa: E402


def timed(function, steps):
    for _ in range(2):
        function()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(True), torch.cuda.Event(True)
    start.record()
    for _ in range(steps):
        result = function()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / steps, result


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--batch', type=int, default=32)
    parser.add_argument('--frames', type=int, default=430)
    parser.add_argument('--steps', type=int, default=3)
    args = parser.parse_args()
    device = torch.device('cuda', 0)
    state = {k: v.to(device) for k, v in init.hifigan_state(1234).items()}
    batch = [t.to(device) for t in inputs.synthesis(args.batch, args.frames)]
    samples = args.batch * args.frames * 256
    torch.backends.cudnn.benchmark = True

    def forward(autocast=False):
        with torch.inference_mode(), torch.autocast('cuda', torch.bfloat16, enabled=autocast):
            return hifigan.generator(state, *batch)

    result = {
        'metric': 'audio samples/sec synthesized (22.05 kHz), PyTorch eager on cuda:0',
        'unit': 'samples/s', 'batch': args.batch, 'frames': args.frames,
        'torch': torch.__version__, 'cudnn': torch.backends.cudnn.version(), 'modes': {}}
    exact = None
    for name, tf32, autocast in (('fp32', False, False), ('tf32', True, False), ('bf16_autocast', True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            ms, audio = timed(lambda: forward(autocast), args.steps)
        except Exception as error:      # a mode that does not run is reported, not fatal
            result['modes'][name] = {'error': repr(error)[:200]}
            continue
        audio = audio.float()
        if exact is None:
            exact = audio
        result['modes'][name] = {
            'ms_per_step': ms, 'value': samples / (ms * 1e-3),
            'error_vs_fp32': float((audio - exact).abs().max() / exact.abs().max())}
    print(json.dumps(result))


if __name__ == '__main__':
    main()
