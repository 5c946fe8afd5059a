"""The reference's own modules as a timed baseline (bench.py only).

`load()` returns the unmodified `promonet` package -- from /root/reference in
the build container, from the git-ignored copy oracle/_ref/ (made by
oracle/vendor_reference.py) on the GPU box -- imported behind oracle.ref_shim's
stub modules, or None when neither tree exists (bench.py then falls back to the
oracle port and says so in `cpu_baseline.kind`)."""
import os
from pathlib import Path

from oracle import ref_shim

VENDORED = Path(__file__).resolve().parent / '_ref'


def root():
    for candidate in (os.environ.get('PROMONET_REFERENCE'), '/root/reference', str(VENDORED)):
        if candidate and os.path.isdir(os.path.join(candidate, 'promonet')):
            return candidate
    return None


def load(config_file=None):
    where = root()
    if where is None:
        return None
    ref_shim.REFERENCE_ROOT = where
    return ref_shim.load(config_file)


def generator(promonet, seed=1234):
    """promonet.model.Generator() as the reference constructs it
    (promonet/train/core.py:58), under the reference's seed, in eval mode"""
    import torch
    torch.manual_seed(seed)
    return promonet.model.Generator().eval()


def forward(model, loudness, pitch, periodicity, ppg, speakers, sbr, lr):
    """Generator.forward (promonet/model/generator.py:116-135), no autocast"""
    previous = model.default_previous_samples.to(loudness.device)
    return model(loudness, pitch, periodicity, ppg, speakers, sbr, lr, previous)
