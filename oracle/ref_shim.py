"""Import the UNMODIFIED reference package from /root/reference behind stubs.

Container-only: /root/reference does not exist on the GPU box, so nothing in
the `-m gpu` tests, smoke() or bench.py calls this.  It is used by
``oracle/make_golden.py`` (to freeze golden vectors) and by the CPU tests that
pin the restatements against the reference when the reference tree is present.

The reference cannot be imported as-is: ``promonet/__init__.py:7-35`` imports
yapecs, GPUtil, ppgs, penn, librosa, torchutil, matplotlib, pypar, pyworld,
resampy, soundfile, jiwer, umap, whisper -- none installed, no network.  The
stubs below provide only import-time surface, except for the few third-party
primitives the reference's own arithmetic is built on, which forward to our
restatements: ``ppgs.sparsify`` (``oracle.features``), and for
``promonet.evaluate.Metrics`` ``torchutil.metrics.{L1, Average, RMSE}``,
``penn.voicing.threshold`` and ``ppgs.distance`` (``oracle.metrics``).
"""
import argparse
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('PROMONET_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'promonet'))


class _Any(types.ModuleType):
    """Attribute-auto-stub module"""

    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        sub = _Any(self.__name__ + '.' + key)
        setattr(self, key, sub)
        return sub

    def __call__(self, *args, **kwargs):
        return self


def _module(name, **kwargs):
    module = types.ModuleType(name)
    module.__dict__.update(kwargs)
    module.__path__ = []
    sys.modules[name] = module
    return module


def load(config_file=None):
    """Returns the reference `promonet` module (one config per process)"""
    if 'promonet' in sys.modules and hasattr(sys.modules['promonet'], 'model'):
        return sys.modules['promonet']
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')

    from oracle import features, metrics

    def configure(name, defaults):
        # yapecs.configure restatement (promonet/__init__.py:10-11)
        if config_file:
            namespace = {}
            with open(config_file) as file:
                exec(file.read(), namespace)
            if namespace.get('MODULE') == name:
                for key, value in namespace.items():
                    if key.isupper() and key != 'MODULE':
                        setattr(defaults, key, value)

    _module(
        'yapecs',
        configure=configure,
        ArgumentParser=argparse.ArgumentParser)
    _module('GPUtil', getGPUs=lambda: [])
    grids = types.SimpleNamespace(
        constant=lambda tensor, ratio: metrics.grid_constant(tensor.shape[-1], ratio),
        of_length=lambda tensor, length: features.grid_of_length(tensor.shape[-1], length))
    _module(
        'ppgs',
        edit=types.SimpleNamespace(grid=grids),
        REPRESENTATION_KIND='ppg',
        sparsify=features.sparsify,
        distance=metrics.ppg_distance,
        PHONEMES=[str(i) for i in range(40)],
        SIMILARITY_EXPONENT=1.2,
        representation_file_extension=lambda: '-ppg.pt')
    for name in [
        'transformers', 'penn', 'librosa', 'pypar', 'pyworld', 'resampy',
        'soundfile', 'jiwer', 'umap', 'whisper', 'whisper.normalizers',
        'torbi', 'matplotlib', 'matplotlib.pyplot', 'huggingface_hub',
    ]:
        if name == 'huggingface_hub':
            try:
                import huggingface_hub  # noqa: F401
                continue
            except Exception:
                pass
        sys.modules[name] = _Any(name)
    sys.modules['umap'].UMAP = object
    sys.modules['whisper.normalizers'].EnglishTextNormalizer = object
    torchutil = _Any('torchutil')
    torchutil.notify = lambda name: (lambda function: function)
    torchutil.metrics = types.ModuleType('torchutil.metrics')
    for cls in ('L1', 'Average', 'RMSE'):
        setattr(torchutil.metrics, cls, getattr(metrics, cls))
    sys.modules['torchutil'] = torchutil
    sys.modules['penn'].voicing.threshold = metrics.voicing_threshold

    argv = sys.argv
    sys.argv = argv[:1]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import promonet
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.argv = argv

    # librosa.filters.mel is used by preprocess/spectrogram.py:118
    from oracle import dsp
    sys.modules['librosa'].filters.mel = \
        lambda sr, n_fft, n_mels: dsp.mel_basis(sr, n_fft, n_mels)
    return promonet
