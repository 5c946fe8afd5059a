"""The C-ABI library builds, loads, and exports every symbol the header declares
(no compute: runs without a GPU)"""
import ctypes
import re

from conftest import ROOT


def declared_symbols():
    header = (ROOT / 'include' / 'promonet_b200.h').read_text()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    return sorted(set(re.findall(r'\b(pmn_[a-z0-9_]+)\s*\(', header)))


def test_header_declares_symbols():
    symbols = declared_symbols()
    assert 'pmn_generator_forward' in symbols
    assert 'pmn_conv1d' in symbols


def test_library_exports_every_declared_symbol():
    from promonet_b200 import build
    library = ctypes.CDLL(str(build.build()))
    missing = [s for s in declared_symbols() if not hasattr(library, s)]
    assert not missing, missing


def test_binding_covers_header():
    from promonet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.library().pmn_version() >= 1


def test_errors_are_reported_without_a_device():
    from promonet_b200 import _lib
    lib = _lib.library()
    status = lib.pmn_conv1d(
        None, None, None, None, None, None, None, 0, 1., 1, 1, 1, 1, 1, 1, 1, 0,
        1., 0, None)
    assert status == -1
    assert b'null' in lib.pmn_last_error()


def test_product_paths_refuse_to_run_without_a_gpu():
    """No CPU fallback anywhere: without CUDA the entry points raise instead of computing"""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import promonet_b200
    from promonet_b200.train import Trainer
    with pytest.raises(RuntimeError, match='CUDA'):
        Trainer()
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.model.Generator()
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.edit.grid.sample(torch.rand(40, 10), torch.linspace(0., 9., 5))
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.preprocess.from_audio(torch.zeros(1, 4096))
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.synthesize.from_features(
            torch.zeros(8, 4), torch.zeros(1, 4), torch.zeros(1, 4), torch.zeros(1, 40, 4))
