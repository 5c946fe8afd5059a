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
