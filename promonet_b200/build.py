"""Build libpromonet_b200.so in-tree with nvcc for sm_100a

The library is plain CUDA C++ behind an extern "C" surface
(include/promonet_b200.h); Python binds it with ctypes (promonet_b200/_lib.py).
"""
import hashlib
import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
SOURCES = sorted((ROOT / 'csrc').glob('*.cu'))
HEADERS = sorted((ROOT / 'csrc').glob('*.cuh')) + [
    ROOT.parent / 'include' / 'promonet_b200.h']
LIBRARY = ROOT / 'lib' / 'libpromonet_b200.so'
STAMP = ROOT / 'lib' / 'libpromonet_b200.stamp'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
    '-lcuda']


def nvcc():
    path = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(path):
        raise RuntimeError('nvcc not found; cannot build libpromonet_b200.so')
    return path


def digest():
    sha = hashlib.sha256()
    for file in SOURCES + HEADERS:
        sha.update(file.name.encode())
        sha.update(file.read_bytes())
    sha.update(' '.join(NVCC_FLAGS).encode())
    return sha.hexdigest()


def is_current():
    return (
        LIBRARY.exists() and
        STAMP.exists() and
        STAMP.read_text().strip() == digest())


def build(force=False, verbose=False):
    """Compile every kernel into one shared library; returns its path"""
    if not force and is_current():
        return LIBRARY
    LIBRARY.parent.mkdir(exist_ok=True)
    command = [nvcc(), *NVCC_FLAGS, '-o', str(LIBRARY)] + [str(s) for s in SOURCES]
    if verbose:
        command.insert(1, '-Xptxas=-v')
    result = subprocess.run(command, capture_output=True, text=True)
    if result.returncode:
        raise RuntimeError(
            'nvcc failed:\n' + ' '.join(command) + '\n' + result.stdout + result.stderr)
    if verbose:
        print(result.stderr)
    STAMP.write_text(digest())
    return LIBRARY


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
