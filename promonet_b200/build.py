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
    '-Xcompiler', '-fPIC']
OBJECTS = ROOT / 'lib' / 'obj'


def nvcc():
    path = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(path):
        raise RuntimeError('nvcc not found; cannot build libpromonet_b200.so')
    return path


def header_digest():
    sha = hashlib.sha256()
    for file in HEADERS:
        sha.update(file.name.encode())
        sha.update(file.read_bytes())
    sha.update(' '.join(NVCC_FLAGS).encode())
    return sha


def source_digest(source, headers=None):
    sha = (headers or header_digest()).copy()
    sha.update(source.name.encode())
    sha.update(source.read_bytes())
    return sha.hexdigest()


def digest():
    headers = header_digest()
    return hashlib.sha256(
        ''.join(source_digest(s, headers) for s in SOURCES).encode()).hexdigest()


def is_current():
    return (
        LIBRARY.exists() and
        STAMP.exists() and
        STAMP.read_text().strip() == digest())


def compile_one(source, verbose=False):
    """source.cu -> lib/obj/source.o unless an object of the same digest exists"""
    obj = OBJECTS / (source.stem + '.o')
    stamp = OBJECTS / (source.stem + '.stamp')
    wanted = source_digest(source)
    if obj.exists() and stamp.exists() and stamp.read_text().strip() == wanted:
        return obj, ''
    command = [nvcc(), *NVCC_FLAGS, '-c', '-o', str(obj), str(source)]
    if verbose:
        command.insert(1, '-Xptxas=-v')
    result = subprocess.run(command, capture_output=True, text=True)
    if result.returncode:
        raise RuntimeError(
            'nvcc failed:\n' + ' '.join(command) + '\n' + result.stdout + result.stderr)
    stamp.write_text(wanted)
    return obj, result.stderr


def build(force=False, verbose=False):
    """Compile every kernel (one object per source, in parallel) and link them
    into one shared library; returns its path"""
    if not force and is_current():
        return LIBRARY
    from concurrent.futures import ThreadPoolExecutor
    OBJECTS.mkdir(parents=True, exist_ok=True)
    if force:
        for stamp in OBJECTS.glob('*.stamp'):
            stamp.unlink()
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda s: compile_one(s, verbose), SOURCES))
    command = [
        nvcc(), '-shared', '-o', str(LIBRARY), *[str(obj) for obj, _ in results],
        '-lcuda']
    result = subprocess.run(command, capture_output=True, text=True)
    if result.returncode:
        raise RuntimeError(
            'link failed:\n' + ' '.join(command) + '\n' + result.stdout + result.stderr)
    if verbose:
        print('\n'.join(log for _, log in results))
    STAMP.write_text(digest())
    return LIBRARY


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
