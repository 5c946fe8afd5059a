"""Oracle: Viterbi decoding (test infrastructure; torbi is un-vendored: PARITY UNPINNED)

`decode` calls the C restatement (oracle/viterbi.c, built by oracle/Makefile);
`decode_numpy` is the same recurrence vectorised in numpy, used to cross-check the
C code on small cases.  Both follow torbi.from_probabilities as it is called at
promonet/preprocess/harmonics.py:270-276.
"""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
LIBRARY = ROOT / '_build' / 'libviterbi_oracle.so'
_library = None


def library():
    global _library
    if _library is None:
        if not LIBRARY.exists():
            subprocess.run(['make', '-C', str(ROOT), '-s'], check=True)
        _library = ctypes.CDLL(str(LIBRARY))
        _library.viterbi_oracle.restype = ctypes.c_int
    return _library


def _logs(observation, transition, initial, log_probs):
    observation = np.ascontiguousarray(observation, dtype=np.float32)
    states = observation.shape[-1]
    if transition is None:
        transition = np.full((states, states), 1. / states, dtype=np.float32)
    if initial is None:
        initial = np.full((states,), 1. / states, dtype=np.float32)
    transition = np.ascontiguousarray(transition, dtype=np.float32)
    initial = np.ascontiguousarray(initial, dtype=np.float32)
    if not log_probs:
        with np.errstate(divide='ignore'):
            observation = np.log(observation)
            transition = np.log(transition)
            initial = np.log(initial)
    return observation, transition, initial


def decode(observation, batch_frames=None, transition=None, initial=None, log_probs=False):
    """observation (B, T, S) -> indices (B, T) int32"""
    observation, transition, initial = _logs(observation, transition, initial, log_probs)
    batch, frames, states = observation.shape
    indices = np.zeros((batch, frames), dtype=np.int32)
    lengths = None
    if batch_frames is not None:
        lengths = np.ascontiguousarray(batch_frames, dtype=np.int32)
    # one call per utterance, spread over host threads (ctypes releases the GIL): only
    # the wall time of the checker depends on this
    function = library().viterbi_oracle

    def one(b):
        return function(
            observation[b:b + 1].ctypes.data_as(ctypes.c_void_p),
            lengths[b:b + 1].ctypes.data_as(ctypes.c_void_p) if lengths is not None else None,
            transition.ctypes.data_as(ctypes.c_void_p),
            initial.ctypes.data_as(ctypes.c_void_p),
            indices[b:b + 1].ctypes.data_as(ctypes.c_void_p),
            1, frames, states)

    if batch > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(batch, os.cpu_count() or 1)) as pool:
            statuses = list(pool.map(one, range(batch)))
    else:
        statuses = [one(b) for b in range(batch)]
    if any(statuses):
        raise MemoryError('viterbi_oracle')
    return indices


def decode_numpy(observation, batch_frames=None, transition=None, initial=None, log_probs=False):
    observation, transition, initial = _logs(observation, transition, initial, log_probs)
    batch, frames, states = observation.shape
    indices = np.zeros((batch, frames), dtype=np.int32)
    for b in range(batch):
        length = frames if batch_frames is None else int(batch_frames[b])
        delta = initial + observation[b, 0]
        psi = np.zeros((length, states), dtype=np.int32)
        for t in range(1, length):
            scores = delta[:, None] + transition        # (i, j), fp32
            psi[t] = scores.argmax(axis=0)              # first maximum = lowest index
            delta = scores.max(axis=0) + observation[b, t]
        state = int(delta.argmax())
        for t in range(length - 1, -1, -1):
            indices[b, t] = state
            if t > 0:
                state = int(psi[t, state])
    return indices
