"""Validation metrics: drop-in for promonet.evaluate.Metrics
(promonet/evaluate/metrics.py:17-83) for the prosody and pronunciation metrics
(loudness / loudness-loud / loudness-quiet RMSE, periodicity RMSE, voiced pitch
error in cents, PPG Jensen-Shannon distance).

The reference's update() is ~40 small kernels and four host synchronisations per
call (boolean-mask indexing, metrics.py:199-204,254-261); here update() is ONE
launch (`pmn_metrics_update`) that adds every running sum into 12 doubles on the
device, and the only device-to-host copy is in __call__.  update() also takes
batches (leading item axis): the sums are over frames, so a batch of utterances
is the same as consecutive calls.

Word error rate (metrics.py:315-319) needs whisper transcripts and jiwer: out of
scope, `predicted_text` / `target_text` raise.
"""
import math

import torch

from promonet_b200 import _lib, config

NAMES = ('loudness', 'loudness-loud', 'loudness-quiet', 'periodicity', 'pitch', 'ppg')


class Metrics:

    def __init__(self, device=None, loudness_threshold=-60., similarity=None, sums=None):
        """loudness_threshold: metrics.py:172; similarity: optional (40, 40) phoneme
        similarity matrix, already raised to ppgs.SIMILARITY_EXPONENT (a data asset
        of the un-vendored ppgs package; without it the distance is the plain
        Jensen-Shannon distance of the sparsified PPGs); sums: optional device buffer of
        PMN_METRICS_SLOTS doubles to accumulate into (a row of a table shared by several
        Metrics, so that data-parallel ranks combine all of them with one all-reduce)"""
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200.evaluate.Metrics needs a CUDA device; there is no CPU path')
        self.device = torch.device(
            'cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.loudness_threshold = float(loudness_threshold)
        self.similarity = None if similarity is None else similarity.to(
            self.device, torch.float32).contiguous()
        if self.similarity is not None and self.similarity.shape != (
                config.PPG_CHANNELS, config.PPG_CHANNELS):
            raise ValueError('similarity must be (40, 40)')
        if sums is None:
            sums = torch.zeros(_lib.METRICS_SLOTS, dtype=torch.float64, device=self.device)
        if sums.shape != (_lib.METRICS_SLOTS,) or sums.dtype != torch.float64 or \
                not sums.is_contiguous() or sums.device != self.device:
            raise ValueError(f'sums must be {_lib.METRICS_SLOTS} contiguous doubles on {self.device}')
        self.sums = sums

    def reset(self):
        self.sums.zero_()

    def update(
        self,
        predicted_loudness,
        predicted_pitch,
        predicted_periodicity,
        predicted_ppg,
        target_loudness,
        target_pitch,
        target_periodicity,
        target_ppg,
        predicted_text=None,
        target_text=None
    ):
        """Arguments as metrics.py:38-53: loudness (bands, F) or (B, bands, F) — the two sides
        may have different row counts (8 bands vs 513 bins), each is averaged over its own;
        pitch, periodicity (1, F) or (B, F); ppg (1, 40, F) or (B, 40, F), or None for both."""
        if predicted_text is not None or target_text is not None:
            raise NotImplementedError('word error rate is outside the accelerated path')
        f32 = lambda t: None if t is None else t.to(self.device, torch.float32).contiguous()
        predicted_loudness, target_loudness = f32(predicted_loudness), f32(target_loudness)
        frames = predicted_pitch.shape[-1]
        items = predicted_pitch.numel() // max(frames, 1)
        bands = []
        for loudness in (predicted_loudness, target_loudness):
            if loudness.shape[-1] != frames or loudness.numel() % max(items * frames, 1):
                raise ValueError('loudness and pitch disagree on the number of frames')
            bands.append(loudness.numel() // max(items * frames, 1))
        contours = [f32(t) for t in (
            predicted_pitch, target_pitch, predicted_periodicity, target_periodicity)]
        if any(t.numel() != items * frames for t in contours):
            raise ValueError('pitch and periodicity must have the same shape')
        if (predicted_ppg is None) != (target_ppg is None):
            raise ValueError('give both PPGs or neither')
        predicted_ppg, target_ppg = f32(predicted_ppg), f32(target_ppg)
        if predicted_ppg is not None and not (
                predicted_ppg.shape == target_ppg.shape and
                predicted_ppg.shape[-2:] == (config.PPG_CHANNELS, frames) and
                predicted_ppg.numel() == items * config.PPG_CHANNELS * frames):
            raise ValueError('ppg must be (B, 40, F)')
        if items == 0 or frames == 0:
            return
        with torch.cuda.device(self.device):
            _lib.check(_lib.library().pmn_metrics_update(
                _lib.ptr(predicted_loudness), bands[0], _lib.ptr(target_loudness), bands[1],
                *[_lib.ptr(t) for t in contours],
                _lib.ptr(predicted_ppg), _lib.ptr(target_ppg), config.PPG_CHANNELS,
                _lib.ptr(self.similarity), items, frames,
                self.loudness_threshold, config.VOICING_THRESHOLD, config.SPARSE_PPG_THRESHOLD,
                _lib.ptr(self.sums), _lib.stream()))

    def __call__(self):
        """-> {'pitch', 'periodicity', 'ppg' (when PPGs were given), 'loudness', 'loudness-loud',
        'loudness-quiet'} (metrics.py:26-36,180-183); a metric that saw no frame is nan"""
        return finish(self.sums.cpu().tolist())       # the one device-to-host copy


def finish(sums):
    """The scalars of metrics.py:26-36 from the PMN_METRICS_SLOTS running sums
    (include/promonet_b200.h, pmn_metrics_update): RMSE = sqrt(squares / count), pitch in cents =
    1200 x mean |log2 ratio| (metrics.py:221-228), PPG distance = total / frames"""
    mean = lambda i: sums[i] / sums[i + 1] if sums[i + 1] else math.nan
    result = {'pitch': 1200. * mean(8), 'periodicity': math.sqrt(mean(6))}
    if sums[11]:
        result['ppg'] = mean(10)
    result.update({
        'loudness': math.sqrt(mean(0)),
        'loudness-loud': math.sqrt(mean(2)),
        'loudness-quiet': math.sqrt(mean(4))})
    return result
