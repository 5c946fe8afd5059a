"""CPU restatement of the in-training validation arithmetic (TEST INFRASTRUCTURE ONLY)

* ``Metrics`` follows promonet/evaluate/metrics.py:17-312 line by line.  The
  parts written in promonet itself (band means and the loud/quiet split, the
  voicing mask and the 1200 x log2 pitch error, the sparsify-then-distance PPG
  metric, the result dictionary) are **pinned**: ``oracle/make_golden.py
  --metrics`` runs the reference's own ``promonet.evaluate.Metrics`` (imported
  unmodified through ``oracle.ref_shim``) on seeded inputs -> ``tests/golden/metrics.npz``.
* The third-party primitives under them are absent from /root/reference and
  **parity unpinned**: ``torchutil.metrics.{RMSE, L1, Average}`` (running sum and
  count), ``penn.voicing.threshold`` (periodicity > threshold) and
  ``ppgs.distance`` (Jensen-Shannon distance per frame, optionally through the
  phoneme-similarity matrix, which is a data asset of ppgs we do not have).
  They are restated below from their published definitions, and the shim hands
  the same restatements to the reference classes.
* ``edit_from_features`` follows promonet/edit/core.py:17-132 on top of
  ``oracle.features.grid_sample`` (pinned, grid.npz).
"""
import math

import torch

from oracle import features

VOICING_THRESHOLD = .1625          # promonet/config/defaults.py:135
SPARSE_PPG_METHOD = 'percentile'   # :113
SPARSE_PPG_THRESHOLD = .85         # :117
FMIN, FMAX = 50., 550.             # :27-28


###############################################################################
# torchutil.metrics / penn.voicing / ppgs.distance (un-vendored: restated)
###############################################################################


class Average:
    """torchutil.metrics.Average: running total / count"""

    def __init__(self):
        self.reset()

    def __call__(self):
        return float(self.total / self.count) if self.count else float('nan')

    def update(self, values, count):
        self.total += float(values)
        self.count += int(count)

    def reset(self):
        self.total, self.count = 0., 0


class L1(Average):
    """torchutil.metrics.L1: mean absolute error over every element seen"""

    def update(self, predicted, target):
        self.total += float((predicted.double() - target.double()).abs().sum())
        self.count += predicted.numel()


class RMSE(Average):
    """torchutil.metrics.RMSE: sqrt(sum of squared errors / elements)"""

    def __call__(self):
        return math.sqrt(self.total / self.count) if self.count else float('nan')

    def update(self, predicted, target):
        self.total += float(((predicted.double() - target.double()) ** 2).sum())
        self.count += predicted.numel()


def voicing_threshold(periodicity, threshold):
    """penn.voicing.threshold"""
    return periodicity > threshold


def ppg_distance(ppgX, ppgY, reduction='mean', normalize=False, exponent=None, similarity=None):
    """ppgs.distance on (channels, frames) PPGs: Jensen-Shannon distance per frame.
    `similarity` (channels, channels), already raised to the exponent, replaces the
    asset ppgs loads when normalize=True (absent here, so the default is no transform)."""
    ppgX = torch.clamp(ppgX, 1e-8, 1 - 1e-8)
    ppgY = torch.clamp(ppgY, 1e-8, 1 - 1e-8)
    if similarity is not None:
        ppgX = similarity.T @ ppgX
        ppgY = similarity.T @ ppgY
    log_average = torch.log((ppgX + ppgY) / 2)
    kl_X = (ppgX * (torch.log(ppgX) - log_average)).sum(dim=0)
    kl_Y = (ppgY * (torch.log(ppgY) - log_average)).sum(dim=0)
    distance = torch.sqrt(((kl_X + kl_Y) / 2).clamp_min(0.))
    if reduction == 'mean':
        return distance.mean()
    if reduction == 'sum':
        return distance.sum()
    return distance


###############################################################################
# promonet.evaluate.Metrics (evaluate/metrics.py:17-312)
###############################################################################


class Loudness:
    """metrics.py:169-209"""

    def __init__(self, threshold=-60.):
        self.threshold = threshold
        self.loud, self.quiet, self.both = RMSE(), RMSE(), RMSE()

    def __call__(self):
        return {
            'loudness': self.both(), 'loudness-loud': self.loud(), 'loudness-quiet': self.quiet()}

    def update(self, predicted, target):
        if predicted.ndim == 3:
            predicted = predicted.squeeze(0)
        if target.ndim == 3:
            target = target.squeeze(0)
        predicted = predicted.mean(dim=-2, keepdim=True)
        target = target.mean(dim=-2, keepdim=True)
        loud = torch.logical_and(predicted > self.threshold, target > self.threshold)
        self.loud.update(predicted[loud], target[loud])
        self.quiet.update(predicted[~loud], target[~loud])
        self.both.update(predicted, target)

    def reset(self):
        for metric in (self.loud, self.quiet, self.both):
            metric.reset()


class Pitch(L1):
    """metrics.py:212-261: mean voiced error in cents"""

    def __call__(self):
        return 1200 * super().__call__()

    def update(self, predicted_pitch, predicted_periodicity, target_pitch, target_periodicity):
        voicing = (
            voicing_threshold(predicted_periodicity, VOICING_THRESHOLD) &
            voicing_threshold(target_periodicity, VOICING_THRESHOLD))
        super().update(torch.log2(predicted_pitch[voicing]), torch.log2(target_pitch[voicing]))


class PPG(Average):
    """metrics.py:269-312 with ppgs.REPRESENTATION_KIND == 'ppg'"""

    def __init__(self, similarity=None):
        super().__init__()
        self.similarity = similarity

    def update(self, predicted, target):
        predicted = features.sparsify(predicted, SPARSE_PPG_METHOD, SPARSE_PPG_THRESHOLD)
        target = features.sparsify(target, SPARSE_PPG_METHOD, SPARSE_PPG_THRESHOLD)
        total = ppg_distance(
            predicted.squeeze(0), target.squeeze(0), reduction='sum', similarity=self.similarity)
        super().update(total, predicted.shape[-1])


class Metrics:
    """metrics.py:17-83 (the WER branch needs whisper transcripts: out of scope)"""

    def __init__(self, similarity=None):
        self.loudness = Loudness()
        self.periodicity = RMSE()
        self.pitch = Pitch()
        self.ppg = PPG(similarity)

    def __call__(self):
        result = {'pitch': self.pitch(), 'periodicity': self.periodicity()}
        if self.ppg.count:
            result['ppg'] = self.ppg()
        return result | self.loudness()

    def update(self, predicted_loudness, predicted_pitch, predicted_periodicity, predicted_ppg,
               target_loudness, target_pitch, target_periodicity, target_ppg):
        self.loudness.update(predicted_loudness, target_loudness)
        self.periodicity.update(predicted_periodicity, target_periodicity)
        self.pitch.update(predicted_pitch, predicted_periodicity, target_pitch, target_periodicity)
        if predicted_ppg is not None and target_ppg is not None:
            self.ppg.update(predicted_ppg, target_ppg)

    def reset(self):
        for metric in (self.loudness, self.periodicity, self.pitch, self.ppg):
            metric.reset()


###############################################################################
# promonet.edit.from_features (edit/core.py:17-132)
###############################################################################


def grid_constant(frames, ratio):
    """promonet.edit.grid.constant -> ppgs.edit.grid.constant (un-vendored, [RECALLED]): a uniform
    grid of round(T / ratio + 1e-4) positions over [0, T - 1], the frame count of the reference's
    own selective-stretch branch (promonet/edit/core.py:82)"""
    return features.grid_of_length(frames, round(frames / ratio + 1e-4))


def edit_from_features(loudness, pitch, periodicity, ppg, pitch_shift_cents=None,
                       time_stretch_ratio=None, loudness_scale_db=None):
    """edit/core.py:49-132 with stretch_unvoiced = stretch_silence = True"""
    if time_stretch_ratio is not None:
        grid = grid_constant(ppg.shape[-1], time_stretch_ratio)                # :54-57
        pitch = 2 ** features.grid_sample(torch.log2(pitch), grid)             # :113
        periodicity = features.grid_sample(periodicity, grid)                  # :114
        loudness = features.grid_sample(loudness, grid)                        # :115
        ppg = features.grid_sample(ppg, grid, 'linear')                        # :116
    if pitch_shift_cents is not None:                                          # :121-124
        pitch = torch.clip(pitch * 2 ** (pitch_shift_cents / 1200), FMIN, FMAX)
    if loudness_scale_db is not None:                                          # :127-128
        loudness = loudness + loudness_scale_db
    return loudness, pitch, periodicity, ppg
