"""Pin oracle/metrics.py (validation metrics and feature edits): against the golden vectors
frozen from the reference's own promonet.evaluate.Metrics / promonet.edit.from_features, and
against the reference itself when /root/reference is present (CPU, no GPU)"""
import math

import numpy as np
import pytest
import torch

from oracle import make_golden, metrics, ref_shim

EDIT_NAMES = ('loudness', 'pitch', 'periodicity', 'ppg')


def same(actual, expected, rel=1e-9):
    return (math.isnan(actual) and math.isnan(expected)) or actual == pytest.approx(expected, rel=rel)


def case_inputs(index, frames, rows, target_rows):
    return (
        make_golden.metric_inputs(100 + index, frames, rows),
        make_golden.metric_inputs(200 + index, frames, target_rows))


def test_metrics_oracle_matches_reference_golden(golden):
    g = golden('metrics')
    names = [str(name) for name in g['names']]
    accumulated = metrics.Metrics()
    for index, (frames, rows, target_rows) in enumerate(g['cases'].tolist()):
        predicted, target = case_inputs(index, frames, rows, target_rows)
        single = metrics.Metrics()
        single.update(*predicted, *target)
        if index < 3:
            accumulated.update(*predicted, *target)
        for name, expected in zip(names, g[f'single_{index}'].tolist()):
            assert same(single()[name], expected), (index, name)
    for name, expected in zip(names, g['accumulated'].tolist()):
        assert same(accumulated()[name], expected), name


def edit_arguments(row):
    keys = ('pitch_shift_cents', 'time_stretch_ratio', 'loudness_scale_db')
    return {key: value for key, value in zip(keys, row) if not math.isnan(value)}


def test_edit_oracle_matches_reference_golden(golden):
    g = golden('metrics')
    loudness, pitch, periodicity, ppg = make_golden.metric_inputs(300, 57, 8)
    for index, row in enumerate(g['edit_arguments'].tolist()):
        outputs = metrics.edit_from_features(loudness, pitch, periodicity, ppg[0], **edit_arguments(row))
        for name, value in zip(EDIT_NAMES, outputs):
            assert torch.equal(value, g[f'edit_{index}_{name}']), (index, name)
    # the frame counts of the two evaluation ratios (config/defaults.py:204)
    # round(T / ratio + 1e-4), the count of the reference's own branch at edit/core.py:82
    assert g['edit_0_pitch'].shape[-1] == round(57 / .717 + 1e-4) == 79
    assert g['edit_1_pitch'].shape[-1] == round(57 / 1.414 + 1e-4) == 40


def test_metrics_special_cases():
    m = metrics.Metrics()
    assert math.isnan(m()['pitch']) and math.isnan(m()['loudness']) and 'ppg' not in m()
    frames = 16
    loudness = torch.full((8, frames), -30.)
    pitch = torch.full((1, frames), 220.)
    voiced = torch.full((1, frames), .5)
    # an octave apart, all voiced and loud: 1200 cents, loud == both, no quiet frame
    m.update(loudness, pitch, voiced, None, loudness + 3., 2 * pitch, voiced, None)
    result = m()
    assert result['pitch'] == pytest.approx(1200.) and result['periodicity'] == 0.
    assert result['loudness'] == pytest.approx(3.) == result['loudness-loud']
    assert math.isnan(result['loudness-quiet'])
    # unvoiced frames do not count towards the pitch error
    m.reset()
    m.update(loudness, pitch, voiced * 0., None, loudness, 2 * pitch, voiced, None)
    assert math.isnan(m()['pitch']) and m()['periodicity'] == pytest.approx(.5)
    # identical PPGs are at distance zero, disjoint ones at sqrt(log 2)
    ppg = torch.zeros(1, 40, frames)
    ppg[:, 3] = 1.
    other = torch.zeros(1, 40, frames)
    other[:, 7] = 1.
    m.update(loudness, pitch, voiced, ppg, loudness, pitch, voiced, ppg)
    assert m()['ppg'] == pytest.approx(0., abs=1e-6)
    m.reset()
    m.update(loudness, pitch, voiced, ppg, loudness, pitch, voiced, other)
    assert m()['ppg'] == pytest.approx(math.sqrt(math.log(2.)), rel=1e-4)


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')
def test_oracle_metrics_against_live_reference():
    """The reference's own classes, through the shim, on fresh inputs"""
    promonet = ref_shim.load()
    theirs, ours = promonet.evaluate.Metrics(), metrics.Metrics()
    for index, (frames, rows, target_rows) in enumerate(((91, 8, 513), (12, 8, 8))):
        predicted = make_golden.metric_inputs(7 + index, frames, rows)
        target = make_golden.metric_inputs(17 + index, frames, target_rows)
        theirs.update(*predicted, *target)
        ours.update(*predicted, *target)
    assert set(theirs()) == set(ours())
    for name, value in theirs().items():
        assert same(ours()[name], value), name
    features = make_golden.metric_inputs(5, 33, 8)
    expected = promonet.edit.from_features(
        *[t.clone() for t in features[:3]], features[3][0].clone(), pitch_shift_cents=250.,
        time_stretch_ratio=.8, loudness_scale_db=-4.)
    actual = metrics.edit_from_features(
        *features[:3], features[3][0], pitch_shift_cents=250., time_stretch_ratio=.8,
        loudness_scale_db=-4.)
    assert all(torch.equal(a, e) for a, e in zip(actual, expected))


def test_validation_entry_points_refuse_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import promonet_b200
    from promonet_b200.train import evaluate
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.evaluate.Metrics()
    with pytest.raises(RuntimeError, match='CUDA'):
        promonet_b200.edit.from_features(
            torch.zeros(8, 4), torch.ones(1, 4), torch.zeros(1, 4), torch.zeros(40, 4),
            time_stretch_ratio=2.)
    with pytest.raises(RuntimeError, match='CUDA'):
        evaluate(None, 0, None, [])


def test_validation_conditions_are_the_reference_ones():
    from promonet_b200.train.evaluate import conditions
    assert conditions() == [
        'reconstruction', 'shifted-071', 'shifted-141', 'stretched-071', 'stretched-141',
        'scaled-071', 'scaled-141']
