"""Parity of the in-training validation path through the C ABI:
promonet_b200.evaluate.Metrics (pmn_metrics_update), promonet_b200.edit.from_features
(pmn_edit_contour / pmn_grid_sample) and promonet_b200.train.evaluate, against the golden
vectors of the reference's own classes and against the CPU oracle.  Scalars: 1e-5 relative
(the sums run in fp32 per frame and in double across frames)."""
import json
import math

import pytest
import torch

from conftest import relative_error
from oracle import dsp, hifigan, inputs, make_golden
from oracle import metrics as oracle_metrics
from promonet_b200.model import init

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-5


def close(actual, expected, rel=TOLERANCE):
    return (math.isnan(actual) and math.isnan(expected)) or actual == pytest.approx(
        expected, rel=rel, abs=1e-7)


def cuda(tensors):
    return [None if t is None else t.cuda() for t in tensors]


@pytest.fixture(scope='module')
def pb():
    import promonet_b200
    return promonet_b200


###############################################################################
# Metrics
###############################################################################


def test_metrics_match_reference_golden(pb, golden):
    """promonet.evaluate.Metrics of the reference itself, per update and accumulated"""
    g = golden('metrics')
    names = [str(name) for name in g['names']]
    accumulated = pb.evaluate.Metrics()
    single = pb.evaluate.Metrics()
    for index, (frames, rows, target_rows) in enumerate(g['cases'].tolist()):
        predicted = make_golden.metric_inputs(100 + index, frames, rows)
        target = make_golden.metric_inputs(200 + index, frames, target_rows)
        single.reset()
        single.update(*cuda(predicted), *cuda(target))
        if index < 3:
            accumulated.update(*cuda(predicted), *cuda(target))
        result = single()
        assert list(result) == ['pitch', 'periodicity', 'ppg', 'loudness', 'loudness-loud', 'loudness-quiet']
        for name, expected in zip(names, g[f'single_{index}'].tolist()):
            assert close(result[name], expected), (index, name, result[name], expected)
    for name, expected in zip(names, g['accumulated'].tolist()):
        assert close(accumulated()[name], expected), name


@pytest.mark.parametrize('frames', [1, 127, 128, 129, 1000])
def test_batched_update_equals_consecutive_updates(pb, frames):
    batch = 3
    predicted = [make_golden.metric_inputs(10 + i, frames, 8) for i in range(batch)]
    target = [make_golden.metric_inputs(20 + i, frames, 513) for i in range(batch)]
    oracle = oracle_metrics.Metrics()
    one_by_one = pb.evaluate.Metrics()
    for p, t in zip(predicted, target):
        oracle.update(*p, *t)
        one_by_one.update(*cuda(p), *cuda(t))
    stack = lambda items, i: torch.stack([item[i] for item in items]).cuda()
    batched = pb.evaluate.Metrics()
    batched.update(
        stack(predicted, 0), stack(predicted, 1)[:, 0], stack(predicted, 2)[:, 0], stack(predicted, 3)[:, 0],
        stack(target, 0), stack(target, 1)[:, 0], stack(target, 2)[:, 0], stack(target, 3)[:, 0])
    for name, expected in oracle().items():
        assert close(one_by_one()[name], expected), name
        assert close(batched()[name], expected), name


def test_metrics_without_ppg_and_with_similarity(pb):
    frames = 77
    predicted = make_golden.metric_inputs(1, frames, 8)
    target = make_golden.metric_inputs(2, frames, 8)
    m = pb.evaluate.Metrics()
    m.update(*cuda(predicted[:3]), None, *cuda(target[:3]), None)
    oracle = oracle_metrics.Metrics()
    oracle.update(*predicted[:3], None, *target[:3], None)
    assert 'ppg' not in m() and set(m()) == set(oracle())
    for name, expected in oracle().items():
        assert close(m()[name], expected), name
    # a phoneme-similarity matrix (rows sum to one) in front of the distance
    generator = torch.Generator().manual_seed(3)
    similarity = torch.softmax(4. * torch.eye(40) + torch.randn(40, 40, generator=generator), 1)
    m = pb.evaluate.Metrics(similarity=similarity)
    m.update(*cuda(predicted), *cuda(target))
    oracle = oracle_metrics.Metrics(similarity=similarity)
    oracle.update(*predicted, *target)
    assert close(m()['ppg'], oracle()['ppg'], 1e-4)
    with pytest.raises(ValueError):
        m.update(*cuda(predicted), *cuda(target[:3]), None)
    with pytest.raises(NotImplementedError):
        m.update(*cuda(predicted), *cuda(target), 'a', 'b')


def test_metrics_special_values(pb):
    frames = 16
    loudness = torch.full((8, frames), -30.).cuda()
    pitch = torch.full((1, frames), 220.).cuda()
    voiced = torch.full((1, frames), .5).cuda()
    m = pb.evaluate.Metrics()
    assert math.isnan(m()['pitch']) and 'ppg' not in m()
    m.update(loudness, pitch, voiced, None, loudness + 3., 2 * pitch, voiced, None)
    result = m()
    assert result['pitch'] == pytest.approx(1200., rel=1e-6) and result['periodicity'] == 0.
    assert result['loudness'] == pytest.approx(3.) == result['loudness-loud']
    assert math.isnan(result['loudness-quiet'])
    m.reset()
    m.update(loudness, pitch, voiced * 0., None, loudness, 2 * pitch, voiced, None)
    assert math.isnan(m()['pitch']) and m()['periodicity'] == pytest.approx(.5)


###############################################################################
# Edits
###############################################################################


def edit_arguments(row):
    keys = ('pitch_shift_cents', 'time_stretch_ratio', 'loudness_scale_db')
    return {key: value for key, value in zip(keys, row) if not math.isnan(value)}


def test_edit_matches_reference_golden(pb, golden):
    """promonet.edit.from_features of the reference itself"""
    g = golden('metrics')
    loudness, pitch, periodicity, ppg = cuda(make_golden.metric_inputs(300, 57, 8))
    kept = loudness.clone()
    for index, row in enumerate(g['edit_arguments'].tolist()):
        outputs = pb.edit.from_features(loudness, pitch, periodicity, ppg[0], **edit_arguments(row))
        for name, value in zip(('loudness', 'pitch', 'periodicity', 'ppg'), outputs):
            expected = g[f'edit_{index}_{name}']
            assert value.shape == expected.shape, (index, name)
            assert relative_error(value, expected) < TOLERANCE, (index, name)
    assert torch.equal(loudness, kept)        # inputs are not edited in place
    *_, grid = pb.edit.from_features(
        loudness, pitch, periodicity, ppg[0], time_stretch_ratio=2., return_grid=True)
    assert grid.shape == (29,) and float(grid[0]) == 0. and float(grid[-1]) == 56.
    with pytest.raises(NotImplementedError):
        pb.edit.from_features(
            loudness, pitch, periodicity, ppg[0], time_stretch_ratio=2., stretch_silence=False)


def test_edit_batched_contours(pb):
    loudness, pitch, periodicity, ppg = inputs.synthesis(3, 50, seed=5)[:4]
    batched = pb.edit.from_features(
        *cuda((loudness, pitch, periodicity, ppg)), pitch_shift_cents=-300., time_stretch_ratio=1.3,
        loudness_scale_db=2.)
    for b in range(3):
        expected = oracle_metrics.edit_from_features(
            loudness[b], pitch[b:b + 1], periodicity[b:b + 1], ppg[b], pitch_shift_cents=-300.,
            time_stretch_ratio=1.3, loudness_scale_db=2.)
        assert relative_error(batched[0][b], expected[0]) < TOLERANCE
        assert relative_error(batched[1][b:b + 1], expected[1]) < TOLERANCE
        assert relative_error(batched[2][b:b + 1], expected[2]) < TOLERANCE
        assert relative_error(batched[3][b], expected[3]) < TOLERANCE


###############################################################################
# train.evaluate
###############################################################################


class Module:
    """Stands in for the generator being trained: evaluate only needs its state dict"""

    def __init__(self, state):
        self.state = state

    def state_dict(self):
        return self.state


def validation_loader(items, frames):
    loader = []
    for index in range(items):
        loudness, pitch, periodicity, ppg, speakers, _, _ = inputs.synthesis(1, frames, seed=40 + index)
        # audio a few samples longer than a whole number of hops: evaluate trims it (:560-562)
        audio = inputs.audio(1, frames * 256 + 100, seed=index)[:, None]
        loader.append((
            ['text'], loudness, pitch, periodicity * .5, ppg, speakers, torch.ones(1), torch.ones(1),
            torch.zeros(1, 513, frames), audio, [f'stem-{index}']))
    return loader


def test_evaluate_matches_the_oracle_composition(pb, tmp_path):
    """Every scalar of train.evaluate equals the CPU oracle's Metrics over (edited inputs,
    features of the audio evaluate generated); the audio of every condition is the generator's
    output for the edited inputs; loudness and periodicity errors of the reconstruction also
    agree with the all-CPU chain (oracle generator -> oracle features)"""
    from oracle import penn as oracle_penn
    from promonet_b200.train import evaluate
    frames, items = 40, 2
    state = init.hifigan_state(1234)
    loader = validation_loader(items, frames)
    scalars, waveforms = evaluate(tmp_path, 0, Module(state), loader + loader, 0, evaluation_steps=items)
    stored = json.loads((tmp_path / 'evaluation-00000000.json').read_text())
    assert stored['step'] == 0 and stored['scalars'].keys() == scalars.keys()
    conditions = [
        'reconstruction', 'shifted-071', 'shifted-141', 'stretched-071', 'stretched-141',
        'scaled-071', 'scaled-141']
    assert sorted(scalars) == sorted(
        f'{c}/{m}' for c in conditions
        for m in ('pitch', 'periodicity', 'loudness', 'loudness-loud', 'loudness-quiet'))
    assert len(waveforms) == items * (len(conditions) + 1)       # + the original at step 0
    generator = pb.model.Generator(state=state)
    expected = {c: oracle_metrics.Metrics() for c in conditions}
    chain = oracle_metrics.Metrics()
    for i, batch in enumerate(loader):
        _, loudness, pitch, periodicity, ppg, speakers, _, _, _, audio, _ = batch
        assert torch.equal(waveforms[f'original/{i:02d}-audio'].cpu(), audio[0, :, :frames * 256])
        edited = {'reconstruction': (loudness, pitch, periodicity, ppg)}
        for ratio, tag in ((.717, '071'), (1.414, '141')):
            edited[f'shifted-{tag}'] = (loudness, ratio * pitch, periodicity, ppg)
            edited[f'scaled-{tag}'] = (loudness + 10 * math.log2(ratio), pitch, periodicity, ppg)
            # the device's own edit (held to the reference in test_edit_matches_reference_golden):
            # a 1e-7 difference in a stretched pitch could cross a pitch-bin edge of the generator
            edited[f'stretched-{tag}'] = tuple(t.cpu() for t in pb.edit.from_features(
                *cuda((loudness, pitch, periodicity, ppg)), time_stretch_ratio=ratio))
            assert edited[f'stretched-{tag}'][1].shape == (1, round(frames / ratio + 1e-4))
        for condition, features in edited.items():
            audio_out = waveforms[f'{condition}/{i:02d}-audio']
            assert audio_out.shape == (1, features[1].shape[-1] * 256)
            alone = generator(
                *cuda(features), speakers.cuda(), torch.ones(1).cuda(), torch.ones(1).cuda())
            assert relative_error(audio_out, alone[0]) < 1e-4, condition
            predicted = pb.preprocess.from_audio(audio_out, gpu=0)
            expected[condition].update(
                *features[:3], None, *[t.cpu() for t in predicted], None)
        with torch.no_grad():
            reference_audio = hifigan.generator(
                state, loudness, pitch, periodicity, ppg, speakers, torch.ones(1), torch.ones(1))[0]
        _, chain_periodicity, _ = oracle_penn.from_audio(oracle_penn.init_state(1234), reference_audio)
        chain.update(
            loudness, pitch, periodicity, None,
            dsp.loudness(reference_audio, 8), pitch, chain_periodicity, None)
    for condition in conditions:
        for name, value in expected[condition]().items():
            assert close(scalars[f'{condition}/{name}'], value), (condition, name)
    for name in ('loudness', 'periodicity'):
        assert close(scalars[f'reconstruction/{name}'], chain()[name], 1e-3), name


def test_evaluate_with_a_ppg_model_and_the_inference_generator(pb):
    from promonet_b200.train import evaluate
    generator = pb.model.Generator(state=init.hifigan_state(7))

    def ppg_model(audio):
        frames = audio.shape[-1] // 256
        torch.manual_seed(frames)
        return torch.softmax(torch.randn(audio.shape[0], 40, frames, device=audio.device), 1)

    scalars, waveforms = evaluate(None, 5, generator, validation_loader(1, 24), ppg_model=ppg_model)
    assert 'original/00-audio' not in waveforms and len(waveforms) == 7
    assert all(0. < scalars[f'{c}/ppg'] < 1. for c in ('reconstruction', 'stretched-141', 'scaled-071'))
