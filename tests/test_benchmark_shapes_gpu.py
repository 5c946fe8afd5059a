"""Parity at the shapes the benchmark numbers are quoted on (bench.py, BASELINE.json
configs[1..3]), so that the headline is not produced on an unverified shape: batch
offsets past 2^31 bytes, > 148 tiles per persistent kernel, two Viterbi waves, 8-item
training batches.  Same bars as the small-shape tests; the CPU oracle runs on a few
sampled utterances."""
import numpy as np
import pytest
import torch

from conftest import relative_error
from oracle import dsp, hifigan, inputs
from oracle import penn as oracle_penn
from oracle import train as oracle_train
from oracle import viterbi as oracle_viterbi

pytestmark = pytest.mark.gpu

TOLERANCE = 1e-4


def test_hifigan_batch_32_by_430_frames():
    """configs[1]: B = 32, F = 430 -> (32, 1, 110 080); utterances 0, 13 and 31 against the
    oracle run one utterance at a time (Generator.forward, generator.py:116-135)"""
    import promonet_b200
    state = promonet_b200.model.init.hifigan_state(promonet_b200.RANDOM_SEED)
    model = promonet_b200.model.Generator(state=state)
    args = inputs.synthesis(32, 430, seed=1234)
    audio = model(*[a.cuda() for a in args])
    again = model(*[a.cuda() for a in args])
    assert audio.shape == (32, 1, 430 * 256)
    assert torch.equal(audio, again)                    # no race between tiles
    assert bool(torch.isfinite(audio).all())
    for index in (0, 13, 31):
        with torch.no_grad():
            expected = hifigan.generator(state, *[a[index:index + 1] for a in args])
        assert relative_error(audio[index:index + 1], expected) < TOLERANCE, index
    # the host entry point (what bench.py's e2e times) returns the same bits
    host = model.forward_host(*args)
    assert torch.equal(host, audio.cpu())


@pytest.mark.parametrize('seed', [1234, 7, 99])
def test_hifigan_with_fp8_corrections_batch_32_by_430_frames(seed):
    """The same shape with the residual blocks of the C = 256 / 128 stages on "fp16 + 2 x fp8"
    operands (Generator(f8=True), pmn_generator_set_f8): three seeds for weights and inputs, two
    utterances each against the oracle; the same 1e-4 bar"""
    import promonet_b200
    state = promonet_b200.model.init.hifigan_state(seed)
    model = promonet_b200.model.Generator(state=state, f8=True)
    args = inputs.synthesis(32, 430, seed=seed)
    audio = model(*[a.cuda() for a in args])
    again = model(*[a.cuda() for a in args])
    assert torch.equal(audio, again)
    assert bool(torch.isfinite(audio).all())
    errors = []
    for index in (3, 31):
        with torch.no_grad():
            expected = hifigan.generator(state, *[a[index:index + 1] for a in args])
        errors.append(relative_error(audio[index:index + 1], expected))
    print(f'fp8-corrected generator, seed {seed}: relative errors {errors}')
    assert max(errors) < TOLERANCE, errors


def test_preprocess_batch_32_by_10_seconds():
    """configs[2]: 32 utterances x 220 500 samples -> 861 frames each"""
    import promonet_b200
    pb = promonet_b200
    audio = inputs.audio(32, 220500, seed=1234)
    features = ['loudness', 'pitch', 'periodicity', 'mels']
    loudness, pitch, periodicity, mels = pb.preprocess.from_audio_batch(audio, features=features)
    assert loudness.shape == (32, 8, 861) and pitch.shape == (32, 861) and mels.shape == (32, 80, 861)
    model = pb.preprocess.core._pitch_model(torch.device('cuda', torch.cuda.current_device()))
    pitch2, periodicity2, logits, bins = model(audio, return_intermediates=True)
    assert torch.equal(pitch2, pitch) and torch.equal(periodicity2, periodicity)
    # Viterbi, all 32 utterances (two waves of clusters), bit-exact on identical log inputs:
    # the device's own logits -> log-softmax on the host -> both decoders
    with torch.no_grad():
        log_posterior = torch.log_softmax(logits.cpu(), dim=-1).numpy()
    transition = oracle_penn.transition_matrix(256 / 22050)
    initial = oracle_penn.initial_distribution()
    with np.errstate(divide='ignore'):
        log_transition, log_initial = np.log(transition.numpy()), np.log(initial.numpy())
    expected = oracle_viterbi.decode(log_posterior, None, log_transition, log_initial, log_probs=True)
    actual = pb.preprocess.viterbi.from_probabilities(
        torch.from_numpy(log_posterior), None, torch.from_numpy(log_transition),
        torch.from_numpy(log_initial), log_probs=True)
    assert np.array_equal(actual.cpu().numpy(), expected)
    # the pipeline's own path (logs taken on the device) agrees with it except at rounding ties
    assert float((bins.cpu().numpy() == expected).mean()) > 0.99
    # the float features of sampled utterances against the CPU chain
    state = oracle_penn.init_state(1234)
    for index in (0, 31):
        one = audio[index:index + 1]
        assert relative_error(loudness[index], dsp.loudness(one, 8)) < TOLERANCE
        assert relative_error(mels[index:index + 1], dsp.linear_to_mel(dsp.magnitude(one.double()))) < TOLERANCE
        _, expected_periodicity, aux = oracle_penn.from_audio(state, one)
        expected_logits, _, _ = oracle_penn.postprocess(aux['logits'])
        live = torch.isfinite(expected_logits)
        assert relative_error(logits[index].cpu()[live], expected_logits[live]) < TOLERANCE
        assert float((periodicity[index].cpu() - expected_periodicity[0]).abs().max()) < TOLERANCE


def test_train_step_8_by_64_frames_in_the_shipped_math():
    """configs[3]: 8 items x 16 384 samples, Trainer's default math='tf32' (tf32 tensor-core
    products, fp32 accumulation and storage) against the fp32 autograd oracle.  Stated bar:
    generated audio and the five losses within 5e-3 relative (10-bit operand mantissas, the
    same as the reference's fp16 autocast, train/core.py:220); each module's whole gradient
    within 3 % of the oracle's in norm and at cosine similarity > 0.999 (discriminator) /
    0.99 (generator, whose gradient passes through the discriminator's tf32 data gradients)."""
    from promonet_b200.model import init
    from promonet_b200.train.core import Trainer
    states = init.hifigan_state(1234), init.discriminator_state(1234)
    batch = oracle_train.batch(8, 64, seed=1234)
    g_state = oracle_train.leaf_state(states[0])
    d_state = oracle_train.leaf_state(states[1])
    losses, g_grads, d_grads, generated = oracle_train.step(g_state, d_state, batch)
    trainer = Trainer(*states)
    assert trainer.math == 'tf32'
    ours = trainer.step(*[t.cuda().contiguous() for t in batch], update=False).cpu()
    assert relative_error(trainer.generated, generated) < 5e-3
    for i, name in enumerate(('discriminator', 'mel', 'feature_matching', 'adversarial', 'generator')):
        assert abs(float(ours[i]) - float(losses[name])) < 5e-3 * abs(float(losses[name])), name
    report = {}
    for kind, module, expected, bar in (
            ('discriminator', trainer.discriminators, d_grads, 0.999),
            ('generator', trainer.generator, g_grads, 0.99)):
        actual = module.params.gradients()
        a = torch.cat([actual[name].double().cpu().reshape(-1) for name in expected])
        b = torch.cat([expected[name].double().reshape(-1) for name in expected])
        cosine = float(torch.dot(a, b) / (a.norm() * b.norm()))
        ratio = float(a.norm() / b.norm())
        report[kind] = (cosine, ratio)
        assert cosine > bar and abs(ratio - 1.) < 3e-2, report
    print('tf32 step vs fp32 autograd (cosine, norm ratio):', report)
