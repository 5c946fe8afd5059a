"""Freeze golden vectors from the UNMODIFIED reference (container only).

    python -m oracle.make_golden

Imports /root/reference through oracle.ref_shim, constructs the reference
modules under seed 1234 and stores small input/output pairs in tests/golden/.
The weights themselves are not stored (57 MB): promonet_b200.model.init draws
bit-identical tensors from the same seed, and the per-tensor checksums stored
here prove it at test time.
"""
from pathlib import Path

import numpy as np
import torch

from oracle import inputs, ref_shim

GOLDEN = Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def checksums(state):
    return {
        k: float(v.double().abs().sum()) for k, v in state.items()
        if v.is_floating_point()}


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    promonet = ref_shim.load()
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator().eval()
    state = generator.state_dict()

    # Generator.forward (generator.py:116-135), 8-band and 513-row loudness
    result = {}
    for rows, frames, batch in ((8, 24, 2), (513, 16, 1)):
        loud, pitch, per, ppg, spk, sbr, lr = inputs.synthesis(
            batch, frames, seed=promonet.RANDOM_SEED + rows, loudness_rows=rows)
        with torch.no_grad():
            features = generator.prepare_features(loud, pitch, per, ppg)
            audio = generator(
                loud, pitch, per, ppg, spk, sbr, lr,
                generator.default_previous_samples)
        tag = f'r{rows}'
        result.update({
            f'{tag}_loudness': loud, f'{tag}_pitch': pitch,
            f'{tag}_periodicity': per, f'{tag}_ppg': ppg,
            f'{tag}_speakers': spk, f'{tag}_sbr': sbr, f'{tag}_lr': lr,
            f'{tag}_features': features, f'{tag}_audio': audio})
    names = sorted(checksums(state))
    result['checksum_names'] = np.array(names)
    result['checksum_values'] = np.array([checksums(state)[n] for n in names])
    np.savez_compressed(
        GOLDEN / 'generator.npz',
        **{k: v.numpy() if torch.is_tensor(v) else v for k, v in result.items()})

    # Block.forward (hifigan.py:198-210) at C=32, k=3 and C=64, k=11
    blocks = {}
    for channels, kernel, length in ((32, 3, 300), (64, 11, 200)):
        torch.manual_seed(promonet.RANDOM_SEED + channels)
        block = promonet.model.hifigan.Block(channels, kernel, (1, 3, 5)).eval()
        x = torch.randn(2, channels, length)
        with torch.no_grad():
            y = block(x)
        tag = f'c{channels}k{kernel}'
        blocks[f'{tag}_x'] = x.numpy()
        blocks[f'{tag}_y'] = y.numpy()
        for key, value in block.state_dict().items():
            blocks[f'{tag}_{key}'] = value.numpy()
    np.savez_compressed(GOLDEN / 'block.npz', **blocks)

    # spectrogram.from_audio (preprocess/spectrogram.py:15-60,111-135)
    audio = inputs.audio(2, 4096, seed=promonet.RANDOM_SEED)[:, None]
    with torch.no_grad():
        linear = promonet.preprocess.spectrogram.from_audio(audio)
        mels = promonet.preprocess.spectrogram.from_audio(audio, mels=True)
    np.savez_compressed(
        GOLDEN / 'spectrogram.npz',
        audio=audio.numpy(), linear=linear.numpy(), mels=mels.numpy())

    # loudness.band_average / normalize (preprocess/loudness.py:84-146)
    loudness = torch.rand(513, 12) * 100. - 100.
    np.savez_compressed(
        GOLDEN / 'loudness_bands.npz',
        loudness=loudness.numpy(),
        averaged=promonet.preprocess.loudness.band_average(loudness, 8).numpy(),
        normalized=promonet.preprocess.loudness.normalize(loudness).numpy())
    print('wrote', sorted(p.name for p in GOLDEN.iterdir()))


def fargan():
    """config/fargan.py needs its own process (derived constants freeze at import)"""
    GOLDEN.mkdir(parents=True, exist_ok=True)
    promonet = ref_shim.load(ref_shim.REFERENCE_ROOT + '/config/fargan.py')
    assert promonet.MODEL == 'fargan'
    torch.manual_seed(promonet.RANDOM_SEED)
    generator = promonet.model.Generator().eval()
    state = generator.state_dict()
    loud, pitch, per, ppg, spk, sbr, lr = inputs.synthesis(2, 12, seed=99)
    previous = 0.1 * torch.randn(2, 1, promonet.NUM_PREVIOUS_SAMPLES)
    with torch.no_grad():
        audio = generator(loud, pitch, per, ppg, spk, sbr, lr, torch.zeros_like(previous))
        audio_previous = generator(loud, pitch, per, ppg, spk, sbr, lr, previous)
    names = sorted(checksums(state))
    np.savez_compressed(
        GOLDEN / 'fargan.npz',
        loudness=loud.numpy(), pitch=pitch.numpy(), periodicity=per.numpy(), ppg=ppg.numpy(),
        speakers=spk.numpy(), sbr=sbr.numpy(), lr=lr.numpy(), previous=previous.numpy(),
        audio=audio.numpy(), audio_previous=audio_previous.numpy(),
        checksum_names=np.array(names),
        checksum_values=np.array([checksums(state)[n] for n in names]))
    print('wrote fargan.npz')


if __name__ == '__main__':
    import sys
    fargan() if '--fargan' in sys.argv else main()
