"""Oracle: STFT / magnitude / log-mel / A-weighted loudness (test infrastructure).

* spectrogram path restates promonet/preprocess/spectrogram.py:15-60,111-135
  (pinned against the reference module, see oracle/make_golden.py);
* loudness path restates promonet/preprocess/loudness.py:17-55,149-160 whose
  arithmetic is librosa's (un-vendored, unpinned in setup.py:18): PARITY
  UNPINNED -- restated from librosa's published definitions of `stft`,
  `amplitude_to_db`, `A_weighting`, `fft_frequencies` and `filters.mel`.
"""
import numpy as np
import torch

SAMPLE_RATE = 22050  # promonet/config/defaults.py:49
HOPSIZE = 256        # :31
NUM_FFT = 1024       # :43
WINDOW_SIZE = 1024   # :52
NUM_MELS = 80        # :40
MIN_DB = -100.       # :37
REF_DB = 20.         # :46


def hann(dtype=torch.float64):
    """Periodic hann: torch.hann_window == scipy get_window('hann', fftbins=True)"""
    return torch.hann_window(WINDOW_SIZE, dtype=dtype)


def frames(audio):
    """Reflect-pad (NUM_FFT - HOPSIZE) // 2 and frame (spectrogram.py:35-37)"""
    size = (NUM_FFT - HOPSIZE) // 2
    padded = torch.nn.functional.pad(audio, (size, size), mode='reflect')
    return padded.unfold(-1, NUM_FFT, HOPSIZE)  # (..., F, 1024)


def stft(audio, window=None):
    """(B, 1, T) or (B, T) -> complex (B, 513, F); center=False"""
    if audio.ndim == 3:
        audio = audio.squeeze(1)
    x = frames(audio[:, None]).squeeze(1)
    if window is not None:
        x = x * window.to(x.dtype)
    return torch.fft.rfft(x, dim=-1).transpose(-1, -2)


def magnitude(audio):
    """spectrogram.py:39-52: sqrt(re^2 + im^2 + 1e-6)"""
    spec = torch.view_as_real(stft(audio, hann(audio.dtype)))
    return torch.sqrt(spec.pow(2).sum(-1) + 1e-6)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / (200. / 3)
    log_region = f >= 1000.
    with np.errstate(divide='ignore', invalid='ignore'):
        mel_log = 15. + np.log(f / 1000.) / (np.log(6.4) / 27.)
    return np.where(log_region, mel_log, mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = m * (200. / 3)
    log_region = m >= 15.
    return np.where(log_region, 1000. * np.exp((np.log(6.4) / 27.) * (m - 15.)), f)


def mel_basis(sr=SAMPLE_RATE, n_fft=NUM_FFT, n_mels=NUM_MELS):
    """librosa.filters.mel(sr, n_fft, n_mels): Slaney scale + Slaney norm, float32"""
    fftfreqs = np.arange(n_fft // 2 + 1, dtype=np.float64) * sr / n_fft
    mel_f = _mel_to_hz(
        np.linspace(_hz_to_mel(0.), _hz_to_mel(sr / 2.), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2. / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def linear_to_mel(spectrogram, threshold=None):
    """spectrogram.py:111-135"""
    basis = torch.from_numpy(mel_basis()).to(spectrogram.dtype)
    mel = torch.log(torch.matmul(basis, spectrogram))
    if threshold is not None:
        mel = torch.clamp(mel, min=threshold)
    return mel


def a_weighting(frequencies, min_db=-80.):
    """librosa.A_weighting"""
    f_sq = np.asarray(frequencies, dtype=np.float64) ** 2
    const = np.array([12194.217, 20.598997, 107.65265, 737.86223]) ** 2
    with np.errstate(divide='ignore'):
        weights = 2. + 20. * (
            np.log10(const[0]) +
            2 * np.log10(f_sq) -
            np.log10(f_sq + const[0]) -
            np.log10(f_sq + const[1]) -
            .5 * np.log10(f_sq + const[2]) -
            .5 * np.log10(f_sq + const[3]))
    return np.maximum(min_db, weights)


def perceptual_weights():
    """loudness.py:149-160: A_weighting(fft_frequencies)[:, None] - REF_DB"""
    frequencies = np.arange(NUM_FFT // 2 + 1, dtype=np.float64) * SAMPLE_RATE / NUM_FFT
    return a_weighting(frequencies)[:, None] - REF_DB


def loudness(audio, bands=8):
    """loudness.py:17-55 for one utterance (1, T) -> (bands or 513, F)

    librosa.stft runs in the input dtype (float32 -> complex64) and
    amplitude_to_db(S) = 10 log10(max(1e-10, S^2)) then max(., max - 80).
    """
    from oracle import features
    x = frames(audio[None].to(torch.float32)).squeeze(0).squeeze(0)
    x = (x * hann(torch.float32)).numpy()
    spec = np.abs(np.fft.rfft(x.astype(np.float64), axis=-1)).T.astype(np.float32)
    power = np.square(spec)
    db = 10. * np.log10(np.maximum(1e-10, power))
    db = np.maximum(db, db.max() - 80.)
    weighted = db + perceptual_weights()
    weighted[weighted < MIN_DB] = MIN_DB
    result = torch.from_numpy(weighted).float()
    return features.band_average(result, bands) if bands is not None else result
