"""Multi-resolution spectral convergence (promonet/train/loss.py:61-150; flag
SPECTRAL_CONVERGENCE_LOSS, config/defaults.py:358: off by default, on in config/fargan.py:23).

Each of the six STFTs (n_fft = window 2560 ... 80, hop n_fft / 4, hann, center=True) is a
1 x 1 convolution of the reflect-padded signal, read in place as overlapping frames, with the
windowed DFT basis as its weight: forward and backward are the training conv kernels.
"""
import torch

from promonet_b200.train import ops

FFT_SIZES = (2560, 1280, 640, 320, 160, 80)   # loss.py:129-131


class MultiResolutionSpectralConvergence:

    def __init__(self, device, math='tf32'):
        self.math = math
        self.resolutions = []
        for n_fft in FFT_SIZES:
            basis = ops.dft_basis(n_fft, device)          # (2 bins, n_fft)
            rows = basis.shape[0]
            if math == 'tf32':
                forward = ops.pack_weight_taps(
                    basis, torch.empty(ops.packed_floats(rows, n_fft, 1), device=device),
                    rows, n_fft, 1, False)
                backward = ops.pack_weight_taps(
                    basis, torch.empty(ops.packed_floats(n_fft, rows, 1), device=device),
                    rows, n_fft, 1, True)
            else:
                forward = basis
                backward = ops.transpose_weight(basis, torch.empty_like(basis), rows, n_fft, 1)
            self.resolutions.append((n_fft, n_fft // 4, rows, forward, backward))
        self.sums = torch.zeros(2, device=device)

    def conv(self, geometry, a, weight, out):
        if self.math == 'tf32':
            return ops.conv_gemm_tc(geometry, False, a, weight, out)
        return ops.conv_gemm(geometry, False, a, weight, out)

    def __call__(self, both, batch, weight, loss, ggenerated):
        """both (2 B, 1, T): target audio then generated audio.  *loss += weight * mean over
        the resolutions of ||s_y - s_x||_1 / ||s_y||_1; ggenerated (B, 1, T) += its gradient."""
        samples = both.shape[-1]
        signals = both.view(2 * batch, samples)
        for n_fft, hop, rows, forward, backward in self.resolutions:
            padded = ops.reflect_pad(signals, n_fft // 2, n_fft // 2)       # center=True
            length = samples + n_fft
            frames = 1 + samples // hop
            geometry = ops.geometry(
                2 * batch, n_fft, rows, (frames, 1), (1, 1), strides=(1, hop, length))
            spec = self.conv(
                geometry, padded, forward, torch.empty(2 * batch, rows, frames, device=both.device))
            gspec = torch.empty(batch, rows, frames, device=both.device)
            ops.spectral_convergence(spec, batch, weight / len(self.resolutions), self.sums, loss, gspec)
            geometry = ops.geometry(batch, rows, n_fft, (frames, 1), (1, 1))
            gframes = self.conv(
                geometry, gspec, backward, torch.empty(batch, n_fft, frames, device=both.device))
            gpadded = torch.zeros(batch, length, device=both.device)
            ops.frame_overlap_add(gframes, gpadded, hop)
            ops.reflect_pad_backward(
                gpadded, ggenerated.view(batch, samples), n_fft // 2, n_fft // 2, accumulate=True)
