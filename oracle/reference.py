"""The reference's own modules as a timed baseline (bench.py only).

`load()` returns the unmodified `promonet` package -- from /root/reference in
the build container, from the git-ignored copy oracle/_ref/ (made by
oracle/vendor_reference.py) on the GPU box -- imported behind oracle.ref_shim's
stub modules, or None when neither tree exists (bench.py then falls back to the
oracle port and says so in `cpu_baseline.kind`)."""
import os
from pathlib import Path

from oracle import ref_shim

VENDORED = Path(__file__).resolve().parent / '_ref'


def root():
    for candidate in (os.environ.get('PROMONET_REFERENCE'), '/root/reference', str(VENDORED)):
        if candidate and os.path.isdir(os.path.join(candidate, 'promonet')):
            return candidate
    return None


def load(config_file=None):
    where = root()
    if where is None:
        return None
    ref_shim.REFERENCE_ROOT = where
    return ref_shim.load(config_file)


def generator(promonet, seed=1234):
    """promonet.model.Generator() as the reference constructs it
    (promonet/train/core.py:58), under the reference's seed, in eval mode"""
    import torch
    torch.manual_seed(seed)
    return promonet.model.Generator().eval()


def forward(model, loudness, pitch, periodicity, ppg, speakers, sbr, lr):
    """Generator.forward (promonet/model/generator.py:116-135), no autocast"""
    previous = model.default_previous_samples.to(loudness.device)
    return model(loudness, pitch, periodicity, ppg, speakers, sbr, lr, previous)


class TrainingStep:
    """One iteration of the reference's training loop body (promonet/train/core.py:183-369)
    composed from the unmodified reference modules, optimizers and losses, for timing it on
    the same GPU as a baseline: `autocast=True` is how the reference trains (fp16 autocast +
    GradScaler, :118,220,262), False is plain fp32 (or TF32 where torch's flags allow it)"""

    def __init__(self, promonet, device, autocast, seed=1234):
        import torch
        self.promonet, self.device, self.autocast = promonet, device, autocast
        torch.manual_seed(seed)
        self.generator = promonet.model.Generator().to(device)
        torch.manual_seed(seed)
        self.discriminators = promonet.model.Discriminator().to(device)
        self.discriminator_optimizer = promonet.OPTIMIZER(self.discriminators.parameters())
        self.generator_optimizer = promonet.OPTIMIZER(self.generator.parameters())
        self.scaler = torch.amp.GradScaler('cuda', enabled=autocast)
        self.previous_samples = torch.zeros(promonet.HOPSIZE, device=device)

    def __call__(self, loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio):
        import torch
        promonet = self.promonet
        context = lambda: torch.autocast('cuda', torch.float16, enabled=self.autocast)
        with context():
            generated = self.generator(
                loudness, pitch, periodicity, ppg, speakers, sbr, lr, self.previous_samples)
            real_logits, fake_logits, _, _ = self.discriminators(audio, generated.detach())
            discriminator_losses, _, _ = promonet.loss.discriminator(
                [logit.float() for logit in real_logits], [logit.float() for logit in fake_logits])
        self.discriminator_optimizer.zero_grad()
        self.scaler.scale(discriminator_losses).backward()
        self.scaler.step(self.discriminator_optimizer)
        with context():
            _, fake_logits, real_maps, fake_maps = self.discriminators(audio, generated)
            mels = promonet.preprocess.spectrogram.linear_to_mel(spectrograms, None)
            generated_mels = promonet.preprocess.spectrogram.from_audio(generated.float(), True, None)
            mel_loss = torch.nn.functional.l1_loss(mels, generated_mels)
            feature_matching_loss = promonet.loss.feature_matching(real_maps, fake_maps)
            adversarial_loss, _ = promonet.loss.generator([logit.float() for logit in fake_logits])
            generator_losses = (
                promonet.MEL_LOSS_WEIGHT * mel_loss +
                promonet.FEATURE_MATCHING_LOSS_WEIGHT * feature_matching_loss +
                promonet.ADVERSARIAL_LOSS_WEIGHT * adversarial_loss)
        self.generator_optimizer.zero_grad()
        self.scaler.scale(generator_losses).backward()
        self.scaler.step(self.generator_optimizer)
        self.scaler.update()
        return generator_losses


def fargan_eager(batch, frames, seed=1234):
    """The reference FARGAN generator (config/fargan.py) run eagerly by PyTorch on cuda:0, one
    forward of `batch` utterances x `frames` frames after a short warm-up: ms and samples / s.
    Run in its own process (`python -m oracle.reference --fargan-eager B F`): the reference
    freezes its configuration at import."""
    import json
    import time
    import torch
    where = root()
    promonet = load(os.path.join(where, 'config', 'fargan.py')) if where else None
    if promonet is None:
        return {'unavailable': 'the reference tree (oracle/_ref) did not travel to this box'}
    from promonet_b200 import synthetic
    device = torch.device('cuda', 0)
    model = generator(promonet, seed).to(device)
    inputs = [t.to(device) for t in synthetic.synthesis(batch, frames, seed=seed)]
    previous = torch.zeros(batch, 1, promonet.NUM_PREVIOUS_SAMPLES, device=device)

    def forward(count):
        args = [t[..., :count] if t.ndim > 1 else t for t in inputs]
        with torch.inference_mode():
            return model(*args, previous)
    forward(min(frames, 4))          # warm-up: kernels loaded, cuBLAS handles made
    torch.cuda.synchronize()
    begin = time.perf_counter()
    audio = forward(frames)
    torch.cuda.synchronize()
    seconds = time.perf_counter() - begin
    return {
        'what': f'promonet.model.Generator (config/fargan.py) forward on cuda, {batch} x {frames} frames, '
                f'torch {torch.__version__} eager fp32 (TF32 off), one forward after a 4-frame warm-up',
        'unit': 'samples/s', 'ms_per_step': seconds * 1e3,
        'value': audio.shape[0] * audio.shape[-1] / seconds,
        'finite': bool(torch.isfinite(audio).all())}


if __name__ == '__main__':
    import json
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == '--fargan-eager':
        print(json.dumps(fargan_eager(int(sys.argv[2]), int(sys.argv[3]))))
