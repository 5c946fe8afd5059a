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
