"""One GAN training step, sequenced as promonet/train/core.py:183-369

    generated = generator(batch)                                 :223
    D(audio, generated.detach()) -> LSGAN loss -> D AdamW step   :239-256
    D(audio, generated) again with the updated D                 :272
    45 x mel L1 + feature matching + LSGAN generator loss        :277-332
    generator backward -> G AdamW step                           :335-369

Differences from the reference, all deliberate: fp32 everywhere instead of fp16
autocast + GradScaler (:118,220,262); the generator step does not deposit (unused)
gradients in the discriminator (:338 does, :254 clears them); under
torch.distributed the gradients of each module are averaged with one NCCL
all-reduce before its optimizer step (the reference is single-GPU).
"""
import os
from pathlib import Path

import torch

from promonet_b200 import config, parallel
from promonet_b200.train import ops
from promonet_b200.train.discriminator import Discriminator
from promonet_b200.train.generator import Generator

LOSSES = (
    'discriminator', 'mel', 'feature_matching', 'adversarial', 'generator', 'spectral_convergence')


class Trainer:

    def __init__(self, generator_state=None, discriminator_state=None, device=None,
                 process_group=None, math='tf32', multi_scale_discriminator=False,
                 spectral_convergence_loss=False, peer_optimizer=True, data_parallel=True,
                 multi_resolution_discriminator=False):
        """math: 'tf32' runs the convolutions' forward and data gradients on the tensor
        cores (tf32 operands, fp32 accumulation; the reference trains under fp16 autocast,
        train/core.py:220); 'fp32' is the exact FMA path used for parity"""
        self.math = math
        self.process_group = process_group
        self.world = 1
        if data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        # Data parallel (data_parallel=False: a replica that ignores the process group): by default the gradient exchange is fused with the optimizer in one kernel
        # over NVLink peer memory (ParamSet.adamw_peer); peer_optimizer=False uses an NCCL
        # all-reduce of the flat gradient buffer followed by the local AdamW kernel
        peer_group = None
        if self.world > 1 and peer_optimizer:
            peer_group = process_group or torch.distributed.group.WORLD
        self.generator = Generator(generator_state, device, math, peer_group)
        # MULTI_SCALE_DISCRIMINATOR (config/defaults.py:180) and SPECTRAL_CONVERGENCE_LOSS (:358)
        # are off in config/promonet.py; BASELINE.json's wording of the training config names both
        self.discriminators = Discriminator(
            discriminator_state, self.generator.device, math, multi_scale_discriminator, peer_group,
            multi_resolution_discriminator)
        self.spectral_convergence = None
        if spectral_convergence_loss:
            from promonet_b200.train.losses import MultiResolutionSpectralConvergence
            self.spectral_convergence = MultiResolutionSpectralConvergence(self.generator.device, math)
        self.device = self.generator.device
        self.step_count = 0
        self.epoch = 0

    ###########################################################################
    # Data-parallel gradient exchange
    ###########################################################################

    def optimize(self, params):
        """Exchange the gradients of one module over the data-parallel ranks and take its AdamW
        step (train/core.py:256,366; torch.optim.AdamW of config/defaults.py:390-394)"""
        settings = (config.LEARNING_RATE, config.ADAM_BETAS, config.ADAM_EPS, config.WEIGHT_DECAY)
        if params.peers is not None:
            params.adamw_peer(*settings)          # one kernel over NVLink peer memory
        else:
            if self.world > 1:
                parallel.all_reduce_sum(params.grad, self.process_group)   # NCCL over NVLink
            params.adamw(*settings, grad_scale=1. / self.world)

    def close(self):
        """Give up the peer-memory optimizer before tearing the process group down: with live
        symmetric-memory mappings torch.distributed.destroy_process_group() can wait on the peers.
        Call on every rank; parameters stay readable (state_dict), the trainer does not step again."""
        torch.cuda.synchronize(self.device)
        for module in (self.generator, self.discriminators):
            module.params.release_peers()
        self.closed = True

    def broadcast_parameters(self, source=0):
        if self.world > 1:
            for params in (self.generator.params, self.discriminators.params):
                torch.distributed.broadcast(params.data, source, group=self.process_group)

    ###########################################################################
    # Step
    ###########################################################################

    def step(self, loudness, pitch, periodicity, ppg, speakers, spectral_balance_ratios,
             loudness_ratios, spectrograms, audio, update=True):
        """Batch tensors as collated by the reference (data/collate.py:43-60), on the device.
        Returns the five losses as a device tensor ordered like LOSSES (no host sync)."""
        if getattr(self, 'closed', False):
            raise RuntimeError('Trainer.step after Trainer.close()')
        batch = (loudness, pitch, periodicity, ppg, speakers, spectral_balance_ratios,
                 loudness_ratios, spectrograms, audio)
        self._discriminator_phase(batch)
        if update:
            self.optimize(self.discriminators.params)
        self._generator_phase(batch)
        if update:
            self.optimize(self.generator.params)
        return self._final_phase()

    # The step in three pieces, cut at the two optimizer steps (where the data-parallel exchange
    # happens), so that each piece can be captured in a CUDA graph (step_graphed)

    def _discriminator_phase(self, batch):
        """generator forward (:223), discriminator forward and backward (:239-255)"""
        G, D = self.generator, self.discriminators
        (loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio) = batch
        count, _, samples = audio.shape
        self.losses = torch.zeros(len(LOSSES), device=self.device)
        slot = self._slot
        G.refresh()
        # generated audio is written next to the real audio: D runs once over both halves
        both = torch.empty(2 * count, 1, samples, device=self.device)
        both[:count].copy_(audio)
        G.forward(loudness, pitch, periodicity, ppg, speakers, sbr, lr, out=both[count:])
        self.both = both
        D.refresh()
        records = D.forward(both)
        gmaps = []
        for logits, maps in zip(D.logits(records), D.feature_maps(records)):
            glogits = torch.empty_like(logits)
            ops.mse_to_target(logits[:count], 1., 1., slot('discriminator'), glogits[:count])
            ops.mse_to_target(logits[count:], 0., 1., slot('discriminator'), glogits[count:])
            gmaps.append([None] * (len(maps) - 1) + [glogits.view(maps[-1].shape)])
        D.layers.zero_grad()
        D.backward(records, gmaps, 0, 2 * count, weights=True)

    def _generator_phase(self, batch):
        """second discriminator forward with the updated discriminator (:272), generator losses
        (:277-332) and generator backward (:335-338)"""
        G, D = self.generator, self.discriminators
        spectrograms, audio = batch[7], batch[8]
        count, _, samples = audio.shape
        slot, both = self._slot, self.both
        D.refresh()
        records = D.forward(both)
        ggenerated = torch.zeros(count, 1, samples, device=self.device)
        gmaps = []
        for logits, maps in zip(D.logits(records), D.feature_maps(records)):
            # feature matching (loss.py:11-26) seeds the gradient of every generated map
            gradients = []
            for fmap in maps:
                g = torch.empty_like(fmap[count:])
                ops.l1_mean(fmap[count:], fmap[:count], config.FEATURE_MATCHING_LOSS_WEIGHT,
                            slot('feature_matching'), g)
                gradients.append(g)
            # adversarial (loss.py:43-53) adds to the gradient of the logits
            gadversarial = torch.empty_like(logits[count:])
            ops.mse_to_target(logits[count:], 1., config.ADVERSARIAL_LOSS_WEIGHT,
                              slot('adversarial'), gadversarial)
            ops.axpby(1., gadversarial.view(-1), 1., gradients[-1].view(-1))
            gmaps.append(gradients)
        D.backward(records, gmaps, count, 2 * count, weights=False, gaudio=ggenerated)
        # mel loss (:277-305)
        target_mels = ops.linear_to_mel(spectrograms)
        magnitude, spectrum = ops.stft_magnitude(both[count:].view(count, samples), 'hann', 1e-6, 0)
        gmagnitude = torch.empty_like(magnitude)
        ops.mel_loss(magnitude, target_mels, 1., slot('mel'), gmagnitude, config.MEL_LOSS_WEIGHT)
        ops.stft_magnitude_backward(
            gmagnitude, spectrum, ggenerated.view(count, samples), 'hann', 1e-6, 0, accumulate=True)
        # multi-resolution spectral convergence (:308-310)
        if self.spectral_convergence is not None:
            self.spectral_convergence(both, count, 1., slot('spectral_convergence'), ggenerated)
        G.layers.zero_grad()
        G.backward(ggenerated)
        self.generated = both[count:]

    def _final_phase(self):
        """the total generator loss (:291,323-332), on the device"""
        slot = self._slot
        ops.axpby(config.MEL_LOSS_WEIGHT, slot('mel'), 0., slot('generator'))
        ops.axpby(1., slot('feature_matching'), 1., slot('generator'))
        ops.axpby(1., slot('adversarial'), 1., slot('generator'))
        ops.axpby(1., slot('spectral_convergence'), 1., slot('generator'))
        self.step_count += 1
        return self.losses

    def _slot(self, name):
        index = LOSSES.index(name)
        return self.losses[index:index + 1]

    ###########################################################################
    # CUDA-graph replay of the step
    ###########################################################################

    def step_graphed(self, *batch):
        """Same step with its ~790 kernel launches replayed from three CUDA graphs, cut at the two
        optimizer steps, which stay eager (a fused peer-memory kernel between two device-side
        barriers, or an NCCL all-reduce and the AdamW kernel).  Shapes are fixed by the first
        call; the batch is copied into static buffers."""
        if getattr(self, 'graphs', None) is None:
            self.static_batch = [t.clone() for t in batch]
            stream = torch.cuda.Stream(self.device)
            stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(stream):
                for _ in range(2):   # warm-up: lazy tables, kernel attributes, allocator pools
                    self.step(*self.static_batch, update=False)
            torch.cuda.current_stream(self.device).wait_stream(stream)
            self.step_count -= 2
            self.graphs = []
            pool = None
            for phase in (
                lambda: self._discriminator_phase(self.static_batch),
                lambda: self._generator_phase(self.static_batch),
                lambda: self._final_phase(),
            ):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, pool=pool):
                    phase()
                pool = graph.pool()
                self.graphs.append(graph)
            self.step_count -= 1   # capturing _final_phase advanced the host counter without running
        for static, tensor in zip(self.static_batch, batch):
            static.copy_(tensor, non_blocking=True)
        self.graphs[0].replay()
        self.optimize(self.discriminators.params)
        self.graphs[1].replay()
        self.optimize(self.generator.params)
        self.graphs[2].replay()
        self.step_count += 1
        return self.losses

    ###########################################################################
    # Checkpoints (torchutil.checkpoint layout: train/core.py:426-438)
    ###########################################################################

    def save(self, directory, epoch=None):
        directory = Path(directory)
        epoch = self.epoch if epoch is None else epoch
        rank = torch.distributed.get_rank(self.process_group) if self.world > 1 else 0
        if rank == 0:
            directory.mkdir(parents=True, exist_ok=True)
        for name, module in (('generator', self.generator), ('discriminator', self.discriminators)):
            optimizer = module.params.optimizer_state()    # a collective when the step is sharded
            if rank == 0:
                torch.save(
                    {'model': module.state_dict(), 'optimizer': optimizer,
                     'step': self.step_count, 'epoch': epoch},
                    directory / f'{name}-{self.step_count:08d}.pt')

    def load(self, directory):
        """Resume from the newest generator-*.pt / discriminator-*.pt (train/core.py:70-105)"""
        directory = Path(directory)
        for name, module in (('generator', self.generator), ('discriminator', self.discriminators)):
            files = sorted(directory.glob(f'{name}-*.pt'))
            if not files:
                continue
            checkpoint = torch.load(files[-1], map_location='cpu')
            module.load_state_dict(checkpoint['model'])
            if checkpoint.get('optimizer') is not None:
                module.params.load_optimizer_state(checkpoint['optimizer'])
            self.step_count = int(checkpoint.get('step', 0))
            self.epoch = int(checkpoint.get('epoch', 0))


def train(directory, dataset='vctk', train_partition='train', valid_partition='valid',
          adapt_from=None, gpu=None, loader=None, steps=None, valid_loader=None,
          peer_optimizer=True, pitch_checkpoint=None):
    """promonet.train (promonet/train/core.py:17-24).  The reference builds its loader
    from a preprocessed dataset on disk (promonet/data, out of scope here): pass `loader`,
    an iterable of batches laid out as data/collate.py:43-60
    (text, loudness, pitch, periodicity, ppg, speakers, spectral_balance_ratios,
    loudness_ratios, spectrograms, audio, stems).  With `valid_loader` (batch-1 batches of the same
    layout) the generator is evaluated every EVALUATION_INTERVAL steps like train/core.py:387-426
    (promonet_b200.train.evaluate).  `steps` defaults to STEPS, or STEPS + ADAPTATION_STEPS when
    adapting (:111-114); `peer_optimizer=False` selects the NCCL all-reduce exchange;
    `pitch_checkpoint` is the FCNF0++ checkpoint the validation's pitch extraction uses."""
    if loader is None:
        raise ValueError(
            'promonet_b200.train needs `loader`: the dataset pipeline of the reference '
            '(promonet.data) is outside the accelerated path')
    device = torch.device('cuda', 0 if gpu is None else gpu)
    torch.cuda.set_device(device)
    trainer = Trainer(device=device, peer_optimizer=peer_optimizer)
    trainer.load(adapt_from if adapt_from is not None else directory)
    trainer.broadcast_parameters()
    if steps is None:
        # :111-114: adaptation runs ADAPTATION_STEPS past the pretraining budget
        steps = config.STEPS + (config.ADAPTATION_STEPS if adapt_from is not None else 0)
    while trainer.step_count < steps:
        sampler = getattr(loader, 'batch_sampler', None)
        if hasattr(sampler, 'set_epoch'):
            sampler.set_epoch(trainer.epoch)             # :133
        for batch in loader:
            (_, loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio, _) = batch
            if audio.shape[-1] < config.CHUNK_SIZE:   # :154
                continue
            tensors = [t.to(device, non_blocking=True) for t in (
                loudness, pitch, periodicity, ppg, speakers, sbr, lr, spectrograms, audio)]
            step = trainer.step_count
            trainer.step(*tensors)
            if valid_loader is not None and step % config.EVALUATION_INTERVAL == 0:
                from promonet_b200.train.evaluate import evaluate
                evaluate(
                    directory, step, trainer.generator, valid_loader, device.index,
                    config.DEFAULT_EVALUATION_STEPS, pitch_checkpoint=pitch_checkpoint)
            if trainer.step_count % config.CHECKPOINT_INTERVAL == 0:
                trainer.save(directory)
            if trainer.step_count >= steps:
                break
        trainer.epoch += 1                                # :382
    trainer.save(directory)
    return trainer
