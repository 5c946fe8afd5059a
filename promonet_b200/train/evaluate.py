"""In-training validation: promonet.train.evaluate (promonet/train/core.py:487-813)

For every validation item the reference synthesizes seven versions (reconstruction,
two pitch shifts, two time stretches, two loudness scalings), each through its own
batch-1 generator call, its own batch-1 preprocess.from_audio and a Metrics.update
that synchronises with the host four times.  Here the five versions that keep the
number of frames go through ONE generator launch sequence and ONE feature-extraction
pass at batch 5 (the two time stretches have their own lengths: one call each), every
edit is a kernel, and every Metrics.update is one launch with no host
synchronisation; scalars come back once at the end.

Outside the accelerated path, as everywhere in this package: the PPG of the generated
audio comes from the foreign pretrained `ppgs` model — pass `ppg_model` (audio (B, T) ->
(B, 40, F) on the device) to get the 'ppg' metric — and figures / tensorboard
(torchutil, matplotlib): the scalars are written as JSON and returned with the audio.

Data parallel (the reference is single-GPU): validation items are independent, so under an
initialised process group rank r takes items r, r + world, ... of the loader, every metric of
every condition accumulates into one (7, 12) table of doubles on the device, and ONE all-reduce
of that table (672 bytes) gives every rank the scalars of the whole validation set.
"""
import json
import math
from pathlib import Path

import torch

from promonet_b200 import _lib, config, edit, evaluate as evaluation, model, parallel, preprocess


def conditions():
    """The metric groups of train/core.py:492-507 in their order"""
    ratios = [f'{int(ratio * 100):03d}' for ratio in config.EVALUATION_RATIOS]
    return ['reconstruction'] + [
        f'{kind}-{ratio}' for kind in ('shifted', 'stretched', 'scaled') for ratio in ratios]


def inference_generator(generator, device):
    """train.evaluate receives the module being trained; the accelerated forward is the
    inference model (folded weights), so a training generator is converted through its
    state dict (weight_g / weight_v keys of the reference checkpoint)"""
    if isinstance(generator, model.Generator):
        return generator
    return model.Generator(device=device, state=generator.state_dict())


def evaluate(directory, step, generator, loader, gpu=None, evaluation_steps=None, ppg_model=None,
             process_group=None, data_parallel=True, pitch_checkpoint=None):
    """Perform model evaluation (train/core.py:487-813)

    Arguments
        directory: where `evaluation-{step:08d}.json` is written (None: nothing is written)
        step: the training step being evaluated
        generator: promonet_b200.model.Generator or promonet_b200.train.generator.Generator
        loader: iterable of batch-1 validation batches laid out as data/collate.py:43-60
        gpu: the GPU index (None = current CUDA device)
        evaluation_steps: stop after this many items (None = the whole loader)
        ppg_model: optional callable audio (B, T) -> ppg (B, 40, F)
        pitch_checkpoint: FCNF0++ checkpoint of the pitch extraction (preprocess.from_audio)
        process_group, data_parallel: under torch.distributed the items are dealt round-robin to
            the ranks of `process_group` and the sums are all-reduced (data_parallel=False: every
            rank evaluates everything on its own)

    Returns
        scalars: {f'{condition}/{metric}': value}
        waveforms: {f'{condition}/{index:02d}-audio': (1, T) device tensor} of this rank's items
    """
    if not torch.cuda.is_available():
        raise RuntimeError('promonet_b200.train.evaluate needs a CUDA device; there is no CPU path')
    device = torch.device('cuda', torch.cuda.current_device() if gpu is None else gpu)
    generator = inference_generator(generator, device)
    rank, world = 0, 1
    if data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized():
        rank = torch.distributed.get_rank(process_group)
        world = torch.distributed.get_world_size(process_group)
    table = torch.zeros(len(conditions()), _lib.METRICS_SLOTS, dtype=torch.float64, device=device)
    metrics = {
        condition: evaluation.Metrics(device, sums=table[index])
        for index, condition in enumerate(conditions())}
    ratios = config.EVALUATION_RATIOS
    waveforms = {}
    ones = lambda count: torch.ones(count, device=device)

    def analyze(audio):
        """audio (B, 1, T) -> loudness (B, 8, F), pitch (B, F), periodicity (B, F), ppg | None"""
        audio = audio[:, 0]
        loudness, pitch, periodicity = preprocess.from_audio_batch(
            audio, gpu=device.index, pitch_checkpoint=pitch_checkpoint)
        return loudness, pitch, periodicity, None if ppg_model is None else ppg_model(audio)

    # without a PPG model for the generated audio there is no pronunciation metric
    reference_ppg = lambda tensor: None if ppg_model is None else tensor

    for i, batch in enumerate(loader):
        if evaluation_steps is not None and i == evaluation_steps:        # :802-803
            break
        if not parallel.owns(i, rank, world):
            continue
        (_, loudness, pitch, periodicity, ppg, speakers, _, _, _, audio, _) = batch
        loudness, pitch, periodicity, ppg, speakers, audio = (
            item.to(device) for item in (loudness, pitch, periodicity, ppg, speakers, audio))
        if loudness.shape[0] != 1:
            raise ValueError('validation batches hold one item (train/core.py:583-593 index item 0)')
        trim = audio.shape[-1] % config.HOPSIZE                           # :560-562
        if trim > 0:
            audio = audio[..., :-trim]
        if step == 0:
            waveforms[f'original/{i:02d}-audio'] = audio[0]               # :565-566

        # The five versions with the item's own frame count, as one batch:
        # reconstruction (:572-617), pitch shifted (:623-668), loudness scaled (:746-795)
        keys = ['reconstruction']
        versions = [(loudness, pitch)]
        for ratio in ratios:
            keys.append(f'shifted-{int(100 * ratio):03d}')
            versions.append((loudness, edit.contour(pitch, scale=ratio)))              # :627
        for ratio in ratios:
            keys.append(f'scaled-{int(ratio * 100):03d}')
            versions.append((                                             # :749-750, convert.py:18-23
                edit.contour(loudness, shift=10. * math.log2(ratio)), pitch))
        count = len(versions)
        batch_loudness = torch.cat([version[0] for version in versions])
        batch_pitch = torch.cat([version[1] for version in versions])
        batch_periodicity = periodicity.expand(count, -1)
        batch_ppg = ppg.expand(count, -1, -1)
        generated = generator(
            batch_loudness, batch_pitch, batch_periodicity, batch_ppg,
            speakers.expand(count), ones(count), ones(count))
        predicted = analyze(generated)
        for index, key in enumerate(keys):
            waveforms[f'{key}/{i:02d}-audio'] = generated[index]
            select = lambda tensor: None if tensor is None else tensor[index:index + 1]
            # the reference passes the inputs in the `predicted` slots and the re-analysed
            # features in the `target` slots (:607-616); every metric is symmetric
            metrics[key].update(
                batch_loudness[index:index + 1], batch_pitch[index:index + 1], periodicity,
                reference_ppg(ppg), *(select(tensor) for tensor in predicted))

        # Time stretching (:674-740): each ratio has its own number of frames
        for ratio in ratios:
            key = f'stretched-{int(ratio * 100):03d}'
            stretched = edit.from_features(
                loudness, pitch, periodicity, ppg, time_stretch_ratio=ratio)
            generated = generator(*stretched, speakers, ones(1), ones(1))
            waveforms[f'{key}/{i:02d}-audio'] = generated[0]
            metrics[key].update(
                *stretched[:3], reference_ppg(stretched[3]), *analyze(generated))

    if world > 1:
        parallel.all_reduce_sum(table, process_group)      # every sum of every condition at once
    scalars = {}
    for condition, metric in metrics.items():                            # :806-808
        for key, value in metric().items():
            scalars[f'{condition}/{key}'] = value
    if directory is not None and rank == 0:
        directory = Path(directory)
        directory.mkdir(parents=True, exist_ok=True)
        with open(directory / f'evaluation-{step:08d}.json', 'w') as file:
            json.dump({'step': step, 'scalars': scalars}, file, indent=1)
    return scalars, waveforms
