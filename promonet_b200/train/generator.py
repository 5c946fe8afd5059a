"""Trainable HiFi-GAN generator: forward with saved activations and the
hand-sequenced backward (promonet/model/generator.py:116-197,
promonet/model/hifigan.py:63-70,97-106,141-145,198-210 in training mode, i.e.
with the weight-norm parametrisation live).  Every arithmetic step is a kernel
of libpromonet_b200; this file only orders the launches.
"""
import torch

from promonet_b200 import config
from promonet_b200.model import init
from promonet_b200.train import ops
from promonet_b200.train.layers import Layers
from promonet_b200.train.params import ParamSet

BUFFERS = ('default_previous_samples', 'ppg_threshold', 'pitch_distribution')
SLOPE = config.LRELU_SLOPE


class Generator:

    def __init__(self, state=None, device=None, math='tf32', peer_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError('promonet_b200.train needs a CUDA device (sm_100a); there is no CPU path')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        state = init.hifigan_state() if state is None else state
        self.params = ParamSet(state, self.device, BUFFERS, peer_group)
        self.ppg_threshold = float(state['ppg_threshold'])
        self.layers = Layers(self.params, math)
        conv = self.layers.conv
        self.input_conv = conv('model.input_feature_conv')
        self.speaker_conv = conv('model.input_speaker_conv')
        self.stages = []
        for i, _ in enumerate(config.HIFIGAN_UPSAMPLE_RATES):
            stage = f'model.model.{i}.model'
            blocks = []
            for j, _ in enumerate(config.HIFIGAN_RESBLOCK_KERNEL_SIZES):
                blocks.append([
                    (conv(f'{stage}.2.model.{j}.convs1.{m}'), conv(f'{stage}.2.model.{j}.convs2.{m}'))
                    for m, _ in enumerate(config.HIFIGAN_RESBLOCK_DILATION_SIZES)])
            self.stages.append((conv(f'{stage}.1'), blocks))
        self.head = conv(f'model.model.{len(config.HIFIGAN_UPSAMPLE_RATES) + 1}')
        self.layers.allocate()
        self.saved = None

    def state_dict(self):
        return self.params.state_dict()

    def load_state_dict(self, state):
        self.params.load_state_dict(state)

    def refresh(self):
        self.layers.refresh()

    ###########################################################################
    # Forward
    ###########################################################################

    @staticmethod
    def _conv_geometry(batch, channels, t, kernel, dilation):
        return ops.geometry(
            batch, channels, channels, (t, 1), (kernel, 1), 1, (dilation, 1),
            (dilation * (kernel - 1) // 2, 0))

    def forward(self, loudness, pitch, periodicity, ppg, speakers, spectral_balance_ratios,
                loudness_ratios, out=None):
        """-> generated (B, 1, 256 F); keeps what backward() needs"""
        P = self.params
        batch, frames = pitch.shape
        new = lambda *shape: torch.empty(*shape, device=self.device)
        saved = {}
        edges = P.buffers['pitch_distribution']
        features = ops.features(
            loudness, pitch, periodicity, ppg, edges, P['pitch_embedding.weight'],
            self.ppg_threshold)
        saved['bins'] = ops.pitch_bins(pitch, edges, config.FMIN, config.FMAX)
        saved['speakers'] = speakers
        gvec = ops.global_features(
            P['speaker_embedding.weight'], speakers, spectral_balance_ratios, loudness_ratios)
        speaker_geometry = ops.geometry(
            batch, config.GLOBAL_CHANNELS, config.HIFIGAN_UPSAMPLE_INITIAL_SIZE, (1, 1), (1, 1))
        speaker_bias = self.speaker_conv.apply(
            speaker_geometry, False, gvec, new(batch, config.HIFIGAN_UPSAMPLE_INITIAL_SIZE),
            bias=self.speaker_conv.bias)
        channels = config.HIFIGAN_UPSAMPLE_INITIAL_SIZE
        input_geometry = ops.geometry(
            batch, config.NUM_FEATURES, channels, (frames, 1), (7, 1), 1, 1, (3, 0))
        x = self.input_conv.apply(
            input_geometry, False, features, new(batch, channels, frames),
            bias=self.input_conv.bias, bias2=speaker_bias)
        saved.update(features=features, gvec=gvec, speaker_geometry=speaker_geometry,
                     input_geometry=input_geometry, stages=[])
        t = frames
        for (up, blocks), rate in zip(self.stages, config.HIFIGAN_UPSAMPLE_RATES):
            x_in = x
            xu = ops.conv_transpose1d(x_in, up.w.view(up.shape), up.bias, rate, SLOPE)
            channels //= 2
            t *= rate
            mrf = new(batch, channels, t)
            records = []
            for j, (kernel, block) in enumerate(zip(config.HIFIGAN_RESBLOCK_KERNEL_SIZES, blocks)):
                current = xu
                record = []
                last = len(block) - 1
                for m, ((c1, c2), dilation) in enumerate(
                        zip(block, config.HIFIGAN_RESBLOCK_DILATION_SIZES)):
                    g1 = self._conv_geometry(batch, channels, t, kernel, dilation)
                    g2 = self._conv_geometry(batch, channels, t, kernel, 1)
                    hidden = c1.apply(g1, False, current, new(batch, channels, t),
                        a_act=ops.ACT_LRELU, a_slope=SLOPE, bias=c1.bias)
                    record.append((current, hidden, g1, g2))
                    if m < last:
                        current = c2.apply(g2, False, hidden, new(batch, channels, t),
                            a_act=ops.ACT_LRELU, a_slope=SLOPE, bias=c2.bias, residual=current)
                    else:
                        # ResidualBlock.forward hifigan.py:141-145: mean over the kernels
                        c2.apply(g2, False, hidden, mrf, a_act=ops.ACT_LRELU, a_slope=SLOPE,
                            bias=c2.bias, residual=current,
                            alpha=1. / len(blocks), accumulate=j > 0)
                records.append(record)
            saved['stages'].append((x_in, xu, records))
            x = mrf
        head_geometry = ops.geometry(batch, channels, 1, (t, 1), (7, 1), 1, 1, (3, 0))
        audio = new(batch, 1, t) if out is None else out
        self.head.apply(head_geometry, False, x, audio, a_act=ops.ACT_LRELU, a_slope=SLOPE,
            out_act=ops.OUT_TANH)
        saved.update(x_last=x, audio=audio, head_geometry=head_geometry)
        self.saved = saved
        return audio

    __call__ = forward

    ###########################################################################
    # Backward
    ###########################################################################

    def backward(self, gaudio):
        """Accumulates d loss / d parameters into params.grad given d loss / d generated"""
        saved, self.saved = self.saved, None
        new = lambda like: torch.empty_like(like)
        num_kernels = len(config.HIFIGAN_RESBLOCK_KERNEL_SIZES)
        audio, x_last, geometry = saved['audio'], saved['x_last'], saved['head_geometry']
        self.head.wgrad(geometry, gaudio, x_last, bias=False, dy_companion=audio,
            dy_act=ops.ACT_TANH_MASK, x_act=ops.ACT_LRELU, x_slope=SLOPE)
        g = self.head.apply_transposed(geometry, True, gaudio, new(x_last), a_companion=audio,
            a_act=ops.ACT_TANH_MASK, mask_src=x_last, mask_slope=SLOPE)
        for (up, blocks), (x_in, xu, records), rate, kernel_size in reversed(list(zip(
                self.stages, saved['stages'], config.HIFIGAN_UPSAMPLE_RATES,
                config.HIFIGAN_UPSAMPLE_KERNEL_SIZES))):
            # every Block sees the MRF gradient / num_kernels
            third = ops.axpby(1. / num_kernels, g, 0., new(g))
            gxu = new(xu)
            for j, (block, record) in enumerate(zip(blocks, records)):
                gcurrent = third
                for m in reversed(range(len(block))):
                    c1, c2 = block[m]
                    current, hidden, g1, g2 = record[m]
                    # x_next = current + c2(lrelu(hidden)) + b2
                    c2.wgrad(g2, gcurrent, hidden, x_act=ops.ACT_LRELU, x_slope=SLOPE)
                    ghidden = c2.apply_transposed(
                        g2, True, gcurrent, new(hidden), mask_src=hidden, mask_slope=SLOPE)
                    # hidden = c1(lrelu(current)) + b1
                    c1.wgrad(g1, ghidden, current, x_act=ops.ACT_LRELU, x_slope=SLOPE)
                    if m > 0:
                        gcurrent = c1.apply_transposed(
                            g1, True, ghidden, new(current), mask_src=current,
                            mask_slope=SLOPE, residual=gcurrent)
                    else:
                        c1.apply_transposed(
                            g1, True, ghidden, gxu, mask_src=current, mask_slope=SLOPE,
                            residual=gcurrent, accumulate=j > 0)
            # xu = ConvTranspose1d(lrelu(x_in)) + b: gradients through the convolution it transposes
            batch, c_in, t_in = x_in.shape
            c_out = xu.shape[1]
            geometry = ops.geometry(
                batch, c_out, c_in, (t_in * rate, 1), (kernel_size, 1), (rate, 1), 1,
                ((kernel_size - rate) // 2, 0), size_out=(t_in, 1))
            up.wgrad(geometry, x_in, gxu, bias=False, dy_act=ops.ACT_LRELU, dy_slope=SLOPE)
            ops.channel_sum(gxu, up.gbias, accumulate=True)
            g = up.apply(geometry, False, gxu, new(x_in), mask_src=x_in, mask_slope=SLOPE)
        # input layer: x0 = conv7(features) + b + speaker projection
        P = self.params
        features, gvec = saved['features'], saved['gvec']
        self.input_conv.wgrad(saved['input_geometry'], g, features)
        gfeatures = self.input_conv.apply_transposed(
            saved['input_geometry'], True, g, new(features))
        ops.embedding_backward(
            gfeatures, saved['bins'], P.gradient('pitch_embedding.weight'),
            channel_offset=config.PPG_CHANNELS)
        batch, channels, frames = g.shape
        gspeaker = ops.row_sum(
            g, torch.empty(batch, channels, device=self.device), batch * channels, frames)
        self.speaker_conv.wgrad(saved['speaker_geometry'], gspeaker, gvec)
        ggvec = self.speaker_conv.apply_transposed(
            saved['speaker_geometry'], True, gspeaker, new(gvec))
        ops.embedding_backward(
            ggvec.view(batch, config.GLOBAL_CHANNELS, 1), saved['speakers'].view(batch, 1),
            P.gradient('speaker_embedding.weight'), channel_offset=0)
        self.layers.finish()
