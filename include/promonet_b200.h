/* promonet_b200 -- C ABI of the B200-native ProMoNet hot path.
 *
 * The reference (maxrmorrison/promonet) has no FFI: its boundary is the Python
 * nn.Module / function contract.  Each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  All
 * pointers are DEVICE pointers to contiguous row-major tensors unless the name
 * ends in `_host`; every call is asynchronous on `stream` (a cudaStream_t passed
 * as void*), allocates nothing, and returns 0 on success or a negative pmn_status
 * with a thread-local message available from pmn_last_error().
 */
#ifndef PROMONET_B200_H_
#define PROMONET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PMN_OK = 0,
    PMN_ERR_ARGUMENT = -1,  /* bad shape / null pointer / unknown tensor name */
    PMN_ERR_STATE = -2,     /* e.g. forward before finalize, missing weights */
    PMN_ERR_WORKSPACE = -3, /* workspace too small */
    PMN_ERR_CUDA = -4       /* a CUDA runtime call failed */
} pmn_status;

const char* pmn_last_error(void);
/* ABI version; bumped on any signature change */
int pmn_version(void);
/* Number of kernels launched by this library in this process (bench gpu_launches) */
int64_t pmn_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: while enabled, each launch
 * of this library is bracketed by CUDA events on its stream.  `kernel` is the
 * kernel's name (e.g. "conv1d_kernel"); read returns the summed event time. */
void pmn_profile_enable(int enabled);
void pmn_profile_reset(void);
int pmn_profile_read(const char* kernel, double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------------ */
/* Generator (HiFi-GAN) -- promonet/model/generator.py:116-135,              */
/* promonet/model/hifigan.py:63-70                                           */
/* ------------------------------------------------------------------------ */

typedef struct pmn_generator pmn_generator;

/* Precision of the dense residual-block convolutions */
typedef enum {
    PMN_MATH_FP32_SIMT = 0,  /* fp32 FMA everywhere */
    PMN_MATH_BF16X3_TC = 1   /* tcgen05, 3-product bf16 hi/lo split, fp32 accumulate */
} pmn_math;

/* config/promonet.py defaults (MODEL='hifigan'): replaces Generator.__init__
 * promonet/model/generator.py:84-114 + HiFiGAN.__init__ hifigan.py:15-61 */
int pmn_generator_create(pmn_generator** out);
void pmn_generator_destroy(pmn_generator* g);

/* Load one state_dict entry by its reference name (e.g.
 * "model.model.0.model.1.weight_g").  fp32 device tensor; copied into
 * library-owned memory.  Replaces torchutil.checkpoint.load at
 * promonet/synthesize/core.py:245. */
int pmn_generator_set_tensor(
    pmn_generator* g, const char* name, const float* data,
    const int64_t* shape, int ndim, void* stream);

/* Fold weight norm (w = g * v / ||v||, hifigan.py:100-106,167-183; core.py:43)
 * and pack weights for the kernels.  Must be called after all tensors are set. */
int pmn_generator_finalize(pmn_generator* g, int math, void* stream);

/* Tensor-core math only: which residual blocks (Block.forward, hifigan.py:198-210) run
 * their three c1 -> c2 pairs as fused pmn_conv_pair_tc-style launches.  Bit
 * 3 * stage + block (block 0 / 1 / 2 = kernel 3 / 7 / 11); stage 0 (C = 256) is never
 * fused.  The default (0x248: the k = 3 block of stages 1 - 3) is the measured-fastest
 * selection.  Results agree within fp32 rounding for every mask, bit for bit where both
 * paths run bf16 x 3 products in the same order. */
int pmn_generator_set_pair_mask(pmn_generator* g, unsigned mask);

/* Tensor-core math only: run the unfused residual blocks of the C = 128 stage (a third of the
 * generator's FLOPs) with "fp16 + 2 x fp8" operands (pmn_conv1d_tc_f8) instead of bf16 x 3: fewer
 * tensor cycles, output within the same 1e-4 bar (measured 0.8 - 1.6e-5 at B = 32 x 430 frames
 * against 0.8e-5).  Call after pmn_generator_finalize. */
int pmn_generator_set_f8(pmn_generator* g, int enabled);

size_t pmn_generator_workspace_bytes(const pmn_generator* g, int batch, int frames);

/* Generator.forward generator.py:116-135.
 *   loudness     (B, loudness_rows, F) fp32, loudness_rows = 8 or 513 (dB)
 *   pitch        (B, F) fp32 Hz;  periodicity (B, F) fp32
 *   ppg          (B, 40, F) fp32
 *   speakers     (B,) int64;  spectral_balance_ratios, loudness_ratios (B,) fp32
 *   audio        (B, 1, 256 F) fp32 out */
int pmn_generator_forward(
    pmn_generator* g,
    const float* loudness, int loudness_rows,
    const float* pitch, const float* periodicity, const float* ppg,
    const int64_t* speakers,
    const float* spectral_balance_ratios, const float* loudness_ratios,
    float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies inputs H2D, runs the
 * forward, copies audio D2H, all on `stream`; `staging` is a device buffer of
 * pmn_generator_staging_bytes().  This is what promonet.synthesize.from_features
 * (promonet/synthesize/core.py:18-59) maps onto. */
size_t pmn_generator_staging_bytes(int batch, int frames, int loudness_rows);
int pmn_generator_forward_host(
    pmn_generator* g,
    const float* loudness_host, int loudness_rows,
    const float* pitch_host, const float* periodicity_host, const float* ppg_host,
    const int64_t* speakers_host,
    const float* spectral_balance_ratios_host, const float* loudness_ratios_host,
    float* audio_host, int batch, int frames,
    void* staging, size_t staging_bytes,
    void* workspace, size_t workspace_bytes, void* stream);

/* Intermediate of the forward, for unit parity: prepare_features
 * generator.py:137-197 -> (B, 113, F) */
int pmn_generator_features(
    pmn_generator* g,
    const float* loudness, int loudness_rows,
    const float* pitch, const float* periodicity, const float* ppg,
    float* features, int batch, int frames, void* stream);

/* ------------------------------------------------------------------------ */
/* Generator (FARGAN, config/fargan.py) -- promonet/model/fargan.py:21-131,  */
/* promonet/model/generator.py:116-135 with MODEL='fargan'                   */
/* ------------------------------------------------------------------------ */

typedef struct pmn_fargan pmn_fargan;
int pmn_fargan_create(pmn_fargan** out);
void pmn_fargan_destroy(pmn_fargan* g);
/* state_dict entries by reference name, e.g.
 * "model.subframe_network.gru1.weight_ih", "model.conditioning_network.0.weight" */
int pmn_fargan_set_tensor(
    pmn_fargan* g, const char* name, const float* data, const int64_t* shape, int ndim, void* stream);
int pmn_fargan_finalize(pmn_fargan* g, void* stream);
size_t pmn_fargan_workspace_bytes(const pmn_fargan* g, int batch, int frames);
/* Same inputs as pmn_generator_forward plus previous_samples (B, 512) fp32 or NULL
 * (zeros: generator.py default_previous_samples); audio (B, 1, 256 F).  Inference
 * semantics: no additive noise (fargan.py:396-403 is training-only). */
int pmn_fargan_forward(
    pmn_fargan* g,
    const float* loudness, int loudness_rows,
    const float* pitch, const float* periodicity, const float* ppg,
    const int64_t* speakers,
    const float* spectral_balance_ratios, const float* loudness_ratios,
    const float* previous_samples,
    float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* Feature extraction -- promonet/preprocess/core.py:17-126                  */
/* ------------------------------------------------------------------------ */

/* STFT-derived features of `batch` equal-length utterances audio (B, T) fp32 at
 * 22 050 Hz; F = T / 256 frames (reflect padding 384, hann 1024, hop 256).
 * Any of the three outputs may be NULL:
 *   magnitude (B, 513, F)  sqrt(re^2 + im^2 + 1e-6)     preprocess/spectrogram.py:35-52
 *   mels      (B, 80, F)   max(log(mel_basis @ magnitude), mel_floor)  spectrogram.py:111-135
 *             (mel_floor = -INFINITY when no dynamic-range threshold is configured)
 *   loudness  (B, bands, F), or (B, 513, F) when bands <= 0: A-weighted dB
 *             promonet/preprocess/loudness.py:17-55 (top_db 80 per utterance, floor -100)
 * workspace (pmn_spectral_workspace_bytes) is needed for loudness only. */
size_t pmn_spectral_workspace_bytes(int batch, int samples);
int pmn_spectral_features(
    const float* audio, int batch, int samples,
    float* magnitude, float* mels, float mel_floor, float* loudness, int loudness_bands,
    void* workspace, size_t workspace_bytes, void* stream);

/* linear_to_mel (preprocess/spectrogram.py:111-135) of an existing magnitude
 * spectrogram (B, 513, F) -> (B, 80, F) */
int pmn_linear_to_mel(
    const float* magnitude, float* mels, float mel_floor, int batch, int frames, void* stream);

/* torbi.from_probabilities (call witnessed at promonet/preprocess/harmonics.py:270-276):
 *   observation (B, T, S), batch_frames (B) int32 valid lengths or NULL,
 *   transition (S, S) [row i -> column j], initial (S); log_probs = 0 takes logs inside
 *   indices (B, T) int32 out; ties resolve to the lowest index */
size_t pmn_viterbi_workspace_bytes(int batch, int frames, int states);
int pmn_viterbi_decode(
    const float* observation, const int32_t* batch_frames, const float* transition,
    const float* initial, int log_probs, int32_t* indices, int batch, int frames, int states,
    void* workspace, size_t workspace_bytes, void* stream);

/* penn.from_audio(audio, sample_rate, hopsize, fmin, fmax, center='half-hop',
 * decoder='viterbi') as called at promonet/preprocess/core.py:71-81: FCNF0++
 * network + Viterbi + local expected value.  Tensor names: layers.{0..5}.conv.{weight,bias},
 * layers.{0..5}.norm.{weight,bias}, layers.6.{weight,bias}. */
typedef struct pmn_pitch pmn_pitch;
int pmn_pitch_create(pmn_pitch** out);
void pmn_pitch_destroy(pmn_pitch* p);
int pmn_pitch_set_tensor(
    pmn_pitch* p, const char* name, const float* data, const int64_t* shape, int ndim, void* stream);
/* math: pmn_math (tensor cores run blocks 1..5; block 0 and the head stay fp32 FMA) */
int pmn_pitch_finalize(pmn_pitch* p, int math, void* stream);
int pmn_pitch_frames(int samples, int sample_rate, double hopsize_seconds);
size_t pmn_pitch_workspace_bytes(
    int batch, int samples, int sample_rate, double hopsize_seconds, int frame_batch);
/*   audio (B, T) fp32 at sample_rate; transition (1440, 1440), initial (1440) probabilities
 *   pitch, periodicity (B, F); optional logits_out (B, F, 1440) masked logits and
 *   bins_out (B, F) int32 decoded bins; frame_batch = frames per network pass */
int pmn_pitch_forward(
    pmn_pitch* p, const float* audio, int batch, int samples, int sample_rate,
    double hopsize_seconds, float fmin, float fmax, const float* transition, const float* initial,
    float* pitch, float* periodicity, float* logits_out, int32_t* bins_out, int frame_batch,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* Operator-level entry points (unit-test granularity)                       */
/* ------------------------------------------------------------------------ */

/* w[d0] = g[d0] * v[d0] / ||v[d0]||_2 over the trailing `inner` elements
 * (torch.nn.utils.weight_norm dim=0; promonet/model/core.py:43-45) */
int pmn_weight_norm_fold(
    const float* v, const float* g, float* w, int dim0, int inner, void* stream);

/* Conv1d weight (C_out, C_in, K) -> packed (C_in, K, C_out) for pmn_conv1d */
int pmn_pack_conv1d_weight(
    const float* w, float* packed, int c_out, int c_in, int k, void* stream);

/* Fused dilated Conv1d (torch.nn.functional.conv1d semantics, stride 1):
 *   y[b,o,t] = act_out( bias[o] + bias2[b,o] + residual[b,o,t]
 *              + sum_{c,j} w[o,c,j] * lrelu_in(x[b,c,t + j*dilation - padding]) )
 * hifigan.py:198-210 is c1/c2 of Block.forward with in_slope=0.1.
 *   packed_weight from pmn_pack_conv1d_weight; bias/bias2/residual may be NULL
 *   in_slope: LeakyReLU slope applied to x (1.0 = none)
 *   out_act: 0 none, 1 tanh, 2 relu
 *   accum / accum_mode: optional second output (B, C_out, T_out):
 *      0 unused, 1 accum = scale*y, 2 accum += scale*y   (MRF mean, hifigan.py:141-145)
 *   out may be NULL when only accum is wanted; out may alias residual. */
int pmn_conv1d(
    const float* x, const float* packed_weight, const float* bias,
    const float* bias2, const float* residual, float* out,
    float* accum, int accum_mode, float accum_scale,
    int batch, int c_in, int c_out, int t_in, int t_out,
    int k, int dilation, int padding, float in_slope, int out_act,
    void* stream);

/* The same convolution ("same" padding, odd k, C_in = C_out in {32, 64, 128, 256})
 * on the tcgen05 tensor cores with a 3-product bf16 hi/lo split and fp32
 * accumulation in TMEM (promonet_b200/csrc/conv1d_tc.cu).  Takes plain fp32
 * tensors: x is converted to hi/lo planes of lrelu(x, in_slope) and `weight`
 * (C_out, C_in, K), already weight-norm folded, is packed into slabs inside
 * `workspace` (pmn_conv1d_tc_workspace_bytes), so this entry point measures
 * parity, not speed.  planes_out, if not NULL, receives hi + lo of the bf16
 * planes of lrelu(y, out_slope) that the next convolution would consume. */
size_t pmn_conv1d_tc_workspace_bytes(int batch, int channels, int t_len, int k);
/* Profiling aid: while `counters` (device, 148 x 10 x 4 int64) is non-NULL every
 * tensor-core conv launch stores per-warp-role cycle counters there
 * ([0] total, [1..3] cycles spent waiting on the pipeline barriers). */
void pmn_debug_tc_counters(void* counters);
int pmn_conv1d_tc(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation,
    float in_slope, float out_slope,
    void* workspace, size_t workspace_bytes, void* stream);

/* The same entry with "fp16 + 2 x fp8" operands (channels in {128, 256}): x and the weights are
 * converted to fp16 of the scaled value plus two e4m3 sections (the value and its fp16 rounding
 * error), products x_m w_m [kind::f16] + two fp8 correction products [kind::f8f6f4] accumulate in
 * one TMEM accumulator at two thirds of the bf16 x 3 tensor cycles (conv1d_tc.cuh).  planes_out
 * receives what the next convolution's operand holds of lrelu(y, out_slope): fp16 part + low part. */
int pmn_conv1d_tc_f8(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation,
    float in_slope, float out_slope,
    void* workspace, size_t workspace_bytes, void* stream);

/* One residual pair of Block.forward (promonet/model/hifigan.py:198-210),
 *     y = x + c2(lrelu(c1(lrelu(x), dilation)))      c1, c2: Conv1d(C, C, k), "same",
 * fused in one tcgen05 kernel that reads the fp32 stream once and writes it once
 * (promonet_b200/csrc/conv_pair_tc.cu); C in {32, 64, 128}, odd k <= 11,
 * (k - 1) * dilation <= 50.  weight1 / weight2 are folded fp32 (C, C, K) and are
 * packed into `workspace` (pmn_conv_pair_tc_workspace_bytes) on every call: a
 * parity entry point.  out and accum (same modes as pmn_conv1d) must not alias x.
 * Bit-identical to pmn_conv1d_tc applied twice. */
size_t pmn_conv_pair_tc_workspace_bytes(int channels, int k);
/* Profiling aid: while `counters` (device, 148 x 10 x 4 int64) is non-NULL every fused-pair
 * launch stores per-warp-role cycle counters there; `variant` >= 0 selects a kernel variant
 * (converter warps / mid-image buffers) for experiments, -1 the default.  Results do not
 * depend on the variant. */
void pmn_debug_pair_tc(void* counters, int variant);
int pmn_conv_pair_tc(
    const float* x, const float* weight1, const float* bias1, const float* weight2,
    const float* bias2, float* out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float slope,
    void* workspace, size_t workspace_bytes, void* stream);

/* General form: C_in -> C_out in {(256,256), (128,128), (64,64), (32,32), (256,32),
 * (32,128), (128,256)}, k <= 32, valid != 0 for no padding (T_out = T - (k-1) d),
 * relu != 0 for max(y, 0): the penn FCNF0++ blocks (promonet/preprocess/core.py:71). */
int pmn_conv1d_tc_general(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int c_in, int c_out, int t_len, int k, int dilation, int valid, int relu,
    float in_slope, float out_slope, void* workspace, size_t workspace_bytes, void* stream);

/* LeakyReLU + ConvTranspose1d with kernel = 2*stride, padding = stride/2
 * (hifigan.py:97-106; PyTorch weight layout (C_in, C_out, K)), T_out = stride*T_in */
int pmn_conv_transpose1d(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope,
    void* stream);

/* The same transposed convolution on the tensor cores (3-tap phase-major
 * formulation, bf16 x 3; (c_in, stride) in {(512, 8), (256, 8), (128, 2), (64, 2)},
 * c_out = c_in / 2).  fp32 in and out; planes and weight slabs are built inside
 * `workspace` (pmn_conv_transpose1d_tc_workspace_bytes): parity entry point. */
size_t pmn_conv_transpose1d_tc_workspace_bytes(int batch, int c_in, int t_in, int stride);
int pmn_conv_transpose1d_tc(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ */
/* Training step -- promonet/train/core.py:183-369 (operator level; the      */
/* step itself is sequenced by promonet_b200/train/)                         */
/* ------------------------------------------------------------------------ */

/* Geometry of a FORWARD 2-D convolution (torch.nn.functional.conv2d semantics,
 * zero padding).  A Conv1d over (B, C, T) is the case w_in = w_out = kw = 1 with
 * time on the H axis: hifigan.py:167-183 (dilated), discriminator.py:67-72
 * ((5,1) kernels, stride (3,1) over (B, C, T/p, p)), :160-170 ((3,9), stride (1,2)). */
typedef struct {
    int batch, c_in, c_out;
    int h_in, w_in, h_out, w_out;
    int kh, kw, sh, sw, dh, dw, ph, pw;
    /* Optional element strides of the gathered tensor of pmn_conv_gemm[_tc] (0 = contiguous
     * (B, C, H, W)): channel c, position (h, w) of item b is read at
     * b * batch_stride + c * channel_stride + (h * W + w) * position_stride.
     * With channel_stride = 1 and position_stride = hop a signal (B, T) is read as its overlapping
     * frames (B, n_fft, frames) without materialising them: an STFT is then a 1 x 1 convolution
     * with the windowed DFT basis (promonet/train/loss.py:61-80). */
    int channel_stride, position_stride, batch_stride;
} pmn_conv_geometry;

/* Activation fused into an operand load */
typedef enum {
    PMN_ACT_NONE = 0,
    PMN_ACT_LRELU = 1,       /* lrelu(value, slope) */
    PMN_ACT_LRELU_MASK = 2,  /* value * (companion > 0 ? 1 : slope): backward of an output LeakyReLU,
                                companion = the activated output */
    PMN_ACT_TANH_MASK = 3    /* value * (1 - companion^2): backward of an output tanh */
} pmn_operand_act;
typedef enum { PMN_OUT_NONE = 0, PMN_OUT_LRELU = 1, PMN_OUT_TANH = 2 } pmn_output_act;

/* Implicit-GEMM convolution.
 *   transposed = 0 (forward): a = x (B, c_in, h_in, w_in), wmat = weight (c_out, c_in, kh, kw),
 *       out (B, c_out, h_out, w_out)
 *   transposed = 1 (data gradient; also the forward of ConvTranspose, hifigan.py:100-106):
 *       a = dy (B, c_out, h_out, w_out), wmat = weight transposed to (c_in, c_out, kh, kw)
 *       (pmn_transpose_weight), out = dx (B, c_in, h_in, w_in)
 *   out = [out +] alpha * (mask(out_act(conv + bias[n] + bias2[b, n])) + residual)
 *   mask: value *= (mask_src > 0 ? 1 : mask_slope), the backward of an INPUT LeakyReLU
 *   bias, bias2, mask_src, residual, a_companion may be NULL; residual and mask_src must not
 *   alias out (use accumulate). */
int pmn_conv_gemm(
    const pmn_conv_geometry* geometry, int transposed,
    const float* a, const float* a_companion, int a_act, float a_slope,
    const float* wmat, const float* bias, const float* bias2,
    int out_act, float out_slope, const float* mask_src, float mask_slope,
    const float* residual, float alpha, int accumulate, float* out, void* stream);

/* pmn_conv_gemm on the tcgen05 tensor cores: tf32 operands (10-bit mantissa), fp32
 * accumulation in TMEM (promonet_b200/csrc/train_conv_tc.cu).  Same arguments, except that the
 * weight is the packing made by pmn_pack_weight_taps (16-byte aligned; the shared-memory
 * images of the kernel's weight tiles, copied with cp.async.bulk):
 *   transposed = 0: pack(weight (c_out, c_in, kh, kw), transposed = 0): rows c_out, reduce over c_in
 *   transposed = 1: pack(weight (c_out, c_in, kh, kw), transposed = 1): rows c_in, reduce over c_out
 * of pmn_conv_tc_packed_floats(rows, reduce, taps) floats.  out_act = PMN_OUT_TANH is not built. */
int pmn_conv_gemm_tc(
    const pmn_conv_geometry* geometry, int transposed,
    const float* a, const float* a_companion, int a_act, float a_slope,
    const float* wpacked, const float* bias, const float* bias2,
    int out_act, float out_slope, const float* mask_src, float mask_slope,
    const float* residual, float alpha, int accumulate, float* out, void* stream);
int pmn_conv_tc_channel_pad(int channels);
size_t pmn_conv_tc_packed_floats(int rows, int reduce, int taps);
/* Profiling aid: while `counters` (device, 8 x int64) is non-NULL, CTA 0 of every tensor-core
 * training convolution stores cycle counters there: [0] K loop, [1] waiting for the stage, [2] operand
 * loads, [3] shared-memory stores + proxy fence, [4] CTA barrier, [5] MMA issue, [6] K steps, [7] until
 * the accumulator is complete */
void pmn_debug_train_tc_counters(void* counters);
/* w (d0, d1, taps) -> [row tile][tap][32-channel block][k / 4][row in tile][4], rounded to tf32,
 * rows = d0 (transposed = 0) or d1 (transposed = 1), padding written as zeros */
int pmn_pack_weight_taps(
    const float* w, float* out, int d0, int d1, int taps, int transposed, void* stream);

/* Weight (and bias) gradient, ACCUMULATED atomically into gw (c_out, c_in, kh, kw) and
 * gbias (c_out, may be NULL):  gw[n, c, i, j] += sum_{b, p} act(dy)[b, n, p] act(x)[b, c, in(p, i, j)].
 * The weight gradient of a ConvTranspose (C_in, C_out, k) is this call on the geometry of the
 * convolution it transposes, with the roles of input and output gradient exchanged. */
int pmn_conv_wgrad(
    const pmn_conv_geometry* geometry,
    const float* dy, const float* dy_companion, int dy_act, float dy_slope,
    const float* x, const float* x_companion, int x_act, float x_slope,
    float* gw, float* gbias, void* stream);

/* pmn_conv_wgrad on the tcgen05 tensor cores (tf32 operands, fp32 accumulation; the bias
 * gradient rides along as a row of ones).  Built for the activation pairs of the training
 * step: (dy_act, x_act) in {(NONE, NONE), (NONE, LRELU), (LRELU_MASK, NONE), (LRELU, NONE)}. */
int pmn_conv_wgrad_tc(
    const pmn_conv_geometry* geometry,
    const float* dy, const float* dy_companion, int dy_act, float dy_slope,
    const float* x, const float* x_companion, int x_act, float x_slope,
    float* gw, float* gbias, void* stream);

/* Everything a module's convolutions need after an optimizer step, in two launches: for each
 * entry of the DEVICE table, w = g v / ||v|| when g is not NULL (else w is unused and the weight
 * is read from v), then packed / packed_t = pmn_pack_weight_taps(weight, transposed = 0 / 1) and
 * wt = pmn_transpose_weight(weight); any of packed, packed_t, wt may be NULL. */
typedef struct {
    const float* v;      /* weight_v, or the plain weight when g is NULL: (dim0, dim1, taps) */
    const float* g;      /* weight_g (dim0) or NULL */
    float* w;            /* folded weight out (when g is not NULL) */
    float* packed;
    float* packed_t;
    float* wt;
    float* dense;        /* (dim0, dim1, taps) dense form of a grouped weight, or NULL */
    int dim0, dim1, taps;
    int groups;          /* > 1: v / w are (dim0, dim1 / groups, taps) (torch Conv1d groups,
                            discriminator.py:218-224); the packings are block-diagonal */
} pmn_weight_desc;
int pmn_prepare_weights(const pmn_weight_desc* table, int layers, int max_dim0, void* stream);
/* gw (dim0, dim1 / groups, taps) = diagonal blocks of a dense weight gradient (dim0, dim1, taps) */
int pmn_extract_grouped(
    const float* dense, float* gw, int dim0, int dim1, int taps, int groups, void* stream);

/* (dim0, dim1, taps) -> (dim1, dim0, taps) */
int pmn_transpose_weight(
    const float* w, float* wt, int dim0, int dim1, int taps, void* stream);

/* Backward of pmn_weight_norm_fold (model/core.py:43-45): gv (dim0, inner), gg (dim0) written */
int pmn_weight_norm_backward(
    const float* v, const float* g, const float* gw, float* gv, float* gg, int dim0, int inner,
    void* stream);

/* The same for every weight-normed convolution of a module in one launch: a DEVICE table with one
 * entry per layer (train/core.py:255,338 call backward() once per step; autograd runs this per
 * parameter).  max_dim0 = the largest dim0 in the table. */
typedef struct {
    const float* v;      /* weight_v (dim0, inner) */
    const float* g;      /* weight_g (dim0) */
    const float* gw;     /* gradient of the folded weight (dim0, inner) */
    float* gv;           /* out: gradient of weight_v */
    float* gg;           /* out: gradient of weight_g */
    int dim0, inner;
} pmn_weight_norm_desc;
int pmn_weight_norm_backward_table(
    const pmn_weight_norm_desc* table, int layers, int max_dim0, void* stream);

/* torch.nn.functional.pad(x, (left, right), 'reflect') over rows of length t_in
 * (discriminator.py:78-81) and its adjoint */
int pmn_reflect_pad(
    const float* x, float* out, int rows, int t_in, int left, int right, void* stream);
int pmn_reflect_pad_backward(
    const float* gout, float* gx, int rows, int t_in, int left, int right, int accumulate,
    void* stream);

/* y = a x + b y (x may be NULL: y = b y) */
int pmn_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream);

/* LSGAN terms, promonet/train/loss.py:29-53: *loss += weight * mean((x - target)^2),
 * grad = d/dx (written; may be NULL) */
int pmn_mse_to_target(
    const float* x, int64_t n, float target, float weight, float* loss, float* grad, void* stream);
/* Feature matching / L1, loss.py:11-26: *loss += weight * mean|fake - real|,
 * gfake (+)= d/dfake (may be NULL) */
int pmn_l1_mean(
    const float* fake, const float* real, int64_t n, float weight, float* loss, float* gfake,
    int accumulate, void* stream);

/* torch.optim.AdamW step over a flat parameter buffer (train/core.py:63-64,256,366;
 * config/defaults.py:390-394); grad is multiplied by grad_scale first (1 / world size).  The
 * step count t of the bias corrections is `step`, or *step_device (device float) when that is
 * not NULL -- so that a captured CUDA graph of the training step stays valid as t advances. */
int pmn_adamw(
    float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
    float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
    const float* step_device, void* stream);

/* The data-parallel exchange of the training step (the all-reduce a DDP wrapper would put after
 * promonet/train/core.py:255 and :338) fused with the optimizer, over NVLink peer memory:
 * grad_peers[r] / param_peers[r] (HOST arrays of `world` DEVICE pointers, e.g. the buffer_ptrs of
 * a symmetric-memory allocation) are rank r's flat gradient / parameter buffers.  This rank
 * averages elements [begin, end) of all gradient buffers, takes the AdamW step on them with its
 * own moments (only that slice of exp_avg / exp_avg_sq is maintained: ZeRO-1) and writes the new
 * parameters into every rank's parameter buffer: reduce-scatter + AdamW + all-gather in one
 * kernel, no NCCL on the data path.  The caller brackets it with cross-rank barriers (gradients
 * complete before, parameter writes landed after).  begin, end multiples of 4; world <= 8. */
int pmn_adamw_peer(
    const float* const* grad_peers, float* const* param_peers, int world, int rank, float* exp_avg,
    float* exp_avg_sq, int64_t begin, int64_t end, float lr, float beta1, float beta2, float eps,
    float weight_decay, int step, const float* step_device, void* stream);

/* out[r] (+)= sum_c x[r, c] */
int pmn_row_sum(const float* x, float* out, int rows, int cols, int accumulate, void* stream);

/* Generator.prepare_features (generator.py:137-197) with explicit tables, for training where
 * pitch_embedding is a parameter: -> features (B, 113, F) */
int pmn_features(
    const float* loudness, int loudness_rows, const float* pitch, const float* periodicity,
    const float* ppg, const float* pitch_distribution, const float* pitch_embedding,
    float ppg_threshold, float* features, int batch, int frames, void* stream);
/* bins (B, F) int64 = clip(searchsorted(edges, clip(pitch, fmin, fmax)), 0, n - 1), generator.py:153-157 */
int pmn_pitch_bins(
    const float* pitch, const float* edges, int64_t* bins, int n, int num_edges,
    float fmin, float fmax, void* stream);
/* gtable[index[b, f], e] += gout[b, channel_offset + e, f]: backward of an embedding lookup
 * whose rows were written to channels [channel_offset, channel_offset + channels) of
 * gout (B, out_channels, F) */
int pmn_embedding_backward(
    const float* gout, const int64_t* index, float* gtable, int batch, int channels, int frames,
    int rows, int out_channels, int channel_offset, void* stream);
/* prepare_global_features (generator.py:49-70): out (B, speaker_channels + 2) */
int pmn_global_features(
    const float* speaker_embedding, const int64_t* speakers, const float* spectral_balance_ratios,
    const float* loudness_ratios, float* out, int batch, int speaker_channels, int num_speakers,
    void* stream);

/* Differentiable STFT magnitude over audio (B, T): reflect pad 384, 1024-point frames, hop 256.
 *   window_kind 0 = periodic hann (preprocess/spectrogram.py:36-52, eps 1e-6),
 *               1 = rectangular (discriminator.py:175-195, eps 0)
 *   layout 0: magnitude (B, 513, F); 1: (B, F, 513) (the CMB discriminator's (B, 1, F, 513));
 *          2: the five CMB bands [0,51) [51,128) [128,256) [256,384) [384,513) of layout 1 stored
 *             one after the other, each a contiguous (B, F, width) tensor (discriminator.py:195)
 *   spectrum (B, F, 513, 2) complex, kept for the backward; either output may be NULL */
int pmn_stft_magnitude(
    const float* audio, int batch, int samples, int window_kind, float eps, int layout,
    float* spectrum, float* magnitude, void* stream);
int pmn_stft_magnitude_backward(
    const float* gmagnitude, const float* spectrum, int batch, int samples, int window_kind,
    float eps, int layout, float* gaudio, int accumulate, void* stream);
/* Mel loss (train/core.py:277-305): L = mean|log(mel_basis @ magnitude) - target|;
 * *loss += loss_weight * L, gmagnitude (B, 513, F) = grad_weight * dL/dmagnitude (may be NULL);
 * magnitude (B, 513, F), target_mels (B, 80, F) */
int pmn_mel_loss(
    const float* magnitude, const float* target_mels, int batch, int frames, float loss_weight,
    float grad_weight, float* loss, float* gmagnitude, void* stream);

/* out[c] (+)= sum_{b, i} x[b, c, i] over x (batch, channels, inner) */
int pmn_channel_sum(
    const float* x, float* out, int batch, int channels, int inner, int accumulate, void* stream);
/* dst[r, dst_offset + j] (+)= src[r, src_offset + j] for j < cols: torch.cat / split along the
 * last axis (discriminator.py:204) */
int pmn_copy_columns(
    const float* src, int src_width, int src_offset, float* dst, int dst_width, int dst_offset,
    int64_t rows, int cols, int accumulate, void* stream);

/* Multi-resolution spectral convergence, promonet/train/loss.py:61-150.  The STFT of one
 * resolution is a 1 x 1 pmn_conv_gemm[_tc] over the reflect-padded signal read as overlapping
 * frames (pmn_conv_geometry strides) with this weight: (2 bins, n_fft), bins = n_fft / 2 + 1,
 * rows [0, bins) = hann[n] cos(2 pi k n / n_fft), rows [bins, 2 bins) = -hann[n] sin(...). */
int pmn_dft_basis(float* out, int n_fft, void* stream);
/* spec (2 B, 2 bins, frames): real then imaginary rows; items [0, B) the target y, [B, 2 B) the
 * prediction x.  With s = sqrt(clamp(|X|, 1e-7)): sums (2 floats, overwritten) = (sum |s_y - s_x|,
 * sum s_y); *loss += weight sums[0] / sums[1]; gspec (B, 2 bins, frames) = weight times the
 * gradient of that ratio with respect to the prediction's spectrum (may be NULL). */
int pmn_spectral_convergence(
    const float* spec, int batch, int bins, int frames, float weight, float* sums, float* loss,
    float* gspec, void* stream);
/* gsignal[b, f hop + n] += gframes[b, n, f]: the adjoint of reading (B, samples) as frames */
int pmn_frame_overlap_add(
    const float* gframes, float* gsignal, int batch, int n_fft, int frames, int hop, int samples,
    void* stream);

/* promonet.edit.grid.sample (promonet/edit/grid.py:12-43): 1-D grid sampling of
 * sequence (items, channels, t_in) at grid (t_out) positions (frames, fractional) ->
 * out (items, channels, t_out); nearest = 0: linear with the final frame replicated, 1: nearest.
 * renormalize != 0 applies softmax(log(p + 1e-8)) over channels afterwards (the PPG resampling of
 * promonet/preprocess/core.py:97-103, the step before synthesize.from_features on the file path). */
int pmn_grid_sample(
    const float* sequence, const float* grid, float* out, int items, int channels, int t_in,
    int t_out, int nearest, int renormalize, void* stream);

/* ---- Multi-resolution spectrogram discriminator front end ----
 * DiscriminatorR.spectrogram, promonet/model/discriminator.py:127-141 (MULTI_RESOLUTION_DISCRIMINATOR,
 * off in config/promonet.py).
 * pmn_dft_basis_rect: the weight of the STFT-as-1-x-1-convolution (see pmn_dft_basis) for
 * torch.stft(window=None, win_length <= n_fft): (2 bins, n_fft), a rectangular window of win_length
 * samples centred in the frame. */
int pmn_dft_basis_rect(float* out, int n_fft, int win_length, void* stream);
/* spec (items, 2 bins, frames), real rows then imaginary rows -> magnitude (items, bins, frames)
 * = sqrt(re^2 + im^2) (no epsilon: torch.norm, :141); backward: gspec (items, 2 bins, frames) =
 * gmagnitude (re, im) / |X|, zero where |X| = 0 */
int pmn_complex_magnitude(
    const float* spec, float* magnitude, int items, int bins, int frames, void* stream);
int pmn_complex_magnitude_backward(
    const float* gmagnitude, const float* spec, float* gspec, int items, int bins, int frames,
    void* stream);

/* ---- In-training validation ----
 * promonet.evaluate.Metrics.update (promonet/evaluate/metrics.py:38-61) in one pass over
 * `items` utterances of `frames` frames: adds into sums[PMN_METRICS_SLOTS] (device doubles)
 *   [0, 1]  loudness: squared error of the row means, count         (metrics.py:185-204)
 *   [2, 3]  the same over frames where both means exceed loudness_threshold ("loud")
 *   [4, 5]  ... over the other frames ("quiet")
 *   [6, 7]  periodicity: squared error, count                        (metrics.py:20,55)
 *   [8, 9]  |log2 predicted_pitch - log2 target_pitch| over frames where both periodicities
 *           exceed voicing_threshold (penn.voicing.threshold), count (metrics.py:249-261)
 *   [10, 11] Jensen-Shannon distance of the sparsified PPGs (ppgs.sparsify 'percentile'
 *           ppg_threshold, then ppgs.distance reduction='sum'), frames (metrics.py:287-312);
 *           similarity: optional (40, 40) phoneme-similarity matrix, already raised to
 *           ppgs.SIMILARITY_EXPONENT, applied as p <- S^T p (ppgs is un-vendored: restated)
 * loudness (items, bands, frames) — predicted and target may have different row counts, each
 * is averaged over its own (metrics.py:192-193); pitch, periodicity (items, frames);
 * ppg (items, 40, frames).  Any predicted/target pair may be NULL (its slots are left alone).
 * RMSE = sqrt(sums[0] / sums[1]) etc. are formed on the host by the caller. */
#define PMN_METRICS_SLOTS 12
int pmn_metrics_update(
    const float* predicted_loudness, int predicted_bands,
    const float* target_loudness, int target_bands,
    const float* predicted_pitch, const float* target_pitch,
    const float* predicted_periodicity, const float* target_periodicity,
    const float* predicted_ppg, const float* target_ppg, int ppg_channels,
    const float* similarity, int items, int frames,
    float loudness_threshold, float voicing_threshold, float ppg_threshold,
    double* sums, void* stream);
/* promonet.edit.from_features for one contour (promonet/edit/core.py:113-128):
 * sequence (items, t_in) resampled at grid (t_out) like pmn_grid_sample (grid == NULL: no
 * resampling, t_out == t_in), in the log2 domain when log2_domain != 0 (pitch), then
 * out = scale * value + shift, clipped to [lo, hi] when lo < hi (pitch shift: FMIN, FMAX). */
int pmn_edit_contour(
    const float* sequence, const float* grid, float* out, int items, int t_in, int t_out,
    int log2_domain, float scale, float shift, float lo, float hi, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* PROMONET_B200_H_ */
