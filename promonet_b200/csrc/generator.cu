// HiFi-GAN generator handle: weight store, weight-norm folding, forward.
//
// Replaces promonet.model.Generator (promonet/model/generator.py:84-197) with
// MODEL='hifigan' (promonet/model/hifigan.py:13-70) for inference.
#include <algorithm>
#include <map>
#include <memory>
#include <new>
#include <vector>

#include "conv1d_tc.cuh"
#include "features.cuh"
#include "generator.cuh"
#include "tensor_store.cuh"

namespace pmn {

namespace {

constexpr int kNumFeatures = 113;   // config/static.py:47-52
constexpr int kPaddedFeatures = 128;  // the input convolution's K on the tensor cores
constexpr int kSpeakerChannels = 256;
constexpr int kGlobalChannels = 258;  // static.py:40-43
constexpr int kNumSpeakers = 109;     // static.py:58-59 (vctk)
constexpr int kInitial = 512;         // HIFIGAN_UPSAMPLE_INITIAL_SIZE defaults.py:256
constexpr int kStages = 4;
constexpr int kUpKernel[kStages] = {16, 16, 4, 4};  // :259
constexpr int kUpRate[kStages] = {8, 8, 2, 2};      // :262
constexpr int kResKernel[3] = {3, 7, 11};           // :250
constexpr int kResDilation[3] = {1, 3, 5};          // :253
constexpr float kSlope = 0.1f;                      // :216
// Which residual blocks take the fused pair kernel by default; bit 3 * stage + block, block =
// kernel 3 / 7 / 11.  Measured on B200 (profiles/r2_pair_breakdown.txt, r2_pair_selection*.txt):
// the longer kernels are bound by the tensor pipe's operand fetch from shared memory, not by HBM,
// and are faster as two launches (larger tiles, half the weight streaming, no converter traffic);
// keeping the pair on chip pays where the two-launch path is HBM-bound, i.e. for the k = 3 blocks:
// of stage 1 (C = 128) since the kernel was written, of stages 2 and 3 (C = 64, 32) since its
// converter and epilogue warps address shared memory as such and step their global pointers
// (each +0.2 ms of 25.7, together 0.55: profiles/r2_pair_selection_final.txt).
constexpr unsigned kDefaultPairMask = 0x248u;


struct PackedConv {
    float* weight = nullptr;  // conv1d: (C_in, K, C_out); conv transpose: (C_in, C_out, K)
    __nv_bfloat16* slabs = nullptr;  // tensor-core path: hi/lo weight slabs (conv1d_tc.cuh)
    const float* bias = nullptr;
    int c_in = 0, c_out = 0, k = 0;
    // C >= 128: the same weights as "fp16 + 2 x fp8" slabs (conv1d_tc.cuh) and their power-of-two scale
    void* f8_slabs = nullptr;
    int f8_shift = 0;
};

}  // namespace

}  // namespace pmn

struct pmn_generator {
    pmn::TensorStore store;
    bool finalized = false;
    int math = PMN_MATH_FP32_SIMT;
    float ppg_threshold = 0.85f;
    // bit (3 * stage + block): run that residual block's three c1 -> c2 pairs as fused
    // conv_pair_tc_kernel launches (tensor-core math, C <= 128) instead of six conv1d_tc ones
    unsigned pair_mask = pmn::kDefaultPairMask;
    // residual blocks of the C = 128 stage with "fp16 + 2 x fp8" operands (pmn_generator_set_f8)
    bool f8 = false;

    pmn::PackedConv up[pmn::kStages];
    pmn::PackedConv conv1[pmn::kStages][3][3];
    pmn::PackedConv conv2[pmn::kStages][3][3];
    float* input_weight = nullptr;  // packed (113, 7, 512)
    __nv_bfloat16* input_slabs = nullptr;   // tensor-core path: the same convolution with 128 input channels
};

namespace pmn {

namespace {

int alloc(pmn_generator* g, size_t count, float** out) { return g->store.alloc(count, out); }

// The stage whose residual blocks take "fp16 + 2 x fp8" operands when the generator's f8 switch is
// on: C = 128 (k = 7 / 11 launches 8 - 13 % faster than bf16 x 3).  C = 256 is built and tested
// (pmn_conv1d_tc_f8) but slower in that form -- its 128-column tiles lose more than the cheaper
// corrections win (k = 11: 0.36 against 0.30 ms, profiles/r2_f8_breakdown.txt)
bool f8_stage(int channels) { return channels == 128; }

int find(const pmn_generator* g, const std::string& name, const Tensor** out) {
    return g->store.find(name, out);
}

// Resolve `<prefix>.weight`, folding `<prefix>.weight_g/_v` when that is what the
// checkpoint holds (weight_norm keys, SURVEY 5 "Checkpoint / resume")
int folded_weight(pmn_generator* g, const std::string& prefix, const Tensor** v_out,
                  const float** w_out, cudaStream_t stream) {
    if (g->store.has(prefix + ".weight")) {
        PMN_TRY(find(g, prefix + ".weight", v_out));
        *w_out = (*v_out)->data;
        return PMN_OK;
    }
    const Tensor *wg, *wv;
    PMN_TRY(find(g, prefix + ".weight_g", &wg));
    PMN_TRY(find(g, prefix + ".weight_v", &wv));
    if (wv->shape.size() != 3 || wg->numel() != (size_t)wv->shape[0])
        return fail(PMN_ERR_STATE, "bad weight_g/weight_v shapes at " + prefix);
    float* w;
    PMN_TRY(alloc(g, wv->numel(), &w));
    PMN_TRY(launch_weight_norm_fold(
        wv->data, wg->data, w, (int)wv->shape[0], (int)(wv->shape[1] * wv->shape[2]), stream));
    *v_out = wv;
    *w_out = w;
    return PMN_OK;
}

int prepare_conv(pmn_generator* g, const std::string& prefix, int channels, int k,
                 PackedConv* conv, cudaStream_t stream) {
    const Tensor* shape;
    const float* w;
    PMN_TRY(folded_weight(g, prefix, &shape, &w, stream));
    if (shape->shape[0] != channels || shape->shape[1] != channels || shape->shape[2] != k)
        return fail(PMN_ERR_STATE, "unexpected conv shape at " + prefix);
    const Tensor* bias;
    PMN_TRY(find(g, prefix + ".bias", &bias));
    conv->c_in = conv->c_out = channels;
    conv->k = k;
    conv->bias = bias->data;
    if (g->math == PMN_MATH_BF16X3_TC) {
        float* slabs;  // two bf16 planes = the bytes of one fp32 tensor (+ the narrow layers' second format)
        PMN_TRY(alloc(g, (tc_weight_elements(channels, channels, k) + 1) / 2, &slabs));
        conv->slabs = reinterpret_cast<__nv_bfloat16*>(slabs);
        if (f8_stage(channels)) {
            float* f8_slabs;
            PMN_TRY(alloc(g, shape->numel(), &f8_slabs));
            PMN_TRY(tc_f8_weight_shift_of(w, shape->numel(), stream, &conv->f8_shift));
            PMN_TRY(launch_pack_tc_weight_f8(w, f8_slabs, channels, channels, k, conv->f8_shift, stream));
            conv->f8_slabs = f8_slabs;
        }
        return launch_pack_tc_weight(w, conv->slabs, channels, channels, k, false, stream);
    }
    PMN_TRY(alloc(g, shape->numel(), &conv->weight));
    return launch_pack_conv1d_weight(w, conv->weight, channels, channels, k, stream);
}

struct Workspace {
    float *features, *speaker_bias, *x_in, *x0, *xt, *cur, *mrf;
    __nv_bfloat16 *a0, *at, *ac;  // tensor-core path: hi/lo planes of lrelu(x0 / xt / cur)
    size_t bytes;
};

Workspace carve(void* base, int batch, int frames, int math) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t count) {
        float* r = reinterpret_cast<float*>(p);
        p += align_up(count * sizeof(float), 256);
        return r;
    };
    const size_t stage = (size_t)batch * 8192 * frames;  // max C*T over stages = 32 * 256 F
    w.features = take((size_t)batch * (kNumFeatures + 1) * frames);
    w.speaker_bias = take((size_t)batch * kInitial);
    w.x_in = take((size_t)batch * kInitial * frames);
    w.x0 = take(stage);
    w.cur = take(stage);
    w.mrf = take(stage);
    w.xt = take(stage);
    w.a0 = w.at = w.ac = nullptr;
    if (math == PMN_MATH_BF16X3_TC) {
        // planes are (B, 2, C, t_pad) bf16; the widest is the last stage (C = 32)
        size_t planes = 0;
        int channels = kInitial, t_len = frames;
        for (int s = 0; s < kStages; ++s) {
            channels /= 2;
            t_len *= kUpRate[s];
            planes = std::max(planes, tc_planes_elements(batch, channels, t_len));
        }
        auto take_planes = [&]() { return reinterpret_cast<__nv_bfloat16*>(take((planes + 1) / 2)); };
        w.a0 = take_planes();
        w.at = take_planes();
        w.ac = take_planes();
    }
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}

}  // namespace

pmn_generator* generator_create() { return new (std::nothrow) pmn_generator(); }
void generator_destroy(pmn_generator* g) { delete g; }

int generator_set_tensor(
    pmn_generator* g, const char* name, const float* data, const int64_t* shape, int ndim,
    cudaStream_t stream) {
    if (g->finalized) return fail(PMN_ERR_STATE, "set_tensor after finalize");
    return g->store.set(name, data, shape, ndim, stream);
}

int generator_finalize(pmn_generator* g, int math, cudaStream_t stream) {
    if (g->finalized) return fail(PMN_ERR_STATE, "generator already finalized");
    if (math != PMN_MATH_FP32_SIMT && math != PMN_MATH_BF16X3_TC)
        return fail(PMN_ERR_ARGUMENT, "generator: unsupported math mode");
    g->math = math;

    const Tensor* t;
    PMN_TRY(find(g, "model.input_feature_conv.weight", &t));
    if (t->shape.size() != 3 || t->shape[0] != kInitial || t->shape[1] != kNumFeatures || t->shape[2] != 7)
        return fail(PMN_ERR_STATE, "unexpected input_feature_conv shape");
    PMN_TRY(alloc(g, t->numel(), &g->input_weight));
    PMN_TRY(launch_pack_conv1d_weight(t->data, g->input_weight, kInitial, kNumFeatures, 7, stream));
    if (math == PMN_MATH_BF16X3_TC) {
        // (512, 113, 7) -> (512, 128, 7) with zero channels, then the tensor-core slabs
        float *padded, *slabs;
        const size_t padded_numel = (size_t)kInitial * kPaddedFeatures * 7;
        PMN_TRY(alloc(g, padded_numel, &padded));
        PMN_TRY(check_cuda(cudaMemsetAsync(padded, 0, padded_numel * sizeof(float), stream), "memset"));
        PMN_TRY(check_cuda(
            cudaMemcpy2DAsync(padded, (size_t)kPaddedFeatures * 7 * sizeof(float), t->data,
                              (size_t)kNumFeatures * 7 * sizeof(float), (size_t)kNumFeatures * 7 * sizeof(float),
                              kInitial, cudaMemcpyDeviceToDevice, stream),
            "pad input weight"));
        PMN_TRY(alloc(g, (tc_weight_elements(kInitial, kPaddedFeatures, 7) + 1) / 2, &slabs));
        g->input_slabs = reinterpret_cast<__nv_bfloat16*>(slabs);
        PMN_TRY(launch_pack_tc_weight(padded, g->input_slabs, kInitial, kPaddedFeatures, 7, false, stream));
    }
    PMN_TRY(find(g, "model.input_feature_conv.bias", &t));
    PMN_TRY(find(g, "model.input_speaker_conv.weight", &t));
    if (t->numel() != (size_t)kInitial * kGlobalChannels)
        return fail(PMN_ERR_STATE, "unexpected input_speaker_conv shape");
    PMN_TRY(find(g, "model.input_speaker_conv.bias", &t));
    PMN_TRY(find(g, "speaker_embedding.weight", &t));
    if (t->numel() != (size_t)kNumSpeakers * kSpeakerChannels)
        return fail(PMN_ERR_STATE, "unexpected speaker_embedding shape");
    PMN_TRY(find(g, "pitch_embedding.weight", &t));
    if (t->numel() != 256 * 64) return fail(PMN_ERR_STATE, "unexpected pitch_embedding shape");
    PMN_TRY(find(g, "pitch_distribution", &t));
    if (t->numel() != 256) return fail(PMN_ERR_STATE, "unexpected pitch_distribution shape");
    PMN_TRY(find(g, "model.model.5.weight", &t));
    if (t->numel() != 32 * 7) return fail(PMN_ERR_STATE, "unexpected output conv shape");
    if (g->store.has("ppg_threshold")) {
        PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "sync"));
        PMN_TRY(check_cuda(
            cudaMemcpy(&g->ppg_threshold, g->store.data("ppg_threshold"), sizeof(float), cudaMemcpyDeviceToHost),
            "read ppg_threshold"));
    }

    int channels = kInitial;
    for (int s = 0; s < kStages; ++s) {
        const std::string stage = "model.model." + std::to_string(s) + ".model";
        const Tensor* shape;
        const float* w;
        PMN_TRY(folded_weight(g, stage + ".1", &shape, &w, stream));
        if (shape->shape[0] != channels || shape->shape[1] != channels / 2 || shape->shape[2] != kUpKernel[s])
            return fail(PMN_ERR_STATE, "unexpected upsample shape at " + stage);
        const Tensor* bias;
        PMN_TRY(find(g, stage + ".1.bias", &bias));
        g->up[s].weight = const_cast<float*>(w);
        if (g->math == PMN_MATH_BF16X3_TC) {
            float* slabs;
            const size_t elements = tc_transpose_weight_elements(channels, channels / 2, kUpRate[s]);
            PMN_TRY(alloc(g, (elements + 1) / 2, &slabs));
            g->up[s].slabs = reinterpret_cast<__nv_bfloat16*>(slabs);
            PMN_TRY(launch_pack_tc_transpose_weight(
                w, g->up[s].slabs, channels, channels / 2, kUpRate[s], stream));
        }
        g->up[s].bias = bias->data;
        g->up[s].c_in = channels;
        g->up[s].c_out = channels / 2;
        g->up[s].k = kUpKernel[s];
        channels /= 2;
        for (int j = 0; j < 3; ++j) {
            const std::string block = stage + ".2.model." + std::to_string(j);
            for (int d = 0; d < 3; ++d) {
                PMN_TRY(prepare_conv(g, block + ".convs1." + std::to_string(d), channels,
                                     kResKernel[j], &g->conv1[s][j][d], stream));
                PMN_TRY(prepare_conv(g, block + ".convs2." + std::to_string(d), channels,
                                     kResKernel[j], &g->conv2[s][j][d], stream));
            }
        }
    }
    PMN_TRY(check_cuda(cudaStreamSynchronize(stream), "finalize sync"));
    g->finalized = true;
    return PMN_OK;
}

int generator_set_f8(pmn_generator* g, bool enabled) {
    if (enabled && g->finalized && g->math != PMN_MATH_BF16X3_TC)
        return fail(PMN_ERR_STATE, "generator: the fp8 form belongs to the tensor-core math mode");
    g->f8 = enabled;
    return PMN_OK;
}

int generator_set_pair_mask(pmn_generator* g, unsigned mask) {
    g->pair_mask = mask;
    return PMN_OK;
}

size_t generator_workspace_bytes(const pmn_generator* g, int batch, int frames) {
    return carve(nullptr, batch, frames, g->math).bytes;
}

int generator_features(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, float* features, int batch, int frames,
    cudaStream_t stream) {
    if (!g->finalized) return fail(PMN_ERR_STATE, "generator not finalized");
    return launch_features(
        loudness, rows, pitch, periodicity, ppg,
        g->store.data("pitch_distribution"), g->store.data("pitch_embedding.weight"),
        g->ppg_threshold, false, features, batch, frames, stream);
}

int generator_forward(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, const int64_t* speakers,
    const float* sbr, const float* lr, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!g->finalized) return fail(PMN_ERR_STATE, "generator not finalized");
    PMN_REQUIRE(batch > 0 && frames > 0, "generator: empty batch");
    PMN_REQUIRE(audio && workspace, "generator: null pointer");
    Workspace w = carve(workspace, batch, frames, g->math);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "generator: workspace too small");

    // G1: features (B, 113, F)
    PMN_TRY(generator_features(g, loudness, rows, pitch, periodicity, ppg, w.features, batch, frames, stream));
    // G2 + speaker 1x1 conv: (B, 512) bias
    PMN_TRY(launch_speaker_bias(
        g->store.data("speaker_embedding.weight"), speakers, sbr, lr,
        g->store.data("model.input_speaker_conv.weight"),
        g->store.data("model.input_speaker_conv.bias"),
        w.speaker_bias, batch, kSpeakerChannels, kInitial, kNumSpeakers, stream));
    // G3: input conv k7 + speaker bias.  Tensor-core path: the 113 feature channels padded to 128,
    // the speaker projection as a per-item bias, and the epilogue writes the LeakyReLU'd planes the
    // first ConvTranspose1d reads (the fp32 tensor is not needed)
    const bool input_on_tensor_cores = g->math == PMN_MATH_BF16X3_TC;
    if (input_on_tensor_cores) {
        PMN_TRY(launch_planes_from_f32(
            w.features, w.a0, batch, kPaddedFeatures, frames, 1.f, stream, kNumFeatures));
        PMN_TRY(launch_zero_plane_pads(w.at, batch, kInitial, frames, stream));
        TcConvArgs a;
        a.x_planes = w.a0; a.w_slabs = g->input_slabs;
        a.bias = g->store.data("model.input_feature_conv.bias"); a.bias_batch = w.speaker_bias;
        a.out_planes = w.at; a.out_slope = kSlope;
        a.batch = batch; a.c_in = kPaddedFeatures; a.c_out = kInitial; a.t_len = frames; a.k = 7;
        PMN_TRY(launch_conv1d_tc(a, stream));
    } else {
        Conv1dArgs a;
        a.x = w.features; a.weight = g->input_weight;
        a.bias = g->store.data("model.input_feature_conv.bias");
        a.bias2 = w.speaker_bias;
        a.out = w.x_in;
        a.batch = batch; a.c_in = kNumFeatures; a.c_out = kInitial;
        a.t_in = a.t_out = frames; a.k = 7; a.padding = 3;
        PMN_TRY(launch_conv1d(a, stream));
    }

    const float* stage_in = w.x_in;
    int t_len = frames;
    bool planes_ready = false;   // w.ac holds the planes of lrelu(stage_in)
    for (int s = 0; s < kStages; ++s) {
        // G4: LeakyReLU + ConvTranspose1d
        const PackedConv& up = g->up[s];
        if (g->math == PMN_MATH_BF16X3_TC) {
            // the planes of lrelu(stage input) come from the epilogue that produced it: the input
            // convolution (stage 0, in w.at) or the last residual-block launch of the stage before
            // (in w.ac), unless that one ran as a fused pair
            const __nv_bfloat16* up_planes = s == 0 ? w.at : w.ac;
            if (!(s == 0 ? input_on_tensor_cores : planes_ready)) {
                PMN_TRY(launch_planes_from_f32(stage_in, w.at, batch, up.c_in, t_len, kSlope, stream));
                up_planes = w.at;
            }
            TcConvArgs a;
            a.x_planes = up_planes; a.w_slabs = up.slabs; a.bias = up.bias; a.out = w.x0;
            a.batch = batch; a.c_in = up.c_in; a.c_out = up.c_out; a.t_len = t_len;
            // ... and writes the planes of lrelu(its output), the residual blocks' first operand
            PMN_TRY(launch_zero_plane_pads(w.a0, batch, up.c_out, t_len * kUpRate[s], stream));
            a.out_planes = w.a0; a.out_slope = kSlope;
            a.out_f8 = g->f8 && f8_stage(up.c_out);
            PMN_TRY(launch_conv_transpose1d_tc(a, kUpRate[s], stream));
        } else {
            PMN_TRY(launch_conv_transpose1d(
                stage_in, up.weight, up.bias, w.x0, batch, up.c_in, up.c_out, t_len,
                up.k, kUpRate[s], kSlope, stream));
        }
        t_len *= kUpRate[s];
        const int channels = up.c_out;
        if (g->math == PMN_MATH_BF16X3_TC) {
            // G5/G6 on tcgen05.  Blocks selected by pair_mask run as three fused pair
            // launches over the fp32 stream (x0 -> cur -> xt -> mean); the others as six
            // launches with the activations travelling between them as bf16 hi/lo planes
            // of lrelu(.) next to the fp32 residual stream
            // stage_f8: every plane of this stage is a "fp16 + 2 x fp8" operand (same bytes, same pad
            // rows) and its convolutions take the fp8 slabs
            const bool stage_f8 = g->f8 && f8_stage(channels);
            bool fused[3], any_planes = false;
            for (int j = 0; j < 3; ++j) {
                fused[j] = ((g->pair_mask >> (3 * s + j)) & 1u) != 0 &&
                           tc_pair_supported(channels, kResKernel[j], kResDilation[2]);
                any_planes = any_planes || !fused[j];
            }
            if (any_planes) {
                PMN_TRY(launch_zero_plane_pads(w.at, batch, channels, t_len, stream));
                PMN_TRY(launch_zero_plane_pads(w.ac, batch, channels, t_len, stream));
            }
            for (int j = 0; j < 3; ++j) {
                if (fused[j]) {
                    const float* in = w.x0;
                    for (int d = 0; d < 3; ++d) {
                        const PackedConv& c1 = g->conv1[s][j][d];
                        const PackedConv& c2 = g->conv2[s][j][d];
                        float* out = d == 0 ? w.cur : d == 1 ? w.xt : nullptr;
                        PMN_TRY(launch_conv_pair_tc(
                            in, c1.slabs, c1.bias, c2.slabs, c2.bias, out,
                            d == 2 ? w.mrf : nullptr, j == 0 ? 1 : 2, 1.f / 3.f,
                            batch, channels, t_len, kResKernel[j], kResDilation[d], kSlope, stream));
                        in = out;
                    }
                    continue;
                }
                for (int d = 0; d < 3; ++d) {
                    TcConvArgs a;
                    a.batch = batch; a.c_in = a.c_out = channels; a.t_len = t_len;
                    a.k = kResKernel[j]; a.out_slope = kSlope;
                    const PackedConv& c1 = g->conv1[s][j][d];
                    const PackedConv& c2 = g->conv2[s][j][d];
                    a.f8x2 = a.out_f8 = stage_f8;
                    a.x_planes = d == 0 ? w.a0 : w.ac;
                    a.w_slabs = stage_f8 ? static_cast<const __nv_bfloat16*>(c1.f8_slabs) : c1.slabs;
                    a.f8_unscale = tc_f8_unscale(c1.f8_shift);
                    a.bias = c1.bias;
                    a.dilation = kResDilation[d];
                    a.out_planes = w.at;
                    PMN_TRY(launch_conv1d_tc(a, stream));
                    a.x_planes = w.at;
                    a.w_slabs = stage_f8 ? static_cast<const __nv_bfloat16*>(c2.f8_slabs) : c2.slabs;
                    a.f8_unscale = tc_f8_unscale(c2.f8_shift);
                    a.bias = c2.bias;
                    a.dilation = 1;
                    a.residual = d == 0 ? w.x0 : w.cur;
                    if (d < 2) {
                        a.out = w.cur;
                        a.out_planes = w.ac;
                    } else {
                        a.out_planes = nullptr;
                        a.accum = w.mrf;
                        a.accum_mode = j == 0 ? 1 : 2;
                        a.accum_scale = 1.f / 3.f;
                        if (j == 2 && s + 1 < kStages) {
                            // the stage's output is complete with this launch: its LeakyReLU'd planes,
                            // the next stage's ConvTranspose1d operand, go to w.ac (last read by this
                            // pair's first convolution)
                            a.out_planes = w.ac;
                            a.planes_from_accum = true;
                            a.out_f8 = false;    // the transposed convolution takes bf16 hi / lo planes
                        }
                    }
                    PMN_TRY(launch_conv1d_tc(a, stream));
                }
            }
            stage_in = w.mrf;
            planes_ready = !fused[2] && s + 1 < kStages;
            continue;
        }
        // G5/G6: three Blocks, mean folded into the last conv of each
        for (int j = 0; j < 3; ++j) {
            const float* block_in = w.x0;
            for (int d = 0; d < 3; ++d) {
                const PackedConv& c1 = g->conv1[s][j][d];
                const PackedConv& c2 = g->conv2[s][j][d];
                Conv1dArgs a;
                a.batch = batch; a.c_in = a.c_out = channels; a.t_in = a.t_out = t_len;
                a.k = c1.k; a.in_slope = kSlope;
                a.x = block_in; a.weight = c1.weight; a.bias = c1.bias; a.out = w.xt;
                a.dilation = kResDilation[d];
                a.padding = kResDilation[d] * (c1.k - 1) / 2;
                PMN_TRY(launch_conv1d(a, stream));
                a.x = w.xt; a.weight = c2.weight; a.bias = c2.bias;
                a.dilation = 1; a.padding = (c2.k - 1) / 2;
                a.residual = block_in;
                if (d < 2) {
                    a.out = w.cur;
                } else {
                    a.out = nullptr;
                    a.accum = w.mrf;
                    a.accum_mode = j == 0 ? 1 : 2;
                    a.accum_scale = 1.f / 3.f;
                }
                PMN_TRY(launch_conv1d(a, stream));
                block_in = w.cur;
            }
        }
        stage_in = w.mrf;
    }
    // G7: LeakyReLU + Conv1d(32 -> 1, k7, no bias) + tanh
    return launch_head(stage_in, g->store.data("model.model.5.weight"), audio, batch, 32,
                       t_len, kSlope, stream);
}

}  // namespace pmn
