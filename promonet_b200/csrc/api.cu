// extern "C" surface of libpromonet_b200.so (declared in include/promonet_b200.h)
#include <new>

#include "common.cuh"
#include "conv1d_tc.cuh"
#include "fargan.cuh"
#include "generator.cuh"
#include "pitch.cuh"
#include "spectral.cuh"

#include <map>
#include <mutex>
#include <vector>

namespace pmn {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};
bool g_profile_enabled = false;

namespace {
struct ProfileEntry {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    cudaEvent_t pending = nullptr;
};
std::mutex g_profile_mutex;
std::map<std::string, ProfileEntry> g_profile;
}  // namespace

void profile_before(const char* kernel, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    ProfileEntry& entry = g_profile[kernel];
    cudaEventCreate(&entry.pending);
    cudaEventRecord(entry.pending, stream);
}

void profile_after(const char* kernel, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    ProfileEntry& entry = g_profile[kernel];
    if (!entry.pending) return;
    cudaEvent_t stop;
    cudaEventCreate(&stop);
    cudaEventRecord(stop, stream);
    entry.events.emplace_back(entry.pending, stop);
    entry.pending = nullptr;
}
}  // namespace pmn

using namespace pmn;

extern "C" {

const char* pmn_last_error(void) { return g_last_error.c_str(); }
int pmn_version(void) { return 3; }
int64_t pmn_launch_count(void) { return g_launch_count.load(); }

void pmn_profile_enable(int enabled) { g_profile_enabled = enabled != 0; }

void pmn_profile_reset(void) {
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    for (auto& item : g_profile)
        for (auto& pair : item.second.events) {
            cudaEventDestroy(pair.first);
            cudaEventDestroy(pair.second);
        }
    g_profile.clear();
}

int pmn_profile_read(const char* kernel, double* total_ms, int64_t* launches) {
    PMN_REQUIRE(kernel && total_ms && launches, "profile_read: null pointer");
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    *total_ms = 0.;
    *launches = 0;
    auto it = g_profile.find(kernel);
    if (it == g_profile.end()) return PMN_OK;
    for (auto& pair : it->second.events) {
        PMN_TRY(check_cuda(cudaEventSynchronize(pair.second), "profile event sync"));
        float ms = 0.f;
        PMN_TRY(check_cuda(cudaEventElapsedTime(&ms, pair.first, pair.second), "profile elapsed"));
        *total_ms += ms;
        *launches += 1;
    }
    return PMN_OK;
}

int pmn_generator_create(pmn_generator** out) {
    PMN_REQUIRE(out, "generator_create: null out");
    *out = generator_create();
    if (!*out) return fail(PMN_ERR_STATE, "out of host memory");
    return PMN_OK;
}

void pmn_generator_destroy(pmn_generator* g) { generator_destroy(g); }

int pmn_generator_set_tensor(
    pmn_generator* g, const char* name, const float* data,
    const int64_t* shape, int ndim, void* stream) {
    PMN_REQUIRE(g && name && data && ndim >= 0 && (ndim == 0 || shape), "set_tensor: bad argument");
    return generator_set_tensor(g, name, data, shape, ndim, (cudaStream_t)stream);
}

int pmn_generator_finalize(pmn_generator* g, int math, void* stream) {
    PMN_REQUIRE(g, "finalize: null generator");
    return generator_finalize(g, math, (cudaStream_t)stream);
}

int pmn_generator_set_f8(pmn_generator* g, int enabled) {
    PMN_REQUIRE(g, "set_f8: null generator");
    return generator_set_f8(g, enabled != 0);
}

int pmn_generator_set_pair_mask(pmn_generator* g, unsigned mask) {
    PMN_REQUIRE(g, "set_pair_mask: null generator");
    return generator_set_pair_mask(g, mask);
}

size_t pmn_generator_workspace_bytes(const pmn_generator* g, int batch, int frames) {
    if (!g || batch <= 0 || frames <= 0) return 0;
    return generator_workspace_bytes(g, batch, frames);
}

int pmn_generator_forward(
    pmn_generator* g, const float* loudness, int loudness_rows, const float* pitch,
    const float* periodicity, const float* ppg, const int64_t* speakers,
    const float* sbr, const float* lr, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, void* stream) {
    PMN_REQUIRE(g, "forward: null generator");
    return generator_forward(
        g, loudness, loudness_rows, pitch, periodicity, ppg, speakers, sbr, lr, audio,
        batch, frames, workspace, workspace_bytes, (cudaStream_t)stream);
}

namespace {
struct Staging {
    float *loudness, *pitch, *periodicity, *ppg, *sbr, *lr, *audio;
    int64_t* speakers;
    size_t bytes;
};
Staging carve_staging(void* base, int batch, int frames, int rows) {
    Staging s;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t bytes) {
        char* r = p;
        p += align_up(bytes, 256);
        return r;
    };
    s.loudness = (float*)take((size_t)batch * rows * frames * 4);
    s.pitch = (float*)take((size_t)batch * frames * 4);
    s.periodicity = (float*)take((size_t)batch * frames * 4);
    s.ppg = (float*)take((size_t)batch * 40 * frames * 4);
    s.sbr = (float*)take((size_t)batch * 4);
    s.lr = (float*)take((size_t)batch * 4);
    s.speakers = (int64_t*)take((size_t)batch * 8);
    s.audio = (float*)take((size_t)batch * 256 * frames * 4);
    s.bytes = (size_t)(p - static_cast<char*>(base));
    return s;
}
}  // namespace

size_t pmn_generator_staging_bytes(int batch, int frames, int loudness_rows) {
    if (batch <= 0 || frames <= 0 || loudness_rows <= 0) return 0;
    return carve_staging(nullptr, batch, frames, loudness_rows).bytes;
}

int pmn_generator_forward_host(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, const int64_t* speakers,
    const float* sbr, const float* lr, float* audio, int batch, int frames,
    void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes,
    void* stream_) {
    PMN_REQUIRE(g && loudness && pitch && periodicity && ppg && speakers && sbr && lr && audio && staging,
                "forward_host: null pointer");
    PMN_REQUIRE(batch > 0 && frames > 0 && rows > 0, "forward_host: bad shape");
    cudaStream_t stream = (cudaStream_t)stream_;
    Staging s = carve_staging(staging, batch, frames, rows);
    if (s.bytes > staging_bytes) return fail(PMN_ERR_WORKSPACE, "forward_host: staging too small");
    auto h2d = [&](void* dst, const void* src, size_t bytes) {
        return check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream), "H2D copy");
    };
    PMN_TRY(h2d(s.loudness, loudness, (size_t)batch * rows * frames * 4));
    PMN_TRY(h2d(s.pitch, pitch, (size_t)batch * frames * 4));
    PMN_TRY(h2d(s.periodicity, periodicity, (size_t)batch * frames * 4));
    PMN_TRY(h2d(s.ppg, ppg, (size_t)batch * 40 * frames * 4));
    PMN_TRY(h2d(s.sbr, sbr, (size_t)batch * 4));
    PMN_TRY(h2d(s.lr, lr, (size_t)batch * 4));
    PMN_TRY(h2d(s.speakers, speakers, (size_t)batch * 8));
    PMN_TRY(generator_forward(
        g, s.loudness, rows, s.pitch, s.periodicity, s.ppg, s.speakers, s.sbr, s.lr, s.audio,
        batch, frames, workspace, workspace_bytes, stream));
    return check_cuda(
        cudaMemcpyAsync(audio, s.audio, (size_t)batch * 256 * frames * 4, cudaMemcpyDeviceToHost, stream),
        "D2H copy");
}

int pmn_generator_features(
    pmn_generator* g, const float* loudness, int rows, const float* pitch,
    const float* periodicity, const float* ppg, float* features, int batch, int frames,
    void* stream) {
    PMN_REQUIRE(g, "features: null generator");
    return generator_features(g, loudness, rows, pitch, periodicity, ppg, features, batch, frames,
                              (cudaStream_t)stream);
}

int pmn_fargan_create(pmn_fargan** out) {
    PMN_REQUIRE(out, "fargan_create: null out");
    *out = fargan_create();
    if (!*out) return fail(PMN_ERR_STATE, "out of host memory");
    return PMN_OK;
}

void pmn_fargan_destroy(pmn_fargan* g) { fargan_destroy(g); }

int pmn_fargan_set_tensor(
    pmn_fargan* g, const char* name, const float* data, const int64_t* shape, int ndim, void* stream) {
    PMN_REQUIRE(g && name && data && ndim >= 0 && (ndim == 0 || shape), "fargan_set_tensor: bad argument");
    return fargan_set_tensor(g, name, data, shape, ndim, (cudaStream_t)stream);
}

int pmn_fargan_finalize(pmn_fargan* g, void* stream) {
    PMN_REQUIRE(g, "fargan_finalize: null generator");
    return fargan_finalize(g, (cudaStream_t)stream);
}

size_t pmn_fargan_workspace_bytes(const pmn_fargan*, int batch, int frames) {
    if (batch <= 0 || frames <= 0) return 0;
    return fargan_workspace_bytes(batch, frames);
}

int pmn_fargan_forward(
    pmn_fargan* g, const float* loudness, int rows, const float* pitch, const float* periodicity,
    const float* ppg, const int64_t* speakers, const float* sbr, const float* lr,
    const float* previous_samples, float* audio, int batch, int frames,
    void* workspace, size_t workspace_bytes, void* stream) {
    PMN_REQUIRE(g, "fargan_forward: null generator");
    return fargan_forward(
        g, loudness, rows, pitch, periodicity, ppg, speakers, sbr, lr, previous_samples, audio,
        batch, frames, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t pmn_spectral_workspace_bytes(int batch, int samples) {
    if (batch <= 0 || samples <= 0) return 0;
    return spectral_workspace_bytes(batch, samples);
}

int pmn_spectral_features(
    const float* audio, int batch, int samples, float* magnitude, float* mels, float mel_floor,
    float* loudness, int loudness_bands, void* workspace, size_t workspace_bytes, void* stream) {
    return launch_spectral_features(
        audio, batch, samples, magnitude, mels, mel_floor, loudness, loudness_bands,
        workspace, workspace_bytes, (cudaStream_t)stream);
}

int pmn_linear_to_mel(
    const float* magnitude, float* mels, float mel_floor, int batch, int frames, void* stream) {
    return launch_linear_to_mel(magnitude, mels, mel_floor, batch, frames, (cudaStream_t)stream);
}

size_t pmn_viterbi_workspace_bytes(int batch, int frames, int states) {
    if (batch <= 0 || frames <= 0 || states <= 0) return 0;
    return viterbi_workspace_bytes(batch, frames, states);
}

int pmn_viterbi_decode(
    const float* observation, const int32_t* batch_frames, const float* transition,
    const float* initial, int log_probs, int32_t* indices, int batch, int frames, int states,
    void* workspace, size_t workspace_bytes, void* stream) {
    return launch_viterbi(
        observation, batch_frames, transition, initial, log_probs != 0, indices, batch, frames,
        states, workspace, workspace_bytes, (cudaStream_t)stream);
}

int pmn_pitch_create(pmn_pitch** out) {
    PMN_REQUIRE(out, "pitch_create: null out");
    *out = pitch_create();
    if (!*out) return fail(PMN_ERR_STATE, "out of host memory");
    return PMN_OK;
}

void pmn_pitch_destroy(pmn_pitch* p) { pitch_destroy(p); }

int pmn_pitch_set_tensor(
    pmn_pitch* p, const char* name, const float* data, const int64_t* shape, int ndim, void* stream) {
    PMN_REQUIRE(p && name && data && ndim >= 0 && (ndim == 0 || shape), "pitch_set_tensor: bad argument");
    return pitch_set_tensor(p, name, data, shape, ndim, (cudaStream_t)stream);
}

int pmn_pitch_finalize(pmn_pitch* p, int math, void* stream) {
    PMN_REQUIRE(p, "pitch_finalize: null model");
    return pitch_finalize(p, math, (cudaStream_t)stream);
}

int pmn_pitch_frames(int samples, int sample_rate, double hopsize_seconds) {
    if (samples <= 0 || sample_rate <= 0 || hopsize_seconds <= 0.) return 0;
    return pitch_frames(samples, sample_rate, hopsize_seconds);
}

size_t pmn_pitch_workspace_bytes(
    int batch, int samples, int sample_rate, double hopsize_seconds, int frame_batch) {
    if (batch <= 0 || samples <= 0 || sample_rate <= 0 || hopsize_seconds <= 0. || frame_batch <= 0)
        return 0;
    return pitch_workspace_bytes(batch, samples, sample_rate, hopsize_seconds, frame_batch);
}

int pmn_pitch_forward(
    pmn_pitch* p, const float* audio, int batch, int samples, int sample_rate,
    double hopsize_seconds, float fmin, float fmax, const float* transition, const float* initial,
    float* pitch, float* periodicity, float* logits_out, int32_t* bins_out, int frame_batch,
    void* workspace, size_t workspace_bytes, void* stream) {
    PMN_REQUIRE(p, "pitch_forward: null model");
    return pitch_forward(
        p, audio, batch, samples, sample_rate, hopsize_seconds, fmin, fmax, transition, initial,
        pitch, periodicity, logits_out, bins_out, frame_batch, workspace, workspace_bytes,
        (cudaStream_t)stream);
}

int pmn_weight_norm_fold(const float* v, const float* g, float* w, int dim0, int inner, void* stream) {
    return launch_weight_norm_fold(v, g, w, dim0, inner, (cudaStream_t)stream);
}

int pmn_pack_conv1d_weight(const float* w, float* packed, int c_out, int c_in, int k, void* stream) {
    return launch_pack_conv1d_weight(w, packed, c_out, c_in, k, (cudaStream_t)stream);
}

int pmn_conv1d(
    const float* x, const float* packed_weight, const float* bias, const float* bias2,
    const float* residual, float* out, float* accum, int accum_mode, float accum_scale,
    int batch, int c_in, int c_out, int t_in, int t_out, int k, int dilation, int padding,
    float in_slope, int out_act, void* stream) {
    Conv1dArgs a;
    a.x = x; a.weight = packed_weight; a.bias = bias; a.bias2 = bias2; a.residual = residual;
    a.out = out; a.accum = accum; a.accum_mode = accum_mode; a.accum_scale = accum_scale;
    a.batch = batch; a.c_in = c_in; a.c_out = c_out; a.t_in = t_in; a.t_out = t_out;
    a.k = k; a.dilation = dilation; a.padding = padding; a.in_slope = in_slope; a.out_act = out_act;
    return launch_conv1d(a, (cudaStream_t)stream);
}

namespace {
struct TcOpWorkspace {
    __nv_bfloat16 *x_planes, *y_planes, *slabs;
    size_t bytes;
};
TcOpWorkspace carve_tc_op(void* base, int batch, int channels, int t_len, int k, int c_out = 0) {
    if (c_out <= 0) c_out = channels;
    TcOpWorkspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](size_t elements) {
        auto* r = reinterpret_cast<__nv_bfloat16*>(p);
        p += align_up(elements * sizeof(__nv_bfloat16), 256);
        return r;
    };
    w.x_planes = take(tc_planes_elements(batch, channels, t_len));
    w.y_planes = take(tc_planes_elements(batch, c_out, t_len));
    w.slabs = take(tc_weight_elements(c_out, channels, k));
    w.bytes = (size_t)(p - static_cast<char*>(base));
    return w;
}
}  // namespace

void pmn_debug_tc_counters(void* counters) { tc_set_debug_counters(static_cast<long long*>(counters)); }

size_t pmn_conv1d_tc_workspace_bytes(int batch, int channels, int t_len, int k) {
    if (batch <= 0 || channels <= 0 || t_len <= 0 || k <= 0) return 0;
    return carve_tc_op(nullptr, batch, channels, t_len, k, 512).bytes;
}

namespace {
int conv1d_tc_entry(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int c_in, int c_out, int t_len, int k, int dilation, int valid, int relu,
    float in_slope, float out_slope, bool f8, void* workspace, size_t workspace_bytes, void* stream_) {
    PMN_REQUIRE(x && weight && workspace, "conv1d_tc: null pointer");
    PMN_REQUIRE(batch > 0 && t_len > 0, "conv1d_tc: empty input");
    PMN_REQUIRE(tc_supported(c_in, c_out, k, dilation), "conv1d_tc: unsupported shape");
    PMN_REQUIRE(c_out <= 512, "conv1d_tc: more than 512 output channels");
    cudaStream_t stream = (cudaStream_t)stream_;
    TcOpWorkspace w = carve_tc_op(workspace, batch, c_in, t_len, k, c_out);
    if (w.bytes > workspace_bytes) return fail(PMN_ERR_WORKSPACE, "conv1d_tc: workspace too small");
    PMN_TRY(launch_planes_from_f32(x, w.x_planes, batch, c_in, t_len, in_slope, stream, 0, f8));
    TcConvArgs a;
    if (f8) {
        // "fp16 + 2 x fp8" operands in and out (conv1d_tc.cuh)
        int shift;
        PMN_TRY(tc_f8_weight_shift_of(weight, (size_t)c_out * c_in * k, stream, &shift));
        PMN_TRY(launch_pack_tc_weight_f8(weight, w.slabs, c_out, c_in, k, shift, stream));
        a.f8x2 = a.out_f8 = true;
        a.f8_unscale = tc_f8_unscale(shift);
    } else {
        PMN_TRY(launch_pack_tc_weight(weight, w.slabs, c_out, c_in, k, false, stream));
    }
    a.x_planes = w.x_planes; a.w_slabs = w.slabs; a.bias = bias; a.residual = residual;
    a.out = out; a.accum = accum; a.accum_mode = accum_mode; a.accum_scale = accum_scale;
    a.batch = batch; a.c_in = c_in; a.c_out = c_out; a.t_len = t_len; a.k = k; a.dilation = dilation;
    a.valid = valid != 0; a.relu = relu != 0;
    a.out_slope = out_slope;
    const int t_out = valid ? t_len - (k - 1) * dilation : t_len;
    PMN_REQUIRE(t_out > 0, "conv1d_tc: input shorter than the kernel");
    if (planes_out) {
        a.out_planes = w.y_planes;
        PMN_TRY(launch_zero_plane_pads(w.y_planes, batch, c_out, t_out, stream));
    }
    PMN_TRY(launch_conv1d_tc(a, stream));
    if (planes_out) PMN_TRY(launch_f32_from_planes(w.y_planes, planes_out, batch, c_out, t_out, stream, f8));
    return PMN_OK;
}
}  // namespace

int pmn_conv1d_tc(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float in_slope, float out_slope,
    void* workspace, size_t workspace_bytes, void* stream_) {
    return pmn_conv1d_tc_general(
        x, weight, bias, residual, out, planes_out, accum, accum_mode, accum_scale, batch,
        channels, channels, t_len, k, dilation, 0, 0, in_slope, out_slope, workspace,
        workspace_bytes, stream_);
}

int pmn_conv1d_tc_general(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int c_in, int c_out, int t_len, int k, int dilation, int valid, int relu,
    float in_slope, float out_slope, void* workspace, size_t workspace_bytes, void* stream_) {
    return conv1d_tc_entry(
        x, weight, bias, residual, out, planes_out, accum, accum_mode, accum_scale, batch, c_in, c_out,
        t_len, k, dilation, valid, relu, in_slope, out_slope, false, workspace, workspace_bytes, stream_);
}

int pmn_conv1d_tc_f8(
    const float* x, const float* weight, const float* bias, const float* residual,
    float* out, float* planes_out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float in_slope, float out_slope,
    void* workspace, size_t workspace_bytes, void* stream_) {
    PMN_REQUIRE(tc_f8_plan(channels, channels, nullptr), "conv1d_tc_f8: 128 or 256 channels");
    return conv1d_tc_entry(
        x, weight, bias, residual, out, planes_out, accum, accum_mode, accum_scale, batch, channels,
        channels, t_len, k, dilation, 0, 0, in_slope, out_slope, true, workspace, workspace_bytes, stream_);
}

void pmn_debug_pair_tc(void* counters, int variant) {
    tc_pair_set_debug(static_cast<long long*>(counters), variant);
}

size_t pmn_conv_pair_tc_workspace_bytes(int channels, int k) {
    if (channels <= 0 || k <= 0) return 0;
    return 2 * align_up(tc_weight_elements(channels, channels, k) * sizeof(__nv_bfloat16), 256);
}

int pmn_conv_pair_tc(
    const float* x, const float* weight1, const float* bias1, const float* weight2,
    const float* bias2, float* out, float* accum, int accum_mode, float accum_scale,
    int batch, int channels, int t_len, int k, int dilation, float slope,
    void* workspace, size_t workspace_bytes, void* stream_) {
    PMN_REQUIRE(x && weight1 && weight2 && workspace, "conv_pair_tc: null pointer");
    PMN_REQUIRE(tc_pair_supported(channels, k, dilation), "conv_pair_tc: unsupported shape");
    if (pmn_conv_pair_tc_workspace_bytes(channels, k) > workspace_bytes)
        return fail(PMN_ERR_WORKSPACE, "conv_pair_tc: workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    auto* slabs1 = static_cast<__nv_bfloat16*>(workspace);
    auto* slabs2 = reinterpret_cast<__nv_bfloat16*>(
        static_cast<char*>(workspace) + pmn_conv_pair_tc_workspace_bytes(channels, k) / 2);
    PMN_TRY(launch_pack_tc_weight(weight1, slabs1, channels, channels, k, false, stream));
    PMN_TRY(launch_pack_tc_weight(weight2, slabs2, channels, channels, k, false, stream));
    return launch_conv_pair_tc(
        x, slabs1, bias1, slabs2, bias2, out, accum, accum_mode, accum_scale, batch, channels,
        t_len, k, dilation, slope, stream);
}

int pmn_conv_transpose1d(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope, void* stream) {
    return launch_conv_transpose1d(x, weight, bias, out, batch, c_in, c_out, t_in, k, stride,
                                   in_slope, (cudaStream_t)stream);
}

size_t pmn_conv_transpose1d_tc_workspace_bytes(int batch, int c_in, int t_in, int stride) {
    if (batch <= 0 || c_in <= 0 || t_in <= 0 || stride <= 0) return 0;
    return align_up(tc_planes_elements(batch, c_in, t_in) * 2, 256) +
           align_up(tc_transpose_weight_elements(c_in, c_in / 2, stride) * 2, 256);
}

int pmn_conv_transpose1d_tc(
    const float* x, const float* weight, const float* bias, float* out,
    int batch, int c_in, int c_out, int t_in, int k, int stride, float in_slope,
    void* workspace, size_t workspace_bytes, void* stream_) {
    PMN_REQUIRE(x && weight && out && workspace, "conv_transpose1d_tc: null pointer");
    PMN_REQUIRE(batch > 0 && t_in > 0, "conv_transpose1d_tc: empty input");
    PMN_REQUIRE(tc_transpose_supported(c_in, c_out, k, stride), "conv_transpose1d_tc: unsupported shape");
    if (pmn_conv_transpose1d_tc_workspace_bytes(batch, c_in, t_in, stride) > workspace_bytes)
        return fail(PMN_ERR_WORKSPACE, "conv_transpose1d_tc: workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    auto* planes = static_cast<__nv_bfloat16*>(workspace);
    auto* slabs = reinterpret_cast<__nv_bfloat16*>(
        static_cast<char*>(workspace) + align_up(tc_planes_elements(batch, c_in, t_in) * 2, 256));
    PMN_TRY(launch_planes_from_f32(x, planes, batch, c_in, t_in, in_slope, stream));
    PMN_TRY(launch_pack_tc_transpose_weight(weight, slabs, c_in, c_out, stride, stream));
    TcConvArgs a;
    a.x_planes = planes; a.w_slabs = slabs; a.bias = bias; a.out = out;
    a.batch = batch; a.c_in = c_in; a.c_out = c_out; a.t_len = t_in;
    return launch_conv_transpose1d_tc(a, stride, stream);
}

}  // extern "C"
