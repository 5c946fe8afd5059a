// STFT-derived features (spectral.cu) and Viterbi decoding (viterbi.cu) launchers
#pragma once

#include "common.cuh"

namespace pmn {

// Device tables shared by the STFT kernels, built once per device
struct SpectralTables {
    float* window = nullptr;       // (1024) periodic hann
    float2* twiddle = nullptr;     // (512) exp(-2 pi i k / 1024)
    float* mel_weights = nullptr;  // (80, 513) Slaney basis
    int* mel_range = nullptr;      // (80, 2) first / one-past-last nonzero bin
    float* a_weights = nullptr;    // (513) A-weighting(f_k) - REF_DB
};
int spectral_tables(const SpectralTables** out);

int spectral_frames(int samples);
size_t spectral_workspace_bytes(int batch, int samples);
int launch_spectral_features(
    const float* audio, int batch, int samples, float* magnitude, float* mels, float mel_floor,
    float* loudness, int loudness_bands, void* workspace, size_t workspace_bytes,
    cudaStream_t stream);

int launch_linear_to_mel(
    const float* magnitude, float* mels, float mel_floor, int batch, int frames, cudaStream_t stream);

size_t viterbi_workspace_bytes(int batch, int frames, int states);
int launch_viterbi(
    const float* observation, const int* batch_frames, const float* transition,
    const float* initial, bool log_probs, int* indices, int batch, int frames, int states,
    void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace pmn
