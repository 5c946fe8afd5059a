// STFT-derived features (spectral.cu) and Viterbi decoding (viterbi.cu) launchers
#pragma once

#include "common.cuh"

namespace pmn {

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
