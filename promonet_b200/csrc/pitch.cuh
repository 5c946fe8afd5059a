// Pitch / periodicity network handle (pitch.cu)
#pragma once

#include "common.cuh"

struct pmn_pitch;

namespace pmn {

pmn_pitch* pitch_create();
void pitch_destroy(pmn_pitch* p);
int pitch_set_tensor(pmn_pitch* p, const char* name, const float* data, const int64_t* shape,
                     int ndim, cudaStream_t stream);
int pitch_finalize(pmn_pitch* p, int math, cudaStream_t stream);
int pitch_frames(int samples, int sample_rate, double hopsize_seconds);
size_t pitch_workspace_bytes(int batch, int samples, int sample_rate, double hopsize_seconds, int frame_batch);
int pitch_forward(
    pmn_pitch* p, const float* audio, int batch, int samples, int sample_rate,
    double hopsize_seconds, float fmin, float fmax, const float* transition, const float* initial,
    float* pitch, float* periodicity, float* logits_out, int* bins_out, int frame_batch,
    void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace pmn
