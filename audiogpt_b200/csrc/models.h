// Internal C++ entry points behind the C ABI (include/agpt_b200.h).
#pragma once
#include "../../include/agpt_b200.h"
#include "common.cuh"

namespace agpt {

void count_launch(long n);

void launch_cf_to_cl(const float* in, float* out, int B, int C, int T, cudaStream_t st);

Handle* hifigan_create(const agpt_hifigan_cfg* cfg, const float* const* W, int nW, int device);
void hifigan_forward(Handle* h, const float* mel, const float* har, int B, int T, float* wav, cudaStream_t st);
void hifigan_vocode_host(Handle* h, const float* mel_host, const float* har_host, int B, int T, float* wav_host);

}  // namespace agpt
