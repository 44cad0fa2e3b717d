// Prompt mel front-end: mel_spectrogram of covomix/util/generate_mel.py:49-72 as ONE kernel.
//   reflect pad (n_fft - hop)/2  ->  hann-windowed frames (torch.stft center=False)  ->  one-sided DFT magnitude
//   sqrt(re^2 + im^2 + 1e-9)  ->  mel filterbank  ->  log(clamp(., 1e-5))
// A prompt is at most 8 s (400 frames of 480 samples, monologue_generation.py:85-88): the whole job is ~0.1 GFLOP, so a
// direct DFT per frame (n_fft = 480 is not a power of two) out of shared memory is all it takes: one CTA per frame, thread
// k owns frequency bin k, the twiddle table cos/sin(2 pi j / n_fft) is indexed by (k * n) mod n_fft, then the first n_mels
// threads take the filterbank rows.  HBM traffic = the waveform once + the mel once.
#pragma once
#include "common.cuh"

namespace covo {

struct MelArgs {
    const float* wav;        // [B, L]
    float* mel;              // [B, n_mels, T]
    const float* window;     // [win]
    const float2* twiddle;   // [n_fft] (cos, sin)(2 pi j / n_fft)
    const float* basis;      // [n_mels, n_freq]
    int L, T, n_fft, hop, win, n_freq, n_mels, pad;
};

__global__ void __launch_bounds__(256) mel_frontend_kernel(const MelArgs a) {
    extern __shared__ float sm[];
    float* frame = sm;                      // [n_fft]
    float2* tw = reinterpret_cast<float2*>(sm + a.n_fft);       // [n_fft]
    float* mag = sm + 3 * a.n_fft;          // [n_freq]
    const int t = blockIdx.x, b = blockIdx.y;
    const float* w = a.wav + static_cast<size_t>(b) * a.L;
    const int off = (a.n_fft - a.win) / 2;  // torch.stft centres a shorter window inside n_fft
    for (int n = threadIdx.x; n < a.n_fft; n += blockDim.x) {
        int p = t * a.hop + n - a.pad;      // index into the un-padded signal; reflect (no edge repeat) at both ends
        if (p < 0) p = -p;
        if (p >= a.L) p = 2 * (a.L - 1) - p;
        const float wn = (n >= off && n < off + a.win) ? a.window[n - off] : 0.f;
        frame[n] = w[p] * wn;
        tw[n] = a.twiddle[n];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.n_freq; k += blockDim.x) {
        float re = 0.f, im = 0.f;
        int idx = 0;
        for (int n = 0; n < a.n_fft; ++n) {
            const float2 c = tw[idx];
            re = fmaf(frame[n], c.x, re);
            im = fmaf(frame[n], -c.y, im);
            idx += k;
            if (idx >= a.n_fft) idx -= a.n_fft;
        }
        mag[k] = sqrtf(re * re + im * im + 1e-9f);
    }
    __syncthreads();
    for (int m = threadIdx.x; m < a.n_mels; m += blockDim.x) {
        const float* br = a.basis + static_cast<size_t>(m) * a.n_freq;
        float s = 0.f;
        for (int k = 0; k < a.n_freq; ++k) s = fmaf(br[k], mag[k], s);
        a.mel[(static_cast<size_t>(b) * a.n_mels + m) * a.T + t] = logf(fmaxf(s, 1e-5f));
    }
}

}  // namespace covo

struct covo_mel {
    covo_mel_cfg cfg;
    covo::DeviceInfo di;
    float* window = nullptr;
    float2* twiddle = nullptr;
    float* basis = nullptr;
    int n_freq = 0;
};
