// NOT YET RUN ON HARDWARE (validated on the CPU emulation of this source against torch.autograd).  Next scope row: the
// backward of the log-mel spectrogram that the mel-reconstruction loss of the VQ-VAE-GAN step differentiates through
// (ttts/vqvae/train.py:357-366,389: F.l1_loss(y_mel, mel_spectrogram_torch(y_hat)) * c_mel; forward = csrc/stft.cu, ttts_stft_mel):
//   x_pad = reflect-pad(wav, pad) ; X_k = sum_n win[n] x_pad[f hop + n] e^{-2 pi i k n / N} ; mag_k = sqrt(re^2 + im^2 + eps)
//   mel_m = sum_k basis[m, k] mag_k ; out = log(max(mel_m, floor))
// One CTA per frame, correctness first: the DFT and its adjoint are evaluated directly (2 x 2.1 M multiply-adds per 2048-point frame --
// a train step differentiates ~20 frames per clip, so this is 5 GFLOP per 64 clips) with a shared-memory twiddle table built by
// sincospif; the frame's gradient is scattered into dwav with atomic adds through the reflect-padding index map (frames overlap).
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace ttts {

struct StftBwdParams {
    const float* wav; const float* window; const float* dlogmel; float* dwav;
    const int32_t* band_lo; const int32_t* band_off; const float* band_w;
    int B, L, n_fft, hop, pad, n_mels, n_frames;
    float eps_inside, log_floor;
};

__global__ void __launch_bounds__(256) stft_mel_bwd_kernel(const StftBwdParams p) {
    TTTS_DYN_SMEM(float, sb);
    const int N = p.n_fft, NB = N / 2 + 1;
    float* tw_c = sb;                 // [N] cos(2 pi j / N)
    float* tw_s = tw_c + N;           // [N] sin(2 pi j / N)
    float* fr = tw_s + N;             // [N] windowed frame
    float* re = fr + N;               // [NB] -> d re
    float* im = re + NB;              // [NB] -> d im
    float* mag = im + NB;             // [NB] -> d mag
    float* dmel = mag + NB;           // [n_mels]
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* xw = p.wav + (size_t)b * p.L;
    auto src_index = [&](int j) {     // padded index -> waveform index (torch reflect padding)
        int i = j - p.pad;
        if (i < 0) i = -i;
        if (i >= p.L) i = 2 * (p.L - 1) - i;
        return i;
    };
    for (int j = tid; j < N; j += 256) {
        float s, c;
        sincospif(2.0f * (float)j / (float)N, &s, &c);
        tw_c[j] = c; tw_s[j] = s;
        fr[j] = p.window[j] * xw[src_index(f * p.hop + j)];
    }
    __syncthreads();
    // forward DFT, bins 0 .. N/2
    for (int k = tid; k < NB; k += 256) {
        float ar = 0.f, ai = 0.f;
        int ph = 0;                                              // k n mod N
        for (int n = 0; n < N; ++n) {
            ar = fmaf(fr[n], tw_c[ph], ar);
            ai = fmaf(-fr[n], tw_s[ph], ai);
            ph += k; if (ph >= N) ph -= N;
        }
        re[k] = ar; im[k] = ai;
        mag[k] = sqrtf(ar * ar + ai * ai + p.eps_inside);
    }
    __syncthreads();
    // mel bands and the gradient of log(max(mel, floor))
    for (int m = tid; m < p.n_mels; m += 256) {
        const int lo = p.band_lo[m], o0 = p.band_off[m], cnt = p.band_off[m + 1] - o0;
        float s = 0.f;
        for (int i = 0; i < cnt; ++i) s = fmaf(p.band_w[o0 + i], mag[lo + i], s);
        const float g = p.dlogmel[((size_t)b * p.n_mels + m) * p.n_frames + f];
        dmel[m] = s > p.log_floor ? g / s : 0.f;
    }
    __syncthreads();
    // d mag -> d re, d im
    for (int k = tid; k < NB; k += 256) {
        float dm = 0.f;
        for (int m = 0; m < p.n_mels; ++m) {
            const int i = k - p.band_lo[m], o0 = p.band_off[m];
            if (i >= 0 && i < p.band_off[m + 1] - o0) dm = fmaf(p.band_w[o0 + i], dmel[m], dm);
        }
        const float inv = dm / mag[k];
        re[k] *= inv; im[k] *= inv;                              // now d re, d im
    }
    __syncthreads();
    // adjoint DFT and window, scattered through the padding map
    for (int n = tid; n < N; n += 256) {
        float s = 0.f;
        int ph = 0;                                              // k n mod N
        for (int k = 0; k < NB; ++k) {
            s = fmaf(re[k], tw_c[ph], s);
            s = fmaf(-im[k], tw_s[ph], s);
            ph += n; if (ph >= N) ph -= N;
        }
        atomicAdd(p.dwav + (size_t)b * p.L + src_index(f * p.hop + n), s * p.window[n]);
    }
}

}  // namespace ttts

/* dwav [B, L] ACCUMULATES (zero it first).  Arguments as ttts_stft_mel; dlogmel [B, n_mels, n_frames] is the gradient of its mel_out. */
extern "C" int ttts_stft_mel_bwd(const float* wav, int32_t B, int32_t L, int32_t n_fft, int32_t hop, int32_t pad, const float* window, float eps_inside,
                                 int32_t n_mels, const int32_t* band_lo, const int32_t* band_off, const float* band_w, float log_floor,
                                 const float* dlogmel, int32_t n_frames, float* dwav, void* stream) {
    using namespace ttts;
    TTTS_CHECK_ARG(wav && window && band_lo && band_off && band_w && dlogmel && dwav, "stft backward: null pointer");
    TTTS_CHECK_ARG(B >= 1 && B <= 65535 && L >= 2 && n_fft >= 2 && n_fft % 2 == 0 && hop >= 1 && pad >= 0 && pad < L && n_mels >= 1, "stft backward: bad shape");
    const int F = 1 + (L + 2 * pad - n_fft) / hop;
    TTTS_CHECK_ARG(F >= 1 && F == n_frames, "stft backward: frame count mismatch (expected %d, got %d)", F, n_frames);
    const size_t smem = ((size_t)3 * n_fft + 3 * (n_fft / 2 + 1) + n_mels) * sizeof(float);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "stft backward: n_fft %d too large for shared memory", n_fft);
#ifndef TTTS_HOST_EMU
    static size_t attr = 48 * 1024;
    if (smem > attr) { TTTS_CUDA(cudaFuncSetAttribute(stft_mel_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
#endif
    StftBwdParams p;
    p.wav = wav; p.window = window; p.dlogmel = dlogmel; p.dwav = dwav; p.band_lo = band_lo; p.band_off = band_off; p.band_w = band_w;
    p.B = B; p.L = L; p.n_fft = n_fft; p.hop = hop; p.pad = pad; p.n_mels = n_mels; p.n_frames = n_frames; p.eps_inside = eps_inside; p.log_floor = log_floor;
    TTTS_CUDA(launch_plain(stft_mel_bwd_kernel, dim3(n_frames, B), dim3(256), smem, (cudaStream_t)stream, p));
    TTTS_LAUNCH_CHECK("stft_mel_bwd");
    return TTTS_OK;
}
