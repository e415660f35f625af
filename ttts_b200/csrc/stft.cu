// Framed real FFT + magnitude + (sparse) mel filterbank + log, one pass over the waveform:
//   spectrogram_torch      ttts/utils/data_utils.py:52-87   reflect-pad, hann(2048), stft(hop 640), sqrt(re^2+im^2+1e-6)
//   spec_to_mel_torch      ttts/utils/data_utils.py:90-103  Slaney mel basis (128 x 1025) @ spec, log(clamp(., 1e-5))
//   mel_spectrogram_torch  ttts/utils/data_utils.py:106-156 (the two fused)
//   MelSpectrogramFeatures ttts/vocoder/feature_extractors.py:28-49 (n_fft 1024, hop 256, 100 HTK mels, center, log(clip(.,1e-7)))
//
// One CTA transforms FPB = 4 consecutive frames of one clip: the windowed, reflect-padded frame is packed as n_fft/2
// complex points in shared memory, transformed with an in-place radix-2 FFT (twiddles from a table computed in double on
// the host), unpacked to the n_fft/2+1 real-FFT bins, and the magnitudes stay in shared memory for the mel stage.  The
// mel basis is stored sparsely (each band = one contiguous run of bins): ~4 kFLOP per frame instead of 262 kFLOP dense.
// Per frame: read hop*4 B of new samples, write (n_fft/2+1)*4 B of spectrogram (when requested) + n_mels*4 B -> HBM-bound.
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int STFT_FPB = 4;
constexpr int STFT_THREADS = 256;

TTTS_DEVICE int reflect_index(int s, int L) {
    if (s < 0) s = -s;
    if (s >= L) s = 2 * (L - 1) - s;
    return min(max(s, 0), L - 1);
}

// tw[k] = exp(-2 pi i k / n_fft), k = 0 .. n_fft/2
__global__ void __launch_bounds__(STFT_THREADS) stft_mel_kernel(const float* __restrict__ wav, int L, int n_fft, int log2_half, int hop, int pad,
                                                                const float* __restrict__ window, const float2* __restrict__ tw, float eps_inside,
                                                                int F, float* __restrict__ spec_out, int n_mels, const int* __restrict__ band_lo,
                                                                const int* __restrict__ band_off, const float* __restrict__ band_w, float log_floor,
                                                                float* __restrict__ mel_out) {
    extern __shared__ float2 st_smem[];
    const int N2 = n_fft >> 1;
    const int bins = N2 + 1;
    float2* z = st_smem;                                             // [FPB][N2]
    float* mag = reinterpret_cast<float*>(z + STFT_FPB * N2);         // [FPB][bins (+pad)]
    const int magld = bins + 3;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * STFT_FPB;
    const float* w = wav + (size_t)b * L;

    // ---- load + window + pack (bit-reversed order for the DIT FFT) ----
    for (int i = tid; i < STFT_FPB * N2; i += STFT_THREADS) {
        const int fr = i / N2, k = i - fr * N2;
        const int f = f0 + fr;
        float2 v = make_float2(0.f, 0.f);
        if (f < F) {
            const int s0 = f * hop - pad + 2 * k;
            v.x = w[reflect_index(s0, L)] * __ldg(window + 2 * k);
            v.y = w[reflect_index(s0 + 1, L)] * __ldg(window + 2 * k + 1);
        }
        const int kr = (int)(__brev((unsigned)k) >> (32 - log2_half));
        z[fr * N2 + kr] = v;
    }
    __syncthreads();
    // ---- radix-2 DIT FFT of size N2 on each of the FPB frames ----
    for (int s = 1; s <= log2_half; ++s) {
        const int half = 1 << (s - 1);
        const int tstride = (n_fft >> s);           // W_m^pos = exp(-2 pi i pos/m) = tw[pos * n_fft/m], m = 2^s
        for (int i = tid; i < STFT_FPB * (N2 >> 1); i += STFT_THREADS) {
            const int fr = i / (N2 >> 1), j = i - fr * (N2 >> 1);
            const int grp = j >> (s - 1), pos = j & (half - 1);
            const int i0 = fr * N2 + (grp << s) + pos, i1 = i0 + half;
            const float2 wv = __ldg(tw + pos * tstride);
            const float2 a = z[i0], c = z[i1];
            const float2 u = make_float2(wv.x * c.x - wv.y * c.y, wv.x * c.y + wv.y * c.x);
            z[i0] = make_float2(a.x + u.x, a.y + u.y);
            z[i1] = make_float2(a.x - u.x, a.y - u.y);
        }
        __syncthreads();
    }
    // ---- unpack to the real-FFT bins and take magnitudes ----
    for (int i = tid; i < STFT_FPB * bins; i += STFT_THREADS) {
        const int fr = i / bins, k = i - fr * bins;
        const float2 zk = z[fr * N2 + (k == N2 ? 0 : k)];
        const float2 zc = z[fr * N2 + ((N2 - k) & (N2 - 1))];          // Z[N2-k], Z[N2] == Z[0]
        // X[k] = (Zk + conj(Zc))/2 - i/2 * tw[k] * (Zk - conj(Zc))
        const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
        const float orr = 0.5f * (zk.x - zc.x), oi = 0.5f * (zk.y + zc.y);
        const float2 t = __ldg(tw + k);
        // -i * t * (orr + i oi) = -i * ((t.x*orr - t.y*oi) + i (t.x*oi + t.y*orr)) = (t.x*oi + t.y*orr) - i (t.x*orr - t.y*oi)
        const float re = er + (t.x * oi + t.y * orr);
        const float im = ei - (t.x * orr - t.y * oi);
        mag[fr * magld + k] = sqrtf(re * re + im * im + eps_inside);
    }
    __syncthreads();
    // ---- spectrogram out: [B, bins, F], 4 consecutive frames per bin ----
    if (spec_out) {
        float* so = spec_out + (size_t)b * bins * F;
        const bool vec = ((F & 3) == 0) && (f0 + STFT_FPB <= F);
        for (int k = tid; k < bins; k += STFT_THREADS) {
            if (vec) {
                *reinterpret_cast<float4*>(so + (size_t)k * F + f0) = make_float4(mag[k], mag[magld + k], mag[2 * magld + k], mag[3 * magld + k]);
            } else {
                for (int fr = 0; fr < STFT_FPB; ++fr) if (f0 + fr < F) so[(size_t)k * F + f0 + fr] = mag[fr * magld + k];
            }
        }
    }
    // ---- sparse mel + log ----
    if (mel_out) {
        float* mo = mel_out + (size_t)b * n_mels * F;
        for (int i = tid; i < STFT_FPB * n_mels; i += STFT_THREADS) {
            const int fr = i / n_mels, m = i - fr * n_mels;
            if (f0 + fr >= F) continue;
            const int lo = band_lo[m], o0 = band_off[m], o1 = band_off[m + 1];
            float s = 0.f;
            for (int o = o0; o < o1; ++o) s = fmaf(__ldg(band_w + o), mag[fr * magld + lo + (o - o0)], s);
            mo[(size_t)m * F + f0 + fr] = logf(fmaxf(s, log_floor));
        }
    }
}

// spec_to_mel_torch on an existing spectrogram [B, bins, F]
__global__ void __launch_bounds__(256) logmel_kernel(const float* __restrict__ spec, int bins, int F, int n_mels, const int* __restrict__ band_lo,
                                                     const int* __restrict__ band_off, const float* __restrict__ band_w, float log_floor,
                                                     float* __restrict__ mel_out) {
    const int b = blockIdx.y, m = blockIdx.x;
    const int lo = band_lo[m], o0 = band_off[m], o1 = band_off[m + 1];
    const float* sp = spec + (size_t)b * bins * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float s = 0.f;
        for (int o = o0; o < o1; ++o) s = fmaf(__ldg(band_w + o), sp[(size_t)(lo + o - o0) * F + f], s);
        mel_out[((size_t)b * n_mels + m) * F + f] = logf(fmaxf(s, log_floor));
    }
}

}  // namespace ttts

using namespace ttts;

extern "C" {

int ttts_stft_mel(const float* wav, int32_t B, int32_t L, int32_t n_fft, int32_t hop, int32_t pad, const float* window, const float* twiddle,
                  float eps_inside, float* spec_out, int32_t n_mels, const int32_t* band_lo, const int32_t* band_off, const float* band_w,
                  float log_floor, float* mel_out, int32_t n_frames, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CHECK_ARG(wav && window && twiddle, "stft: null pointer");
    TTTS_CHECK_ARG(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "stft: n_fft must be a power of two in [64, 4096]");
    TTTS_CHECK_ARG(pad < L, "stft: reflect pad %d needs a longer clip (%d samples)", pad, L);
    const int F = 1 + (L + 2 * pad - n_fft) / hop;
    TTTS_CHECK_ARG(F >= 1 && F == n_frames, "stft: frame count mismatch (expected %d, got %d)", F, n_frames);
    TTTS_CHECK_ARG(!mel_out || (band_lo && band_off && band_w && n_mels > 0), "stft: mel requested without a filterbank");
    int log2_half = 0;
    while ((1 << log2_half) < (n_fft >> 1)) ++log2_half;
    const int N2 = n_fft >> 1;
    const size_t smem = (size_t)STFT_FPB * N2 * sizeof(float2) + (size_t)STFT_FPB * (N2 + 4) * sizeof(float);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        TTTS_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    dim3 grid((F + STFT_FPB - 1) / STFT_FPB, B);
    stft_mel_kernel<<<grid, STFT_THREADS, smem, st>>>(wav, L, n_fft, log2_half, hop, pad, window, reinterpret_cast<const float2*>(twiddle), eps_inside, F,
                                                      spec_out, n_mels, band_lo, band_off, band_w, log_floor, mel_out);
    TTTS_LAUNCH_CHECK("stft_mel");
    return TTTS_OK;
}

int ttts_logmel(const float* spec, int32_t B, int32_t bins, int32_t F, int32_t n_mels, const int32_t* band_lo, const int32_t* band_off,
                const float* band_w, float log_floor, float* mel_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CHECK_ARG(spec && band_lo && band_off && band_w && mel_out, "logmel: null pointer");
    logmel_kernel<<<dim3(n_mels, B), 64, 0, st>>>(spec, bins, F, n_mels, band_lo, band_off, band_w, log_floor, mel_out);
    TTTS_LAUNCH_CHECK("logmel");
    return TTTS_OK;
}
}
