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

// ------------------------------------------------------------------------------------------------------------
// Register-resident mixed-radix form for the two transform sizes the reference uses (n_fft 2048 -> 1024 complex points = 16 x 16 x 4,
// n_fft 1024 -> 512 = 16 x 16 x 2); default for those sizes, TTTS_STFT_V1=1 selects the radix-2 kernel above.
// The radix-2 kernel makes 10 passes over shared memory with a barrier each and one butterfly per thread per pass: 41 M frames/s =
// 126 GB/s, 2 % of the HBM roof (r1g bench).  Here a frame is transformed by N2/16 threads holding 16 complex points each in registers:
//   pass 1: 16-point DFTs over n1 (stride N2/16), twiddle W_N2^(n2 k1)           -> smem [k1][n2]
//   pass 2: 16-point DFTs over n2a,               twiddle W_M^(n2b k2a), M = N2/16 -> smem [n2b][k2a 16 + k1]
//   pass 3: R3-point DFTs over n2b (R3 = N2/256)                                 -> smem, natural order k = k1 + 16 k2a + 256 k2b
// i.e. three exchanges instead of ten, all shared-memory accesses conflict-free by the row pitches chosen below; the 15 twiddles of a
// thread are built from 4 table entries (w, w^2, w^4, w^8).  After the transform: real-FFT unpack, magnitudes, spectrogram / sparse mel
// as in the kernel above.  What bounds it: 1024-point FFT = 51 kFLOP of mostly non-fused adds per frame against 7.2 KB of HBM traffic
// -> the fp32 pipe, not HBM.
// ------------------------------------------------------------------------------------------------------------
TTTS_DEVICE float2 cmulf(const float2 a, const float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
TTTS_DEVICE void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {       // forward: W4 = -i
    const float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
    const float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
    x0 = make_float2(s02.x + s13.x, s02.y + s13.y);
    x2 = make_float2(s02.x - s13.x, s02.y - s13.y);
    x1 = make_float2(d02.x + d13.y, d02.y - d13.x);
    x3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}
// in place, natural order in and out: A[k] = sum_n a[n] W16^(n k)
TTTS_DEVICE void dft16(float2 (&a)[16]) {
    constexpr float C = 0.92387953251128674f, S = 0.38268343236508977f, H = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(a[b], a[4 + b], a[8 + b], a[12 + b]);        // t[b][c] in a[4 c + b]
    // t[b][c] *= W16^(b c)
    a[4 * 1 + 1] = cmulf(a[4 * 1 + 1], make_float2(C, -S));
    a[4 * 2 + 1] = cmulf(a[4 * 2 + 1], make_float2(H, -H));
    a[4 * 3 + 1] = cmulf(a[4 * 3 + 1], make_float2(S, -C));
    a[4 * 1 + 2] = cmulf(a[4 * 1 + 2], make_float2(H, -H));
    a[4 * 2 + 2] = make_float2(a[4 * 2 + 2].y, -a[4 * 2 + 2].x);
    a[4 * 3 + 2] = cmulf(a[4 * 3 + 2], make_float2(-H, -H));
    a[4 * 1 + 3] = cmulf(a[4 * 1 + 3], make_float2(S, -C));
    a[4 * 2 + 3] = cmulf(a[4 * 2 + 3], make_float2(-H, -H));
    a[4 * 3 + 3] = cmulf(a[4 * 3 + 3], make_float2(-C, S));
    // A[c + 4 d] = DFT4 over b of t[b][c]; t[b][c] sits in a[4 c + b], the results go to a[4 c + d] -> one transposition at the end
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { const float2 t = a[4 * c + d]; a[4 * c + d] = a[4 * d + c]; a[4 * d + c] = t; }
}
// a[k] *= w^k, k = 1 .. 15, from w, w^2, w^4, w^8
TTTS_DEVICE void twiddle16(float2 (&a)[16], const float2 w1, const float2 w2, const float2 w4, const float2 w8) {
    const float2 w3 = cmulf(w1, w2), w5 = cmulf(w4, w1), w6 = cmulf(w4, w2), w7 = cmulf(w4, w3);
    a[1] = cmulf(a[1], w1); a[2] = cmulf(a[2], w2); a[3] = cmulf(a[3], w3); a[4] = cmulf(a[4], w4);
    a[5] = cmulf(a[5], w5); a[6] = cmulf(a[6], w6); a[7] = cmulf(a[7], w7); a[8] = cmulf(a[8], w8);
    a[9] = cmulf(a[9], cmulf(w8, w1)); a[10] = cmulf(a[10], cmulf(w8, w2)); a[11] = cmulf(a[11], cmulf(w8, w3));
    a[12] = cmulf(a[12], cmulf(w8, w4)); a[13] = cmulf(a[13], cmulf(w8, w5)); a[14] = cmulf(a[14], cmulf(w8, w6));
    a[15] = cmulf(a[15], cmulf(w8, w7));
}

template <int LOG2N2>
struct StftR16 {
    static constexpr int N2 = 1 << LOG2N2, R3 = N2 / 256, M = N2 / 16, FPB = STFT_THREADS / M;
    static constexpr int PA = M + R3;                 // pass-1 layout [k1][n2]: pass-2 thread (k1, n2b) reads bank (k1 PA + n2b) mod 32, all distinct
    static constexpr int PB = 256 + 32 / R3;          // pass-2 layout [n2b][k2a 16 + k1]: writes of a warp hit banks (n2b PB + k1) mod 32, all distinct
    static constexpr int BUF0 = 16 * PA > N2 ? 16 * PA : N2;          // floats per component
    static constexpr int BUF1 = R3 * PB > N2 + 4 ? R3 * PB : N2 + 4;  // also holds the N2 + 1 magnitudes
    static constexpr int BUF = BUF0 > BUF1 ? BUF0 : BUF1;
    static constexpr int FR = 2 * BUF;                // floats per frame: ONE re / im buffer pair, every pass reads it into registers,
                                                      // barrier, writes it back in the next layout (34 KB per CTA -> 4 CTAs per SM, register-bound)
    static constexpr size_t kSmem = (size_t)FPB * FR * sizeof(float);
};

template <int LOG2N2>
__global__ void __launch_bounds__(STFT_THREADS) stft_mel_r16_kernel(const float* __restrict__ wav, int L, int hop, int pad, const float* __restrict__ window,
                                                                    const float2* __restrict__ tw, float eps_inside, int F, float* __restrict__ spec_out,
                                                                    int n_mels, const int* __restrict__ band_lo, const int* __restrict__ band_off,
                                                                    const float* __restrict__ band_w, float log_floor, float* __restrict__ mel_out) {
    using C = StftR16<LOG2N2>;
    constexpr int N2 = C::N2, R3 = C::R3, M = C::M, FPB = C::FPB, PA = C::PA, PB = C::PB;
    constexpr int bins = N2 + 1;
    extern __shared__ float st16_smem[];
    const int tid = threadIdx.x;
    const int fr = tid / M, t = tid - fr * M;
    float* re0 = st16_smem + fr * C::FR;
    float* im0 = re0 + C::BUF;
    float* re1 = re0;                                 // same storage, next layout
    float* im1 = im0;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FPB;
    const int f = f0 + fr;
    const float* w = wav + (size_t)b * L;
    float2 a[16];

    // ---- pass 1: thread n2 = t takes z[n1 M + n2], n1 = 0 .. 15 (z[m] = windowed samples 2m, 2m + 1) ----
    {
        const int s0 = f * hop - pad;
        const bool live = f < F;
        const bool inside = live && s0 >= 0 && s0 + 2 * N2 <= L;
        const bool vec = inside && ((((size_t)b * L + s0) & 1) == 0) && ((reinterpret_cast<uintptr_t>(wav) & 7) == 0);
        // one branch around each whole batch of 16 loads: with the case distinction inside the loop every load sat in its own basic block
        // and was waited for before the next was issued (r1n profile: a third of all stall samples on the 16 window multiplies)
        if (vec) {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) a[n1] = *reinterpret_cast<const float2*>(w + s0 + 2 * (n1 * M + t));
        } else if (inside) {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) { a[n1].x = w[s0 + 2 * (n1 * M + t)]; a[n1].y = w[s0 + 2 * (n1 * M + t) + 1]; }
        } else if (live) {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) { a[n1].x = w[reflect_index(s0 + 2 * (n1 * M + t), L)]; a[n1].y = w[reflect_index(s0 + 2 * (n1 * M + t) + 1, L)]; }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) a[n1] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float2 wn = __ldg(reinterpret_cast<const float2*>(window) + n1 * M + t);
            a[n1] = make_float2(a[n1].x * wn.x, a[n1].y * wn.y);
        }
        dft16(a);
        twiddle16(a, __ldg(tw + 2 * t), __ldg(tw + 4 * t), __ldg(tw + 8 * t), __ldg(tw + 16 * t));        // W_N2^(n2 k1) = tw[2 n2 k1]
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) { re0[k1 * PA + t] = a[k1].x; im0[k1 * PA + t] = a[k1].y; }
    }
    __syncthreads();
    // ---- pass 2: thread (k1, n2b) = (t / R3, t % R3) takes y[k1][n2a R3 + n2b], n2a = 0 .. 15 ----
    {
        const int k1 = t / R3, n2b = t - k1 * R3;
#pragma unroll
        for (int n2a = 0; n2a < 16; ++n2a) a[n2a] = make_float2(re0[k1 * PA + n2a * R3 + n2b], im0[k1 * PA + n2a * R3 + n2b]);
        __syncthreads();                              // every thread has its 16 points: the buffer may be overwritten
        dft16(a);
        twiddle16(a, __ldg(tw + 32 * n2b), __ldg(tw + 64 * n2b), __ldg(tw + 128 * n2b), __ldg(tw + 256 * n2b));   // W_M^(n2b k2a) = tw[32 n2b k2a]
#pragma unroll
        for (int k2a = 0; k2a < 16; ++k2a) { re1[n2b * PB + k2a * 16 + k1] = a[k2a].x; im1[n2b * PB + k2a * 16 + k1] = a[k2a].y; }
    }
    __syncthreads();
    // ---- pass 3: R3-point DFTs over n2b for j = k2a 16 + k1 = t + M i ; Z[j + 256 k2b] -> buffer 0, natural order ----
#pragma unroll
    for (int i = 0; i < 16 / R3; ++i) {
        const int j = t + M * i;
        if (R3 == 4) {
            float2 x0 = make_float2(re1[j], im1[j]), x1 = make_float2(re1[PB + j], im1[PB + j]);
            float2 x2 = make_float2(re1[2 * PB + j], im1[2 * PB + j]), x3 = make_float2(re1[3 * PB + j], im1[3 * PB + j]);
            dft4(x0, x1, x2, x3);
            a[4 * i] = x0; a[4 * i + 1] = x1; a[4 * i + 2] = x2; a[4 * i + 3] = x3;
        } else {
            const float2 x0 = make_float2(re1[j], im1[j]), x1 = make_float2(re1[PB + j], im1[PB + j]);
            a[2 * i] = make_float2(x0.x + x1.x, x0.y + x1.y);
            a[2 * i + 1] = make_float2(x0.x - x1.x, x0.y - x1.y);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16 / R3; ++i)
#pragma unroll
        for (int k2b = 0; k2b < R3; ++k2b) { re0[t + M * i + 256 * k2b] = a[R3 * i + k2b].x; im0[t + M * i + 256 * k2b] = a[R3 * i + k2b].y; }
    __syncthreads();
    // ---- unpack to the real-FFT bins, magnitudes (bins t + M i in registers, then back into re[0 .. N2]) ----
    float mg[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        const int k = t + M * i;
        mg[i] = 0.f;
        if (k < bins) {
            const int kz = k & (N2 - 1), kc = (N2 - k) & (N2 - 1);
            const float zkx = re0[kz], zky = im0[kz], zcx = re0[kc], zcy = im0[kc];
            const float er = 0.5f * (zkx + zcx), ei = 0.5f * (zky - zcy);
            const float orr = 0.5f * (zkx - zcx), oi = 0.5f * (zky + zcy);
            const float2 tk = __ldg(tw + k);
            const float re = er + (tk.x * oi + tk.y * orr);
            const float im = ei - (tk.x * orr - tk.y * oi);
            mg[i] = sqrtf(re * re + im * im + eps_inside);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 17; ++i) { const int k = t + M * i; if (k < bins) re1[k] = mg[i]; }
    __syncthreads();
    const float* mag0 = st16_smem;                     // frame fr2: mag0 + fr2 * FR
    // ---- spectrogram out: [B, bins, F], FPB consecutive frames per bin ----
    if (spec_out) {
        float* so = spec_out + (size_t)b * bins * F;
        const bool vec4 = ((F & 3) == 0) && (f0 + FPB <= F);
        for (int k = tid; k < bins; k += STFT_THREADS) {
            if (vec4) {
#pragma unroll
                for (int q = 0; q < FPB / 4; ++q)
                    *reinterpret_cast<float4*>(so + (size_t)k * F + f0 + 4 * q) =
                        make_float4(mag0[(4 * q) * C::FR + k], mag0[(4 * q + 1) * C::FR + k], mag0[(4 * q + 2) * C::FR + k], mag0[(4 * q + 3) * C::FR + k]);
            } else {
                for (int q = 0; q < FPB; ++q) if (f0 + q < F) so[(size_t)k * F + f0 + q] = mag0[q * C::FR + k];
            }
        }
    }
    // ---- sparse mel + log ----
    if (mel_out) {
        float* mo = mel_out + (size_t)b * n_mels * F;
        for (int i = tid; i < FPB * n_mels; i += STFT_THREADS) {
            const int m = i / FPB, q = i - m * FPB;          // frames fastest: neighbouring threads write neighbouring addresses
            if (f0 + q >= F) continue;
            const int lo = band_lo[m], o0 = band_off[m], o1 = band_off[m + 1];
            const float* mg = mag0 + q * C::FR + lo;
            float s = 0.f;
            for (int o = o0; o < o1; ++o) s = fmaf(__ldg(band_w + o), mg[o - o0], s);
            mo[(size_t)m * F + f0 + q] = logf(fmaxf(s, log_floor));
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
    TTTS_CHECK_ARG(B >= 1 && B <= 65535, "stft: batch %d outside [1, 65535] (one grid row per clip)", B);
    int log2_half = 0;
    while ((1 << log2_half) < (n_fft >> 1)) ++log2_half;
    const int N2 = n_fft >> 1;
    const size_t smem = (size_t)STFT_FPB * N2 * sizeof(float2) + (size_t)STFT_FPB * (N2 + 4) * sizeof(float);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        TTTS_CUDA(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("TTTS_STFT_V1"); v1 = (e && e[0] == '1') ? 1 : 0; }
    if (!v1 && (n_fft == 2048 || n_fft == 1024) && (reinterpret_cast<uintptr_t>(window) & 7) == 0) {
        static bool attr16 = false;
        if (!attr16) {
            TTTS_CUDA(cudaFuncSetAttribute(stft_mel_r16_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftR16<10>::kSmem));
            TTTS_CUDA(cudaFuncSetAttribute(stft_mel_r16_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftR16<9>::kSmem));
            attr16 = true;
        }
        const float2* tw2 = reinterpret_cast<const float2*>(twiddle);
        if (n_fft == 2048) {
            dim3 grid16((F + StftR16<10>::FPB - 1) / StftR16<10>::FPB, B);
            stft_mel_r16_kernel<10><<<grid16, STFT_THREADS, StftR16<10>::kSmem, st>>>(wav, L, hop, pad, window, tw2, eps_inside, F, spec_out, n_mels, band_lo,
                                                                                        band_off, band_w, log_floor, mel_out);
        } else {
            dim3 grid16((F + StftR16<9>::FPB - 1) / StftR16<9>::FPB, B);
            stft_mel_r16_kernel<9><<<grid16, STFT_THREADS, StftR16<9>::kSmem, st>>>(wav, L, hop, pad, window, tw2, eps_inside, F, spec_out, n_mels, band_lo,
                                                                                      band_off, band_w, log_floor, mel_out);
        }
        TTTS_LAUNCH_CHECK("stft_mel_r16");
        return TTTS_OK;
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
