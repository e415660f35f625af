// OFF BY DEFAULT (TTTS_CONV_SPLIT=1) UNTIL IT HAS RUN ON HARDWARE: written after the round's GPU budget was spent; validated on the CPU
// emulation of this source (tests/emu, tests/test_emu_conv_split_cpu.py) against torch.
//
// Split-reduction form of the pipelined implicit-GEMM convolution (conv1d.cu: conv1d_igemm_pipe_kernel<32>) for the layers that are
// LATENCY-bound, not throughput-bound: the 16 WN layers of PosteriorAudioEncoder (192 -> 384 channels, kernel 5, 36 frames per clip:
// 2 304 positions at B = 64, ttts/vqvae/modules.py:136-222) and the level-2 / level-3 ResBlock convolutions.  Their grids are 216 - 432
// CTAs of 4 warps -- one wave, 1 - 3 CTAs per SM -- and every CTA walks a reduction of Cin*K = 960 - 1 408 rows as 60 - 88 dependent
// 16-row chunks, so a layer costs 60 - 90 us however few FLOPs it has (0.85 GFLOP), and 115 of them in sequence are ~2/3 of an encode
// (profiles/r1v_launches_vqenc.csv: grids (36,6) / (36,12) / (144,3) / (72,4)).
//
// Here G groups of 4 warps share one output tile [32 channels x 64 positions]: group g runs the SAME 4-stage cp.async pipeline over
// its contiguous quarter of the chunks (own shared-memory ring), all groups in lockstep; the G partial register tiles are then summed
// through shared memory in group order by group 0, which applies the unchanged fused epilogue (bias / Mish / residual / scale / mask /
// GLU / WN gate).  Same CTAs, G x the warps in flight per SM, 1/G of the dependent chain.  Deterministic; NOT bit-identical to the
// single-group kernel (the sum over r is associated differently).
#include <stdlib.h>
#ifdef TTTS_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"
#define TTTS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif
#include "conv_params.h"

namespace ttts {

template <int CO_T, int G>
struct SplitConv {
    static constexpr int NT = CO_T * 4;                          // threads per group: (CO_T / 4) x 16
    static constexpr int LDA = CO_T + 4, LDB = IG_P + 4;
    static constexpr int NB = IG_R * IG_P / NT;                  // B-tile elements per thread
    static constexpr int S = IG_STAGES;
    static constexpr int A_ST = IG_R * LDA, B_ST = IG_R * LDB;   // floats per stage
    static constexpr int RING_F = S * (A_ST + B_ST);             // floats per group
    static constexpr int RED_F = (G - 1) * 16 * NT;              // partial tiles of groups 1 .. G-1, [g-1][i*4+j][tid]
    static constexpr int GATE_F = CO_T * (IG_P + 1);             // gated epilogue exchange, placed after the partials
    static_assert(RED_F + GATE_F <= G * RING_F, "reduction buffers alias the rings");
    static constexpr size_t kSmem = (size_t)G * RING_F * sizeof(float);
};

template <int CO_T, int G>
__global__ void __launch_bounds__(CO_T * 4 * G, G == 4 ? 2 : 3) conv1d_igemm_split_kernel(const ConvParams p) {
    using C = SplitConv<CO_T, G>;
    constexpr int NT = C::NT, LDA = C::LDA, LDB = C::LDB, NB = C::NB, S = C::S, A_ST = C::A_ST, B_ST = C::B_ST;
    TTTS_DYN_SMEM(float, sp_smem);
    const int grp = threadIdx.x / NT, tid = threadIdx.x - grp * NT;
    float* const sA = sp_smem + grp * C::RING_F;   // [S][IG_R][LDA]
    float* const sB = sA + S * A_ST;               // [S][IG_R][LDB]
    const int gated = (p.post == 1 || p.post == 3);
    const int Chalf = p.Cout >> 1;
    const int tx = tid & 15, ty = tid >> 4;
    const int p0 = blockIdx.x * IG_P;
    const int co0 = blockIdx.y * (gated ? CO_T / 2 : CO_T);
    const int R = p.Cin * p.K;
    const int Ptot = p.B * p.Tout;
    const int nchunks = (R + IG_R - 1) / IG_R;
    const int per = (nchunks + G - 1) / G;                       // iterations of every group (lockstep)
    const int c_begin = grp * per;
    const int n_local = max(0, min(nchunks, c_begin + per) - c_begin);

    // ---- B loader: this thread always loads column cb (one output position), rows rb0 + (NT/64)*i
    const int cb = tid & 63, rb0 = tid >> 6;
    const int posb = p0 + cb;
    const bool pos_ok = posb < Ptot;
    const int bb = pos_ok ? posb / p.Tout : 0;
    const int tb = pos_ok ? posb - bb * p.Tout : 0;
    const float* xb = p.x + (size_t)bb * p.Cin * p.Tin;
    const int ti0 = tb * p.stride - p.pad;
    // ---- A loader: row (output channel) ca, r-columns ra4..ra4+3
    const int ca = tid >> 2, ra4 = (tid & 3) * 4;
    int coa; bool coa_ok;
    if (gated) { const int cl = ca < CO_T / 2 ? ca : ca - CO_T / 2; coa_ok = (co0 + cl) < Chalf; coa = (ca < CO_T / 2 ? 0 : Chalf) + co0 + cl; }
    else { coa = co0 + ca; coa_ok = coa < p.Cout; }
    const float* wa = p.w + (size_t)coa * R;

    // (ci, k) of each of this thread's B elements at the group's first chunk, advanced by IG_R rows per issued chunk
    int bci[NB], bk[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) { const int rr = c_begin * IG_R + rb0 + (NT / 64) * i; bci[i] = rr / p.K; bk[i] = rr - bci[i] * p.K; }
    const int c16 = IG_R / p.K, k16 = IG_R - c16 * p.K;
    auto issue = [&](int lc) {                                   // local chunk lc -> stage lc % S
        const int r0 = (c_begin + lc) * IG_R;
        float* a = sA + (lc % S) * A_ST;
        float* b = sB + (lc % S) * B_ST;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r0 + ra4 + i;
            const bool ok = coa_ok && rr < R;
            cp_async4(&a[(ra4 + i) * LDA + ca], ok ? wa + rr : p.w, ok);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int rr = r0 + rb0 + (NT / 64) * i;
            const int ti = ti0 + bk[i] * p.dil;
            const bool ok = pos_ok && rr < R && ti >= 0 && ti < p.Tin;
            cp_async4(&b[(rb0 + (NT / 64) * i) * LDB + cb], ok ? xb + (size_t)bci[i] * p.Tin + ti : p.x, ok);
            bk[i] += k16; bci[i] += c16;
            if (bk[i] >= p.K) { bk[i] -= p.K; ++bci[i]; }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < S - 1; ++s) { if (s < n_local) issue(s); cp_async_commit(); }
    const bool lrelu = p.pre_lrelu != 0;
    for (int lc = 0; lc < per; ++lc) {
        cp_async_wait<S - 2>();                  // this group's chunk lc has landed
        const bool live = lc < n_local;
        if (lrelu && live) {                     // leaky ReLU once per element, by the thread that copied it
            float* bw = sB + (lc % S) * B_ST;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                float* e = &bw[(rb0 + (NT / 64) * i) * LDB + cb];
                const float v = *e;
                *e = v > 0.f ? v : 0.1f * v;
            }
        }
        __syncthreads();                         // ... for every thread; everybody is done computing on stage (lc - 1) % S
        if (lc + S - 1 < n_local) issue(lc + S - 1);
        cp_async_commit();
        if (live) {
            const float* a = sA + (lc % S) * A_ST;
            const float* b = sB + (lc % S) * B_ST;
#pragma unroll
            for (int r = 0; r < IG_R; ++r) {
                const float4 av = *reinterpret_cast<const float4*>(&a[r * LDA + ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&b[r * LDB + tx * 4]);
                const float a4[4] = {av.x, av.y, av.z, av.w};
                const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
            }
        }
    }

    // ---------------- sum the G partial tiles in group order (the rings are dead after the barrier) ----------------
    float* const red = sp_smem;                  // [G-1][16][NT]
    __syncthreads();
    if (grp > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[((grp - 1) * 16 + i * 4 + j) * NT + tid] = acc[i][j];
    }
    __syncthreads();
    const bool lead = grp == 0;
    if (lead) {
        for (int g = 1; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += red[((g - 1) * 16 + i * 4 + j) * NT + tid];
    }

    // ---------------- epilogue (conv1d_igemm_pipe_kernel's, by group 0; the gated read-out by all threads) ----------------
    if (!gated) {
        if (!lead) return;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int pos = p0 + tx * 4 + j;
            if (pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            const float mk = p.mask ? p.mask[(size_t)b * p.Tout + t] : 1.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = co0 + ty * 4 + i;
                if (co >= p.Cout) continue;
                float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
                if (p.post == 2) v = mish_f(v);
                const size_t o = ((size_t)b * p.Cout + co) * p.Tout + t;
                if (p.resid) v += p.resid[o];
                v *= p.out_scale;
                if (p.mask) v *= mk;
                p.y[o] = p.accumulate ? p.y[o] + v : v;
            }
        }
    } else {
        // rows < CO_T/2 of the tile are "a" channels, rows >= CO_T/2 the matching "b" channels -> exchange through shared memory
        float* const sgate = sp_smem + C::RED_F;                 // [CO_T][IG_P + 1], does not overlap the partials
        if (lead) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sgate[(ty * 4 + i) * (IG_P + 1) + tx * 4 + j] = acc[i][j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (CO_T / 2) * IG_P; i += NT * G) {
            const int cl = i / IG_P, pl = i - cl * IG_P;
            const int c = co0 + cl, pos = p0 + pl;
            if (c >= Chalf || pos >= Ptot) continue;
            const int b = pos / p.Tout, t = pos - b * p.Tout;
            float a = sgate[cl * (IG_P + 1) + pl] + (p.bias ? __ldg(p.bias + c) : 0.f);
            float g = sgate[(cl + CO_T / 2) * (IG_P + 1) + pl] + (p.bias ? __ldg(p.bias + Chalf + c) : 0.f);
            float v;
            if (p.post == 1) {
                v = a * (1.f / (1.f + expf(-g)));                                         // GLU
            } else {
                if (p.cond) { a += p.cond[(size_t)b * p.cond_ld + c]; g += p.cond[(size_t)b * p.cond_ld + Chalf + c]; }
                v = tanhf(a) * (1.f / (1.f + expf(-g)));                                  // WN gate
            }
            const size_t o = ((size_t)b * Chalf + c) * p.Tout + t;
            if (p.resid) v += p.resid[o];
            v *= p.out_scale;
            if (p.mask) v *= p.mask[(size_t)b * p.Tout + t];
            p.y[o] = p.accumulate ? p.y[o] + v : v;
        }
    }
}

template <int G>
static int conv1d_split_launch(const ConvParams& p, dim3 grid, cudaStream_t st) {
    using C = SplitConv<32, G>;
    static bool attr = false;
    if (!attr) {
        TTTS_CUDA(cudaFuncSetAttribute(conv1d_igemm_split_kernel<32, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
        attr = true;
    }
    TTTS_CUDA(launch_plain(conv1d_igemm_split_kernel<32, G>, grid, dim3(C::NT * G), C::kSmem, st, p));
    TTTS_LAUNCH_CHECK("conv1d_igemm_split");
    return TTTS_OK;
}

// Called by ttts_conv1d_f32 for a layer it would give to conv1d_igemm_pipe_kernel<32> on `grid`.  -1: not taken (switch off, the grid
// already fills the machine, or the reduction is short) -- the caller goes on to the single-group kernel.  force_groups (tests): 2 or 4.
int conv1d_split_try(const ConvParams& p, dim3 grid, int force_groups, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_CONV_SPLIT"); on = (e && e[0] == '1') ? 1 : 0; }
    const int nchunks = (p.Cin * p.K + IG_R - 1) / IG_R;
    int G = force_groups;
    if (G == 0) {
        if (!on) return -1;
        // all CTAs of the layer resident at once: 2 per SM with 4 groups (64 registers x 512 threads, 104 KB), 3 per SM with 2 groups
        const long long ctas = (long long)grid.x * grid.y;
        if (nchunks < 16) return -1;                             // nothing to split
        if (ctas <= 2ll * num_sms() && nchunks >= 32) G = 4;
        else if (ctas <= 3ll * num_sms()) G = 2;
        else return -1;                                          // enough CTAs in flight already
    }
    if (G == 4) return conv1d_split_launch<4>(p, grid, st);
    if (G == 2) return conv1d_split_launch<2>(p, grid, st);
    return -1;
}

}  // namespace ttts
