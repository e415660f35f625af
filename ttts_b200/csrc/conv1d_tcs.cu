// Stride-1 "same" Conv1d of the VQ-VAE encoder (ResBlock1 / WN layers: 32 ... 192 channels, kernel 1 ... 11, dilation 1 / 3 / 5;
// ttts/vqvae/modules.py:136-318) on the 5th-gen tensor cores with SPLIT bf16 operands and NO im2col:
//
//     x = xh + xl, w = wh + wl   (xh = bf16(x), xl = bf16(x - xh))        y ~= xh wh + xh wl + xl wh     (fp32 accumulation in TMEM)
//
// which keeps the whole encoder within ~1.5e-5 of the fp32 reference and all golden codes (tools/split_bf16_conv_study.py).
//
// The first tensor-core draft (conv1d_tc.cu, round 1, measured in round 2: 8.9 ms per 64-clip encode against 7.9 ms for the fp32
// kernels) rebuilt an im2col tile [128 t][64 (tap, ci)] per k-block in shared memory, so every input sample was converted, split and
// stored K times by 128 threads with one CTA per SM: the transform, not the MMAs, bounded it.  Here a convolution tap is a ROW SHIFT:
//
//   * the CTA stages its input window ONCE, channel-last: Xs[row][ci] (hi and lo copies), row = position in a packed row space, 64
//     channels = 128 bytes per row = one SWIZZLE_128B atom column (Cin > 64: several atoms side by side).  Tap k of the convolution
//     reads rows [k dil, k dil + 128) of that window: the A operand of tap k is the SAME tile with its start address advanced by
//     k dil rows (128 B each).  Measured on B200 (profiles/r2d_pytest_tcs_*.log): the 128-byte swizzle is a function of the ABSOLUTE
//     shared-memory address bits, so the unaligned start needs NO base offset in the descriptor (with the offset set to
//     (addr >> 7) & 7, as the wgmma-era documentation suggests for starts off the 1024-byte pattern, every result is wrong).
//   * the weights are pre-split ONCE per layer into [tap][ci atom][hi | lo][co][64 ci] bf16 (conv_tcs_prep_weights) and streamed by
//     TMA, 128B-swizzled, as K-major B operands [NT co][64 ci] through a ring of stages;
//   * D[128 rows][NT co] accumulates over taps x ci atoms x 4 k-steps x 3 products in TMEM; epilogue: one row (= one output frame)
//     per thread, + bias, + residual, * scale, (* mask), coalesced along t.
//
// Packed row space: clip b occupies rows [b P, b P + T), P = T + pad; the `pad` rows between clips are zero in the window, so no tap
// ever reads a neighbouring clip and short clips (T = 36 frames at the deep levels) still fill 128-row MMA tiles.
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int TCS_MAX_ROWS = 184;                        // 128 + 2 * 25 (kernel 11, dilation 5) rounded up to 8 rows
constexpr int TCS_MAX_STAGES = 8;                       // weight-tile ring: as many stages as fit (small tiles are latency-, not bandwidth-bound)
constexpr int TCS_WORKERS = 8;                           // worker warps: stage the window (thread = window row), then the epilogue (TMEM lane
                                                         // quadrant = warp & 3, the two warps of a quadrant take alternate 16-column chunks)
constexpr int TCS_THREADS = (TCS_WORKERS + 2) * 32;      // + warp 8: MMA issuer, warp 9: weight TMA

struct ConvTcsParams {
    const float* x; const float* bias; float* y; const float* resid; const float* mask;
    int B, Cin, Cout, T, K, dil, pad;
    int P;                 // row pitch of a clip in the packed row space = T + pad
    int rows;              // B * P
    int A;                 // ci atoms = ceil(Cin / 64)
    int NT;                // output channels per CTA
    int wrows;             // window rows = 128 + 2 * pad
    int arows;             // rows of an atom column in shared memory: wrows rounded up to 8 (atom stride arows * 128 B is a multiple of 1024)
    int pre_lrelu, accumulate, base_off;
    int stages;            // weight ring depth, 2 .. TCS_MAX_STAGES
    int gated;             // WN gate (post = 3): Cout = 2 H, output channel c = tanh(a_c + cond_c) * sigmoid(g_c + cond_{H + c}), a = rows [0, H), g = rows [H, 2H)
    const float* cond; int cond_ld;      // [B, 2H] conditioning (may be null)
    float out_scale;
    uint32_t p_magic;      // ceil(2^32 / P)
};

TTTS_DEVICE uint64_t tcs_desc(uint32_t saddr, int base_off) {
    uint64_t d = make_smem_desc_sw128(saddr, 16, 1024);
    if (base_off) d |= (uint64_t)((saddr >> 7) & 7u) << 49;         // swizzle phase of a start address that is not 1024-byte aligned
    return d;
}

struct TcsSmem {
    static constexpr int oBar = 0;                                    // mbarriers + tmem holder (256 B)
    static constexpr int oX = 1024;                                   // [hi | lo][A atoms][184 rows][128 B]
    static size_t x_bytes(int A, int arows) { return (size_t)2 * A * arows * 128; }
    static size_t w_stage_bytes(int NT) { return (size_t)2 * NT * 128; }
    static size_t total(int A, int arows, int NT, int stages) { return 1024 + x_bytes(A, arows) + stages * w_stage_bytes(NT) + 1024; }
};

__global__ void __launch_bounds__(TCS_THREADS, 2) conv1d_tcs_kernel(const __grid_constant__ CUtensorMap tmW, const ConvTcsParams p) {
    extern __shared__ uint8_t tcs_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tcs_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TcsSmem::oBar);
    uint64_t* w_full = bars;                       // [stages] TMA bytes landed
    uint64_t* w_empty = bars + TCS_MAX_STAGES;     // [stages] tcgen05.commit: the MMAs that read the stage are done
    uint64_t* x_full = bars + 2 * TCS_MAX_STAGES;  // [3] one per ci atom, 4 arrivals each: that atom column of the window is staged
    uint64_t* acc_full = bars + 2 * TCS_MAX_STAGES + 3;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * TCS_MAX_STAGES + 4);
    const int TCS_STAGES = p.stages;
    const uint32_t TCS_ATOM_BYTES = (uint32_t)p.arows * 128u;             // one [rows x 64 bf16] atom column
    const uint32_t sX = smem_u32(smem + TcsSmem::oX);
    const uint32_t sXlo = sX + p.A * TCS_ATOM_BYTES;
    const uint32_t sW = sX + 2 * p.A * TCS_ATOM_BYTES;
    const uint32_t w_stage = 2 * p.NT * 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g0 = blockIdx.x * 128;               // first packed output row of this tile
    const int n0 = blockIdx.y * p.NT;              // first output channel
    const int tmem_cols = p.NT <= 32 ? 32 : (p.NT <= 64 ? 64 : 128);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 3; ++a) mbar_init(&x_full[a], TCS_WORKERS);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == TCS_WORKERS) tmem_alloc(tmem_holder, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const int n_kb = p.K * p.A;                    // weight tiles: (tap, ci atom)

    if (warp == TCS_WORKERS + 1) {
        // ---------------- TMA: weight tiles [NT co][64 ci], hi then lo, one stage per (tap, atom) ----------------
        for (int it = 0; it < n_kb; ++it) {
            const int s = it % TCS_STAGES;
            const int a = it / p.K, k = it - a * p.K;       // atom-major order: the MMAs of atom 0 run while atoms 1, 2 are still being staged
            const int kb = k * p.A + a;                     // tile index in the pre-split weight tensor [tap][atom]
            mbar_wait(&w_empty[s], ((it / TCS_STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&w_full[s], w_stage);
                uint8_t* dst = smem + TcsSmem::oX + (size_t)2 * p.A * TCS_ATOM_BYTES + (size_t)s * w_stage;
                if (!p.gated) {
                    tma_load_2d(dst, &tmW, &w_full[s], 0, (kb * 2 + 0) * p.Cout + n0);
                    tma_load_2d(dst + p.NT * 128, &tmW, &w_full[s], 0, (kb * 2 + 1) * p.Cout + n0);
                } else {                                  // tile = NT / 2 `a` channels followed by the NT / 2 `g` channels of the same outputs
                    const int hN = p.NT / 2, c0 = blockIdx.y * hN, H = p.Cout / 2;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tma_load_2d(dst + h * p.NT * 128, &tmW, &w_full[s], 0, (kb * 2 + h) * p.Cout + c0);
                        tma_load_2d(dst + h * p.NT * 128 + hN * 128, &tmW, &w_full[s], 0, (kb * 2 + h) * p.Cout + H + c0);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == TCS_WORKERS) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = make_idesc_bf16(128, p.NT, false, false);
        uint32_t first = 1;
        for (int it = 0; it < n_kb; ++it) {
            const int s = it % TCS_STAGES;
            const int a = it / p.K, k = it - a * p.K;
            if (k == 0) mbar_wait(&x_full[a], 0);
            mbar_wait(&w_full[s], (it / TCS_STAGES) & 1);
            tc_fence_after();
            const uint32_t row_off = (uint32_t)(k * p.dil) * 128u;          // tap k = the window shifted by k * dil rows
            const uint32_t aHi = sX + a * TCS_ATOM_BYTES + row_off, aLo = sXlo + a * TCS_ATOM_BYTES + row_off;
            const uint32_t bHi = sW + s * w_stage, bLo = bHi + p.NT * 128;
            const int ksteps = min(4, (p.Cin - a * 64 + 15) / 16);
            if (elect_one()) {
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t dAh = tcs_desc(aHi + ks * 32, p.base_off), dAl = tcs_desc(aLo + ks * 32, p.base_off);
                    const uint64_t dBh = make_smem_desc_sw128(bHi + ks * 32, 16, 1024), dBl = make_smem_desc_sw128(bLo + ks * 32, 16, 1024);
                    umma_bf16(tmem_base, dAh, dBh, idesc, first ? 0u : 1u);
                    umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
                    umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
                    first = 0;
                }
                umma_commit(&w_empty[s]);
                if (it == n_kb - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------- workers: stage the window (channel-last, hi / lo), then the epilogue ----------------
        const int tid = threadIdx.x;                                         // 0 .. 255 = window row
        {
            const int j = tid;
            const int gi = g0 - p.pad + j;                                   // packed input row
            bool row_ok = (j < p.wrows) && gi >= 0 && gi < p.rows;
            int b_ = 0, t_ = 0;
            if (row_ok) { b_ = (int)__umulhi((uint32_t)gi, p.p_magic); if (p.P == 1) b_ = gi; t_ = gi - b_ * p.P; row_ok = t_ < p.T; }
            const float* xp = p.x + ((size_t)b_ * p.Cin) * p.T + t_;
            for (int a = 0; a < p.A; ++a) {                                  // atom by atom: the MMA warp starts on atom 0 while 1, 2 are staged
                if (j < p.arows) {
                    // four 8-channel groups per iteration: 32 independent loads in flight per thread (the loads of a channel are coalesced
                    // across the warp: consecutive lanes = consecutive frames)
                    for (int cg = a * 8; cg < a * 8 + 8; cg += 4) {
                        float v[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const int ci = cg * 8 + e;
                            v[e] = (row_ok && ci < p.Cin) ? __ldg(xp + (size_t)ci * p.T) : 0.f;
                        }
                        if (p.pre_lrelu) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) v[e] = v[e] > 0.f ? v[e] : 0.1f * v[e];
                        }
#pragma unroll
                        for (int h4 = 0; h4 < 4; ++h4) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float h0 = bf16_round(v[h4 * 8 + 2 * e]), h1 = bf16_round(v[h4 * 8 + 2 * e + 1]);
                                hi[e] = pack_bf16(h0, h1);
                                lo[e] = pack_bf16(v[h4 * 8 + 2 * e] - h0, v[h4 * 8 + 2 * e + 1] - h1);
                            }
                            const int c = cg + h4;
                            const uint32_t off = (uint32_t)(c >> 3) * TCS_ATOM_BYTES + (uint32_t)j * 128u + ((uint32_t)((c ^ j) & 7) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sX + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sXlo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]));
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&x_full[a]);
            }
        }

        // epilogue: thread = packed output row g0 + tid = TMEM lane
        const int quad = warp & 3, half = warp >> 2;                         // TMEM lane quadrant; which of the alternating 16-column chunks
        const int g = g0 + quad * 32 + lane;
        bool ok = g < p.rows;
        int b = 0, t = 0;
        if (ok) { b = (int)__umulhi((uint32_t)g, p.p_magic); if (p.P == 1) b = g; t = g - b * p.P; ok = t < p.T; }
        const float mk = (ok && p.mask) ? p.mask[(size_t)b * p.T + t] : 1.f;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        if (p.gated) {
            const int hN = p.NT / 2, H = p.Cout / 2, cbase = blockIdx.y * hN;
            const float* cnd = (ok && p.cond) ? p.cond + (size_t)b * p.cond_ld : nullptr;
            for (int c0 = half * 16; c0 < hN; c0 += 32) {
                uint32_t ra[16], rg[16];
                __syncwarp();
                tmem_ld_32x16(tmem_base + lane_off + c0, ra);
                tmem_ld_32x16(tmem_base + lane_off + hN + c0, rg);
                tmem_ld_wait();                   // .sync.aligned: must be executed by the converged warp, never inside the `ok` branch
                if (ok) {
                    float ba[16], bg[16], res[16], old[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = cbase + c0 + i;
                        const bool cok = c0 + i < hN;
                        ba[i] = cok ? ((p.bias ? __ldg(p.bias + c) : 0.f) + (cnd ? __ldg(cnd + c) : 0.f)) : 0.f;
                        bg[i] = cok ? ((p.bias ? __ldg(p.bias + H + c) : 0.f) + (cnd ? __ldg(cnd + H + c) : 0.f)) : 0.f;
                        const size_t o = ((size_t)b * H + c) * p.T + t;
                        res[i] = (p.resid && cok) ? __ldg(p.resid + o) : 0.f;
                        old[i] = (p.accumulate && cok) ? p.y[o] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = cbase + c0 + i;
                        if (c0 + i < hN) {
                            const float a = __uint_as_float(ra[i]) + ba[i], g = __uint_as_float(rg[i]) + bg[i];
                            const float v = tanhf(a) * (1.f / (1.f + __expf(-g)));
                            p.y[((size_t)b * H + c) * p.T + t] = old[i] + (v + res[i]) * (p.out_scale * mk);
                        }
                    }
                }
            }
        } else {
        for (int c0 = half * 16; c0 < p.NT; c0 += 32) {
            uint32_t r[16];
            __syncwarp();
            tmem_ld_32x16(tmem_base + lane_off + c0, r);
            tmem_ld_wait();                       // .sync.aligned: converged warp only (r2f: inside the divergent `ok` branch it deadlocked)
            if (ok) {
                // every load of the chunk is issued before the first store: written load -> store -> load per channel the compiler must keep
                // that order (y may alias resid for all it knows) and a 32-channel tile becomes a chain of 32 dependent DRAM round trips
                // (r2e launch list: 80 us for a level-0 layer whose MMAs take 2 us)
                const size_t o0 = ((size_t)b * p.Cout + n0 + c0) * p.T + t;
                float res[16], old[16], bs[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool cok = n0 + c0 + i < p.Cout;
                    res[i] = (p.resid && cok) ? __ldg(p.resid + o0 + (size_t)i * p.T) : 0.f;
                    old[i] = (p.accumulate && cok) ? p.y[o0 + (size_t)i * p.T] : 0.f;
                    bs[i] = (p.bias && cok) ? __ldg(p.bias + n0 + c0 + i) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (n0 + c0 + i < p.Cout) p.y[o0 + (size_t)i * p.T] = old[i] + (__uint_as_float(r[i]) + bs[i] + res[i]) * (p.out_scale * mk);
                }
            }
        }
        }
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == TCS_WORKERS) { __syncwarp(); tmem_dealloc(tmem_base, tmem_cols); }
}

// w [Cout][Cin][K] fp32  ->  ws [K][A][hi | lo][Cout][64] bf16 (ci >= Cin: zero)
__global__ void conv_tcs_prep_weights_kernel(const float* __restrict__ w, bf16* __restrict__ ws, int Cout, int Cin, int K, int A) {
    const size_t n = (size_t)K * A * Cout * 64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(i & 63);
        const size_t r = i >> 6;
        const int co = (int)(r % Cout);
        const int ka = (int)(r / Cout);
        const int a = ka % A, k = ka / A;
        const int ci = a * 64 + cc;
        const float v = ci < Cin ? w[((size_t)co * Cin + ci) * K + k] : 0.f;
        const float h = bf16_round(v);
        ws[(((size_t)ka * 2 + 0) * Cout + co) * 64 + cc] = __float2bfloat16_rn(h);
        ws[(((size_t)ka * 2 + 1) * Cout + co) * 64 + cc] = __float2bfloat16_rn(v - h);
    }
}

int64_t conv_tcs_weight_elems(int Cout, int Cin, int K) { return (int64_t)K * ((Cin + 63) / 64) * 2 * Cout * 64; }

int conv_tcs_prep_weights(const float* w, void* ws, int Cout, int Cin, int K, cudaStream_t st) {
    TTTS_CHECK_ARG(w && ws && Cout > 0 && Cin > 0 && K > 0, "conv_tcs_prep_weights: bad arguments");
    const int A = (Cin + 63) / 64;
    const size_t n = (size_t)K * A * Cout * 64;
    int blocks = (int)((n + 255) / 256);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    conv_tcs_prep_weights_kernel<<<blocks, 256, 0, st>>>(w, (bf16*)ws, Cout, Cin, K, A);
    TTTS_LAUNCH_CHECK("conv_tcs_prep_weights");
    return TTTS_OK;
}

bool conv_tcs_covers(int Cin, int Cout, int K, int stride, int dil, int pad, int post) {
    if (stride != 1 || !(post == 0 || (post == 3 && Cout == 384)) || K < 1 || dil < 1) return false;
    if (pad * 2 != dil * (K - 1) || 128 + 2 * pad > TCS_MAX_ROWS) return false;
    if (Cin % 8 != 0 || Cin < 16 || Cin > 192) return false;
    if (!(Cout == 32 || Cout == 64 || Cout == 96 || Cout == 128 || Cout == 192 || Cout == 384)) return false;
    return true;
}

int conv1d_tcs(const float* x, const void* ws, const float* bias, float* y, int B, int Cin, int T, int Cout, int K, int dil, int pre_lrelu,
               const float* resid, float out_scale, int accumulate, const float* mask, int post, const float* cond, int cond_ld, int base_off,
               cudaStream_t st) {
    const int pad = dil * (K - 1) / 2;
    TTTS_CHECK_ARG(x && ws && y && B > 0 && T > 0, "conv1d_tcs: bad arguments");
    TTTS_CHECK_ARG(conv_tcs_covers(Cin, Cout, K, 1, dil, pad, post), "conv1d_tcs: layer not covered (stride 1, same padding, Cin %% 8 == 0, Cin <= 192, Cout in {32, 64, 96, 128, 192, 384})");
    ConvTcsParams p = {};
    p.x = x; p.bias = bias; p.y = y; p.resid = resid; p.mask = mask;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.T = T; p.K = K; p.dil = dil; p.pad = pad;
    p.P = T + pad;
    TTTS_CHECK_ARG((long long)B * p.P < (1ll << 31), "conv1d_tcs: problem too large");
    p.rows = B * p.P;
    p.A = (Cin + 63) / 64;
    p.NT = Cout <= 128 ? Cout : 96;
    p.gated = post == 3; p.cond = cond; p.cond_ld = cond_ld;
    p.wrows = 128 + 2 * pad;
    p.arows = (p.wrows + 7) & ~7;
    p.pre_lrelu = pre_lrelu; p.accumulate = accumulate; p.base_off = base_off; p.out_scale = out_scale;
    p.p_magic = (uint32_t)((0x100000000ULL + (uint64_t)p.P - 1) / (uint64_t)p.P);
    CUtensorMap tm;
    const uint64_t wrows = (uint64_t)K * p.A * 2 * Cout;
    int rc = make_tmap_2d(&tm, ws, 2, 64, wrows, 64, 64, (uint32_t)(p.gated ? p.NT / 2 : p.NT), 1);
    if (rc) return rc;
    // ring depth: what fits next to the window, at most one stage per weight tile; shallow for the big tiles so that 2 CTAs share an SM
    const int n_kb = K * p.A;
    int stages = TCS_MAX_STAGES;
    // single-atom layers (level 0 / 1: thousands of small CTAs, DRAM-latency-bound): 2 CTAs per SM (the register file allows no more at
    // 320 threads x 96 registers); the others: as deep as fits
    while (stages > 2 && (stages > n_kb || TcsSmem::total(p.A, p.arows, p.NT, stages) > (p.A == 1 ? 110 : 227) * 1024)) --stages;
    p.stages = stages;
    const size_t smem = TcsSmem::total(p.A, p.arows, p.NT, stages);
    static size_t attr = 0;
    if (smem > attr) {
        TTTS_CUDA(cudaFuncSetAttribute(conv1d_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        // r2h capture: 70 KB per CTA and still ONE CTA per SM -- 133 registers x 320 threads (now capped by the launch bounds) and a
        // shared-memory carve-out sized for one block; ask for the full carve-out so that 2 - 3 of the small CTAs share an SM
        TTTS_CUDA(cudaFuncSetAttribute(conv1d_tcs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        attr = 227 * 1024;
    }
    TTTS_CHECK_ARG(smem <= 227 * 1024, "conv1d_tcs: shared memory");
    dim3 grid((p.rows + 127) / 128, Cout / p.NT);            // gated: Cout / NT = H / (NT / 2) tiles of NT / 2 output channels
    prof_begin(5, st, 2.0 * B * T * (double)Cin * Cout * K);
    conv1d_tcs_kernel<<<grid, TCS_THREADS, smem, st>>>(tm, p);
    prof_end(5, st);
    TTTS_LAUNCH_CHECK("conv1d_tcs");
    return TTTS_OK;
}

}  // namespace ttts

extern "C" {
int64_t ttts_conv1d_tcs_weight_elems(int32_t Cout, int32_t Cin, int32_t K) { return ttts::conv_tcs_weight_elems(Cout, Cin, K); }
int ttts_conv1d_tcs_prep_weights(const float* w, void* ws_bf16, int32_t Cout, int32_t Cin, int32_t K, void* stream) {
    return ttts::conv_tcs_prep_weights(w, ws_bf16, Cout, Cin, K, (cudaStream_t)stream);
}
int ttts_conv1d_tcs(const float* x, const void* ws_bf16, const float* bias, float* y, int32_t B, int32_t Cin, int32_t T, int32_t Cout, int32_t K,
                    int32_t dil, int32_t pre_lrelu, const float* resid, float out_scale, int32_t accumulate, const float* mask, int32_t post,
                    const float* cond, int32_t cond_ld, int32_t flags, void* stream) {
    return ttts::conv1d_tcs(x, ws_bf16, bias, y, B, Cin, T, Cout, K, dil, pre_lrelu, resid, out_scale, accumulate, mask, post, cond, cond_ld,
                            (flags & 1) ? 1 : 0, (cudaStream_t)stream);
}
}
