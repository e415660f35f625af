// EXPERIMENTAL, OFF BY DEFAULT (TTTS_CONV_TC=1), NOT YET RUN ON HARDWARE -- round-2 work in progress, see DESIGN.md section 7.
//
// The stride-1 ResBlock1 convolutions of the encoder's waveform branch (32 / 64 channels, kernel 3 / 7 / 11, dilation 1 / 3 / 5) on the
// 5th-gen tensor cores with SPLIT bf16 operands: x = hi + lo, w = hi + lo (hi = bf16(v), lo = bf16(v - hi)),
//     y ~= hi_x * hi_w + hi_x * lo_w + lo_x * hi_w          (fp32 accumulation in TMEM, ~2^-16 relative per product)
// which the CPU study tools/split_bf16_conv_study.py shows to keep the whole encoder within 1.7e-5 of the reference (golden tolerance
// 5e-4, no code flips).  Both CUDA-core kernels (conv1d.cu) are issue-bound at ~20 TFLOP/s on these layers.
//
// One CTA = 128 consecutive frames of one clip (MMA M) x all C output channels (MMA N), reduction index r = k * C + ci in blocks of 64:
//   warps 1-4 (128 threads, thread = frame t = TMEM lane) stage the fp32 input window [C][128 + (K-1) DIL] once (leaky ReLU applied), then
//   per k-block build -- the shifted windows cannot be addressed in place, a UMMA descriptor cannot start k*DIL elements into a swizzled
//   row -- the im2col tile [128 t][64 r] as bf16 hi and lo in the 128B-swizzle K-major layout (exactly how the attention kernels write P),
//   and the weight tile [C co][64 r] hi / lo the same way; warp 0 issues 12 tcgen05.mma per k-block (3 products x 4 k-steps of 16) into a
//   [128 x C] fp32 accumulator in TMEM; double-buffered by two full/empty mbarrier pairs.  Epilogue: tcgen05.ld (one frame per lane),
//   + bias, + residual, * scale, coalesced stores along t.  Accumulation order differs from the fp32 kernels: NOT bit-identical to them.
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

struct ConvTcParams {
    const float* x; const float* w; const float* bias; float* y;
    int B, T, pad;
    int pre_lrelu;
    const float* resid;
    float out_scale;
    int accumulate;
};

TTTS_DEVICE void ctc_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 16-byte chunk c16 (0..7) of row r inside a [rows x 64 bf16] K-major SWIZZLE_128B tile
TTTS_DEVICE void ctc_store_chunk(uint32_t base, int r, int c16, uint4 v) {
    const uint32_t addr = base + r * 128 + (((c16 ^ r) & 7) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TTTS_DEVICE void ctc_split8(const float (&v)[8], uint4& hi, uint4& lo) {
    float h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { h[e] = bf16_round(v[e]); l[e] = v[e] - h[e]; }
    hi = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    lo = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
}

template <int C, int K, int DIL>
struct ConvTc {
    static constexpr int TAPS = 64 / C;                        // kernel taps per 64-wide k-block (C = 32: 2, C = 64: 1)
    static constexpr int NKB = (K + TAPS - 1) / TAPS;          // k-blocks
    static constexpr int W = 128 + (K - 1) * DIL;              // input window per channel
    static constexpr int WP = (W + 3) & ~3;
    static constexpr int A_TILE = 128 * 128;                   // [128 t][64 bf16]
    static constexpr int B_TILE = C * 128;                     // [C co][64 bf16]
    static constexpr int oA = 0;                               // [2 stages][hi | lo]
    static constexpr int oB = oA + 4 * A_TILE;                 // [2 stages][hi | lo]
    static constexpr int oWin = oB + 4 * B_TILE;               // fp32 [C][WP]
    static constexpr int oBar = oWin + C * WP * 4;
    static constexpr size_t kSmem = oBar + 64 + 1024;          // + alignment slack
    static constexpr int TMEM_COLS = C <= 32 ? 32 : 64;
};

template <int C, int K, int DIL>
__global__ void __launch_bounds__(160, 1) conv1d_tc_kernel(const ConvTcParams p) {
    using S = ConvTc<C, K, DIL>;
    extern __shared__ uint8_t ctc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ctc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::oBar);
    uint64_t* full = bars;            // [2] 4 arrivals (one per worker warp): tiles of the stage are written
    uint64_t* empty = bars + 2;       // [2] tcgen05.commit: the MMAs that read the stage have completed
    uint64_t* acc_full = bars + 4;    // tcgen05.commit after the last k-block
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);
    float* win = reinterpret_cast<float*>(smem + S::oWin);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, t0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_holder, S::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc_bf16(128, C, false, false);
        for (int kb = 0; kb < S::NKB; ++kb) {
            const int s = kb & 1;
            mbar_wait(&full[s], (kb >> 1) & 1);
            tc_fence_after();
            const uint32_t aHi = smem_base + S::oA + (2 * s) * S::A_TILE, aLo = aHi + S::A_TILE;
            const uint32_t bHi = smem_base + S::oB + (2 * s) * S::B_TILE, bLo = bHi + S::B_TILE;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dAh = make_smem_desc_sw128(aHi + ks * 32, 16, 1024), dAl = make_smem_desc_sw128(aLo + ks * 32, 16, 1024);
                    const uint64_t dBh = make_smem_desc_sw128(bHi + ks * 32, 16, 1024), dBl = make_smem_desc_sw128(bLo + ks * 32, 16, 1024);
                    umma_bf16(tmem_base, dAh, dBh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    umma_bf16(tmem_base, dAh, dBl, idesc, 1u);
                    umma_bf16(tmem_base, dAl, dBh, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (kb == S::NKB - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------- workers: thread = frame t (= TMEM lane) ----------------
        const int quad = warp & 3;                       // the TMEM lane quadrant this warp may read
        const int t = quad * 32 + lane;
        const int q = (warp - 1) * 32 + lane;            // 0 .. 127, used to deal out the cooperative copies
        // input window [C][W]: x[b, ci, t0 - pad + u], zero outside the clip, leaky ReLU applied once
        {
            const float* xb = p.x + (size_t)b * C * p.T;
            const int in0 = t0 - p.pad;
            for (int i = q; i < C * S::W; i += 128) {
                const int ci = i / S::W, u = i - ci * S::W;
                const int ti = in0 + u;
                float v = (ti >= 0 && ti < p.T) ? xb[(size_t)ci * p.T + ti] : 0.f;
                if (p.pre_lrelu) v = v > 0.f ? v : 0.1f * v;
                win[ci * S::WP + u] = v;
            }
        }
        ctc_bar_sync(1, 128);
        for (int kb = 0; kb < S::NKB; ++kb) {
            const int s = kb & 1;
            if (kb >= 2) mbar_wait(&empty[s], ((kb >> 1) & 1) ^ 1);     // the MMAs of k-block kb - 2 have read this stage
            const uint32_t aHi = smem_base + S::oA + (2 * s) * S::A_TILE, aLo = aHi + S::A_TILE;
            const uint32_t bHi = smem_base + S::oB + (2 * s) * S::B_TILE, bLo = bHi + S::B_TILE;
            // im2col row of frame t: column j = tt * C + ci  <->  tap k = kb * TAPS + tt
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
                const int tt = (c16 * 8) / C, ci0 = (c16 * 8) % C;
                const int k = kb * S::TAPS + tt;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (k < K) ? win[(ci0 + e) * S::WP + t + k * DIL] : 0.f;
                uint4 hi, lo;
                ctc_split8(v, hi, lo);
                ctc_store_chunk(aHi, t, c16, hi);
                ctc_store_chunk(aLo, t, c16, lo);
            }
            // weight tile [C co][64 r]: w[co][ci][k] (global layout [C][C][K]), 16-byte chunks dealt out over the 128 workers
#pragma unroll
            for (int i = 0; i < (C * 8) / 128; ++i) {
                const int id = q + 128 * i;
                const int co = id >> 3, c16 = id & 7;
                const int tt = (c16 * 8) / C, ci0 = (c16 * 8) % C;
                const int k = kb * S::TAPS + tt;
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (k < K) ? __ldg(p.w + ((size_t)co * C + ci0 + e) * K + k) : 0.f;
                uint4 hi, lo;
                ctc_split8(v, hi, lo);
                ctc_store_chunk(bHi, co, c16, hi);
                ctc_store_chunk(bLo, co, c16, lo);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        // ---------------- epilogue ----------------
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int tg = t0 + t;
#pragma unroll
        for (int c = 0; c < C / 32; ++c) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld_32x32(tmem_base + lane_off + c * 32, r);
            tmem_ld_wait();
            if (tg < p.T) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int co = c * 32 + i;
                    float v = __uint_as_float(r[i]) + (p.bias ? __ldg(p.bias + co) : 0.f);
                    const size_t o = ((size_t)b * C + co) * p.T + tg;
                    if (p.resid) v += p.resid[o];
                    v *= p.out_scale;
                    p.y[o] = p.accumulate ? p.y[o] + v : v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, S::TMEM_COLS); }
}

template <int C, int K, int DIL>
static int conv1d_tc_launch(const ConvTcParams& p, cudaStream_t st) {
    using S = ConvTc<C, K, DIL>;
    static bool attr = false;
    if (!attr) { TTTS_CUDA(cudaFuncSetAttribute(conv1d_tc_kernel<C, K, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kSmem)); attr = true; }
    dim3 grid((p.T + 127) / 128, p.B);
    conv1d_tc_kernel<C, K, DIL><<<grid, 160, S::kSmem, st>>>(p);
    TTTS_LAUNCH_CHECK("conv1d_tc");
    return TTTS_OK;
}

template <int C>
static int conv1d_tc_dispatch(const ConvTcParams& p, int K, int dil, cudaStream_t st) {
#define TTTS_CTC(KK, DD) if (K == KK && dil == DD) return conv1d_tc_launch<C, KK, DD>(p, st)
    TTTS_CTC(3, 1); TTTS_CTC(3, 3); TTTS_CTC(3, 5);
    TTTS_CTC(7, 1); TTTS_CTC(7, 3); TTTS_CTC(7, 5);
    TTTS_CTC(11, 1); TTTS_CTC(11, 3); TTTS_CTC(11, 5);
#undef TTTS_CTC
    return -1;
}

// -1: not a layer this kernel covers (or TTTS_CONV_TC != 1): the caller goes on to the fp32 kernels
int conv1d_tc_try(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int T, int Cout, int K, int stride, int dil, int pad,
                  int pre_lrelu, const float* resid, float out_scale, int accumulate, const float* mask, int post, int force, cudaStream_t st) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_CONV_TC"); on = (e && e[0] == '1') ? 1 : 0; }
    if (!(on || force) || stride != 1 || post != 0 || mask != nullptr || Cin != Cout || !(Cin == 32 || Cin == 64) || T < 128 || B > 65535) return -1;
    if (pad * 2 != dil * (K - 1)) return -1;
    ConvTcParams p;
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.B = B; p.T = T; p.pad = pad; p.pre_lrelu = pre_lrelu; p.resid = resid; p.out_scale = out_scale;
    p.accumulate = accumulate;
    return Cin == 32 ? conv1d_tc_dispatch<32>(p, K, dil, st) : conv1d_tc_dispatch<64>(p, K, dil, st);
}

}  // namespace ttts
