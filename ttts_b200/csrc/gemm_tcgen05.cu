// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage smem ring -> tcgen05.mma (cta_group::1, 128xBNx16)
//   -> fp32 accumulators in TMEM (two 128xBN accumulator stages, so the epilogue of tile i overlaps the
//   main loop of tile i+1) -> tcgen05.ld -> fused epilogue (bias / gelu_new / residual+dropout / dgelu /
//   split-K fp32 reduction).
//
// Replaces: HF Conv1D addmm (HF: pytorch_utils.py:119-123) at c_attn / c_proj / c_fc / mlp.c_proj
// (HF: modeling_gpt2.py:185,223,239-242), nn.Linear heads (ttts/gpt/model.py:348-349,432-437) and the
// autograd dgrad/wgrad GEMMs of all of them.
//
// Warp roles (256 threads): warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane) + TMEM
// alloc/dealloc, warps 2-3 idle, warps 4-7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace ttts {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;


template <int BN>
struct GemmSmem {
    static constexpr int kStages = (BN == 256) ? 4 : 6;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierOffset = kStages * kStageBytes;
    static constexpr int kTotalBytes = kBarrierOffset + 256 + 1024;  // + alignment slack
};

TTTS_DEVICE void decode_item(const GemmParams& p, int item, int& m_blk, int& n_blk, int& split) {
    int tiles = p.num_m_blocks * p.num_n_blocks;
    split = item / tiles;
    int t = item - split * tiles;
    int group_size = p.group_m * p.num_n_blocks;
    int g = t / group_size;
    int r = t - g * group_size;
    int m_first = g * p.group_m;
    int gm = min(p.group_m, p.num_m_blocks - m_first);
    n_blk = r / gm;
    m_blk = m_first + (r - n_blk * gm);
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using S = GemmSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarrierOffset);
    uint64_t* empty_bar = full_bar + S::kStages;
    uint64_t* tfull_bar = empty_bar + S::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_items = p.num_m_blocks * p.num_n_blocks * p.split_k;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < S::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    // role warps: all 32 lanes run the loop, one elected lane issues the TMA / tcgen05 instructions (common.cuh elect_one)
    if (warp == 0) {
        // ================= TMA producer =================
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            int m_blk, n_blk, split;
            decode_item(p, item, m_blk, n_blk, split);
            const int m0 = m_blk * BM, n0 = n_blk * BN;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * S::kStageBytes;
                uint8_t* sb = sa + S::kABytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, kb * BK);
                    } else {
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, kb * BK);
                    } else {
                        tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
                    }
                }
                __syncwarp();
                if (++stage == S::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
        constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4);
        constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
        const uint32_t smem_base = smem_u32(smem);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            int m_blk, n_blk, split;
            decode_item(p, item, m_blk, n_blk, split);
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * S::kStageBytes;
                const uint32_t sb = sa + S::kABytes;
                const uint64_t adesc0 = A_MN ? make_smem_desc_sw128(sa, BK * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
                const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(sb, BK * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(tmem_d, adesc0 + k * kStepA, bdesc0 + k * kStepB, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == S::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull_bar[as]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;
        int it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            int m_blk, n_blk, split;
            decode_item(p, item, m_blk, n_blk, split);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int row = m_blk * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                __syncwarp();
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                gemm_epilogue_chunk(p, row, n_blk * BN + c * 32, r);
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 2 * BN); }
}

template <bool A_MN, bool B_MN, int BN>
static int launch_gemm(const ttts_gemm_args& a, const GemmParams& p, int grid, cudaStream_t stream) {
    using S = GemmSmem<BN>;
    CUtensorMap tmA, tmB;
    int rc;
    if (A_MN) rc = make_tmap_2d(&tmA, a.A, 2, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, 64, BK, true);
    else      rc = make_tmap_2d(&tmA, a.A, 2, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, BK, BM, true);
    if (rc) return rc;
    if (B_MN) rc = make_tmap_2d(&tmB, a.B, 2, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, 64, BK, true);
    else      rc = make_tmap_2d(&tmB, a.B, 2, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, BK, BN, true);
    if (rc) return rc;
    auto kern = gemm_bf16_kernel<A_MN, B_MN, BN>;
    static bool attr_set = false;  // per template instantiation
    if (!attr_set) {
        TTTS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotalBytes));
        attr_set = true;
    }
    prof_gemm_begin(stream, 2.0 * (double)a.M * (double)a.N * (double)a.K);
    kern<<<grid, GEMM_THREADS, S::kTotalBytes, stream>>>(tmA, tmB, p);
    prof_gemm_end(stream);
    TTTS_LAUNCH_CHECK("gemm_bf16_kernel");
    return TTTS_OK;
}

// The CTA-pair kernel (gemm2_tcgen05.cu, 256x256 tile per SM pair) is the default: it moves 1/3 less operand traffic L2 -> SM per
// FLOP than a 128x256 single-CTA tile, and L2 -> SM bandwidth is what bounds these GEMMs (ncu: 5.8 kB/clk chip-wide against a
// ~6.3 kB/clk cap, profiles/r1_notes.md).  TTTS_GEMM_2CTA=0 selects the single-CTA variant of the same kernel (A/B testing).
bool use_2cta(int M, int N) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_GEMM_2CTA"); on = (e && e[0] == '0') ? 0 : 1; }
    return on && N > 128 && M > 128;
}

bool use_legacy_gemm() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_GEMM_LEGACY"); on = (e && e[0] == '1') ? 1 : 0; }
    return on != 0;
}

int gemm_bf16(const ttts_gemm_args& a, cudaStream_t stream) {
    TTTS_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "gemm: bad shape %d %d %d", a.M, a.N, a.K);
    TTTS_CHECK_ARG(a.A && a.B && a.out, "gemm: null pointer");
    TTTS_CHECK_ARG(a.epi >= 0 && a.epi <= TTTS_EPI_F32, "gemm: bad epilogue %d", a.epi);
    TTTS_CHECK_ARG(a.split_k <= 1 || a.epi == TTTS_EPI_F32_ADD, "gemm: split_k needs TTTS_EPI_F32_ADD");
    TTTS_CHECK_ARG(a.epi != TTTS_EPI_RESID || a.aux, "gemm: RESID needs aux");
    TTTS_CHECK_ARG(a.epi != TTTS_EPI_DGELU || a.aux, "gemm: DGELU needs aux");
    // vectorised epilogue accesses need 16B-aligned rows
    const bool out_f32 = (a.epi == TTTS_EPI_RESID || a.epi == TTTS_EPI_F32_ADD || a.epi == TTTS_EPI_F32);
    TTTS_CHECK_ARG((a.ldo * (out_f32 ? 4 : 2)) % 16 == 0 && ((uintptr_t)a.out & 15) == 0, "gemm: out not 16B aligned (ldo=%d)", a.ldo);
    if (a.epi == TTTS_EPI_RESID) TTTS_CHECK_ARG((a.ldaux * 4) % 16 == 0 && ((uintptr_t)a.aux & 15) == 0, "gemm: resid not aligned");
    if (a.epi == TTTS_EPI_DGELU) TTTS_CHECK_ARG((a.ldaux * 2) % 16 == 0 && ((uintptr_t)a.aux & 15) == 0, "gemm: aux not aligned");
    if (a.epi == TTTS_EPI_GELU && a.aux_out) TTTS_CHECK_ARG((a.ldaux_out * 2) % 16 == 0 && ((uintptr_t)a.aux_out & 15) == 0, "gemm: aux_out not aligned");
    if (a.bias) TTTS_CHECK_ARG(((uintptr_t)a.bias & 15) == 0, "gemm: bias not 16B aligned");

    // N > 128: the pipelined-epilogue kernel (gemm2_tcgen05.cu), CTA pair by default, single CTA with TTTS_GEMM_2CTA=0;
    // TTTS_GEMM_LEGACY=1 keeps everything on the simple kernel below (A/B testing).
    if (a.N > 128 && !use_legacy_gemm()) return gemm2_bf16(a, use_2cta(a.M, a.N), stream);
    const int BN = (a.N > 128) ? 256 : 128;
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.num_m_blocks = (a.M + BM - 1) / BM;
    p.num_n_blocks = (a.N + BN - 1) / BN;
    const int sms = num_sms();
    p.group_m = sms / p.num_n_blocks;
    if (p.group_m < 1) p.group_m = 1;
    if (p.group_m > p.num_m_blocks) p.group_m = p.num_m_blocks;
    p.num_k_blocks = (a.K + BK - 1) / BK;
    int split = a.split_k < 1 ? 1 : a.split_k;
    if (split > p.num_k_blocks) split = p.num_k_blocks;
    p.kb_per_split = (p.num_k_blocks + split - 1) / split;
    p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.epi = a.epi;
    p.out = a.out; p.ldo = a.ldo; p.bias = a.bias;
    p.aux = a.aux; p.ldaux = a.ldaux; p.aux_out = a.aux_out; p.ldaux_out = a.ldaux_out;
    p.drop_thresh16 = a.drop_thresh16; p.drop_scale = a.drop_scale; p.drop_seed = a.drop_seed;
    p.a3d = p.b3d = 0; p.quad = 0; p.l2pf = 0;
    const int items = p.num_m_blocks * p.num_n_blocks * p.split_k;
    const int grid = items < sms ? items : sms;

#define TTTS_GEMM_DISPATCH(AMN, BMN)                                                      \
    (BN == 256 ? launch_gemm<AMN, BMN, 256>(a, p, grid, stream) : launch_gemm<AMN, BMN, 128>(a, p, grid, stream))
    if (!a.a_mn && !a.b_mn) return TTTS_GEMM_DISPATCH(false, false);
    if (!a.a_mn && a.b_mn) return TTTS_GEMM_DISPATCH(false, true);
    if (a.a_mn && a.b_mn) return TTTS_GEMM_DISPATCH(true, true);
    return TTTS_GEMM_DISPATCH(true, false);
#undef TTTS_GEMM_DISPATCH
}

// Pick the split-K factor (1..16) that fills the SMs best for a weight-gradient GEMM.
int pick_split_k(int M, int N, int K) {
    if (N > 128 && !use_legacy_gemm()) return pick_split_k2(M, N, K, use_2cta(M, N));
    const int BN = (N > 128) ? 256 : 128;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int kblocks = (K + BK - 1) / BK;
    const int sms = num_sms();
    int best = 1; double best_eff = -1.0;
    for (int s = 1; s <= 16 && s <= kblocks; ++s) {
        if (kblocks / s < 8 && s > 1) break;
        long items = (long)tiles * s;
        long waves = (items + sms - 1) / sms;
        double eff = (double)items / (double)(waves * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

}  // namespace ttts

extern "C" int ttts_gemm_bf16(const ttts_gemm_args* args, void* stream) {
    if (!args) { ttts::set_error("gemm: null args"); return TTTS_ERR_INVALID; }
    return ttts::gemm_bf16(*args, (cudaStream_t)stream);
}
