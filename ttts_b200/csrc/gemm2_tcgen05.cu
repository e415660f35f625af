// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on one TPC computes a 256x256
// output tile.  Each CTA TMA-loads its own 128 rows of A and HALF of the B tile (128 of the 256 n-rows); the leader CTA
// issues tcgen05.mma.cta_group::2 (M=256, N=256, K=16) which reads A/B from both CTAs' shared memory and writes 128x256
// fp32 accumulators into each CTA's TMEM.  Versus the 1-CTA kernel this cuts L2->SM and shared-memory operand traffic per
// FLOP by 1/3 (32 KB instead of 48 KB per 128x256x64 MAC block per SM), which is what limits the 1-CTA kernel
// (profiles/r1_notes.md).
//
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM alloc, warps 2-9 =
// epilogue (two warps per TMEM lane quadrant, splitting the 256 columns; outputs and auxiliary inputs go through the TMA unit,
// gemm_epilogue_tma.cuh).  Barriers: full[] live in the leader (count 1:
// the leader's arrive.expect_tx covers both CTAs' bytes; both CTAs' TMA complete_tx there), empty[] / tmem_full[] are
// per CTA and signalled by one multicast tcgen05.commit, tmem_empty[] lives in the leader (2 x 8 epilogue warps).
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "gemm_epilogue_tma.cuh"
#include "kernels.h"

namespace ttts {

constexpr int G2_BM = 128;          // rows per CTA (256 per pair)
constexpr int G2_BN = 256;          // columns per tile
constexpr int G2_BK = 64;
constexpr int G2_THREADS = 320;
constexpr int G2_A_BYTES = G2_BM * G2_BK * 2;            // 16 KB

// PAIR = true : CTA pair (cta_group::2), each CTA stages half of B (16 KB), 6 stages
// PAIR = false: single CTA (cta_group::1), full B tile (32 KB), 4 stages -- same warp-specialised pipelined epilogue
template <bool PAIR>
struct G2Cfg {
    static constexpr int kStages = PAIR ? 6 : 4;
    static constexpr int kBRows = PAIR ? G2_BN / 2 : G2_BN;
    static constexpr int kBBytes = kBRows * G2_BK * 2;
    static constexpr int kStageBytes = G2_A_BYTES + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kStageOff = kBarOffset + 1024;                          // barriers | per-warp epilogue staging units (1 KB aligned)
    static constexpr int kSmemBytes = kStageOff + 8 * EPI_WARP_BYTES + 1024;    // + alignment slack
    static constexpr int kTileM = PAIR ? 2 * G2_BM : G2_BM;
    static constexpr int kCtas = PAIR ? 2 : 1;
};

TTTS_DEVICE void decode_item2(const GemmParams& p, int item, int& m_pair, int& n_blk, int& split) {
    const int tiles = p.num_m_blocks * p.num_n_blocks;      // num_m_blocks counts 256-row pairs here
    split = item / tiles;
    const int t = item - split * tiles;
    const int group_size = p.group_m * p.num_n_blocks;
    const int g = t / group_size;
    const int r = t - g * group_size;
    const int m_first = g * p.group_m;
    const int gm = min(p.group_m, p.num_m_blocks - m_first);
    n_blk = r / gm;
    m_pair = m_first + (r - n_blk * gm);
}

template <bool A_MN, bool B_MN, bool PAIR, int EPI>
__global__ void __launch_bounds__(G2_THREADS, 1)   // 10 warps: 3 on one sub-partition -> 16384/3/32 = 168 registers per thread
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmAuxOut, const GemmParams p) {
    using C = G2Cfg<PAIR>;
    constexpr int G2_STAGES = C::kStages;
    constexpr int G2_STAGE_BYTES = C::kStageBytes;
    constexpr int G2_BAR_OFFSET = C::kBarOffset;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_BAR_OFFSET);
    uint64_t* empty_bar = full_bar + G2_STAGES;
    uint64_t* tfull_bar = empty_bar + G2_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* epi_ld_bar = tempty_bar + 2;                     // [8] one per epilogue warp: aux-input TMA loads
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(epi_ld_bar + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // cluster = 1 CTA, one CTA pair, or (p.quad) two CTA pairs on neighbouring n-blocks that share their A rows by TMA multicast:
    // the GEMM is bound by L2 -> SM operand delivery (profiles/r1_notes.md), and a 256 x 512 cluster tile moves 24 KB per CTA and k-block
    // instead of 32 KB.
    const bool quad = PAIR && p.quad;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;     // 0..1 (pair) or 0..3 (quad)
    const uint32_t rank = crank & 1u;                          // position inside the CTA pair
    const uint32_t pidx = quad ? (crank >> 1) : 0u;            // which pair of the cluster
    const bool leader = rank == 0;
    const int ncta = PAIR ? (quad ? 4 : 2) : 1;
    const int cluster_id = blockIdx.x / ncta;
    const int num_clusters = gridDim.x / ncta;
    const int nmul = quad ? 2 : 1;                             // n-blocks per item
    const int total_items = p.num_m_blocks * p.num_n_blocks * p.split_k;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        tma_prefetch_desc(&tmAux);
        tma_prefetch_desc(&tmAuxOut);
        for (int w = 0; w < 8; ++w) mbar_init(&epi_ld_bar[w], 1);
        // full[]: ONE arrival (the leader's arrive.expect_tx for both CTAs' bytes); the peer's TMA only complete_tx's there
        for (int s = 0; s < G2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], quad ? 2 : 1); }   // quad: both pairs' MMAs release a stage
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8 * C::kCtas); }     // one arrival per epilogue warp
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_2sm(tmem_holder, 512); else tmem_alloc(tmem_holder, 512); }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // PDL: everything above ran while the previous kernel of the stream was finishing; from here on its outputs are read
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ================= TMA producer (both CTAs): the whole warp runs the loop, one elected lane issues =================
        int stage = 0; uint32_t phase = 0;
        for (int item = cluster_id; item < total_items; item += num_clusters) {
            int m_pair, n_blk, split;
            decode_item2(p, item, m_pair, n_blk, split);
            const int m0 = m_pair * C::kTileM + (int)rank * G2_BM;
            const int n0 = (n_blk * nmul + (int)pidx) * G2_BN + (int)rank * (G2_BN / 2);
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                uint8_t* sb = sa + G2_A_BYTES;
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], C::kCtas * G2_STAGE_BYTES);
                    auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
                        if (PAIR) tma_load_2d_2sm(dst, tm, &full_bar[stage], c0, c1);
                        else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
                    };
                    if (quad) {
                        // this CTA fetches 64 of the 128 A rows its pair position needs and multicasts them to the CTA at the same position in
                        // the other pair; that CTA fetches the other 64 rows
                        const uint16_t mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));
                        if (A_MN) tma_load_2d_2sm_mc(sa + pidx * (G2_BK * 128), &tmA, &full_bar[stage], m0 + 64 * (int)pidx, kb * G2_BK, mask);
                        else tma_load_2d_2sm_mc(sa + pidx * (G2_BK * 128), &tmA, &full_bar[stage], kb * G2_BK, m0 + 64 * (int)pidx, mask);
                    } else if (A_MN) {
                        if (!PAIR && p.a3d) {
                            tma_load_3d(sa, &tmA, &full_bar[stage], 0, kb * G2_BK, m0 >> 6);
                        } else {
#pragma unroll
                            for (int j = 0; j < G2_BM / 64; ++j) load(sa + j * (G2_BK * 128), &tmA, m0 + 64 * j, kb * G2_BK);
                        }
                    } else {
                        load(sa, &tmA, kb * G2_BK, m0);
                    }
                    if (B_MN) {
                        if (!PAIR && p.b3d) {
                            tma_load_3d(sb, &tmB, &full_bar[stage], 0, kb * G2_BK, n0 >> 6);
                        } else {
#pragma unroll
                            for (int j = 0; j < C::kBRows / 64; ++j) load(sb + j * (G2_BK * 128), &tmB, n0 + 64 * j, kb * G2_BK);
                        }
                    } else {
                        load(sb, &tmB, kb * G2_BK, n0);
                    }
                }
                __syncwarp();
                if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ================= MMA issuer (leader CTA only): warp-uniform loop, elected lane issues =================
            constexpr uint32_t idesc = make_idesc_bf16(C::kTileM, G2_BN, A_MN, B_MN);
            constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4);      // descriptor start-address step per K=16 slice
            constexpr uint32_t kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
            const uint32_t smem_base = smem_u32(smem);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
                int m_pair, n_blk, split;
                decode_item2(p, item, m_pair, n_blk, split);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * G2_BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
                    const uint32_t sb = sa + G2_A_BYTES;
                    const uint64_t adesc0 = A_MN ? make_smem_desc_sw128(sa, G2_BK * 128, 1024) : make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(sb, G2_BK * 128, 1024) : make_smem_desc_sw128(sb, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < G2_BK / 16; ++k) {
                            if (PAIR) umma_bf16_2sm(tmem_d, adesc0 + k * kStepA, bdesc0 + k * kStepB, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                            else umma_bf16(tmem_d, adesc0 + k * kStepA, bdesc0 + k * kStepB, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        if (PAIR) umma_commit_2sm_mask(&empty_bar[stage], quad ? (uint16_t)0xF : (uint16_t)0x3); else umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) { if (PAIR) umma_commit_2sm_mask(&tfull_bar[as], (uint16_t)(0x3u << (2 * pidx))); else umma_commit(&tfull_bar[as]); }
                __syncwarp();
            }
        }
    } else if (warp >= 2) {
        // ================= epilogue (both CTAs, 8 warps) =================
        // Software-pipelined: this warp's 128 bias values (4 per lane) and the first chunk's residual / pre-activation TMA load are
        // issued BEFORE waiting for the accumulator (overlapping the main loop); inside the tile the TMEM load of chunk c+1 and the
        // aux load of chunk c+1 are in flight while chunk c is processed; TMEM is released right after the last load.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        EpiTmaCtx ec;
        ec.tm_out = &tmOut; ec.tm_aux = &tmAux; ec.tm_aux_out = &tmAuxOut;
        ec.U = smem_u32(smem + C::kStageOff + (warp - 2) * EPI_WARP_BYTES);
        ec.ldbar = &epi_ld_bar[warp - 2];
        ec.ld_phase = 0; ec.nstore = 0;
        int it = 0;
        for (int item = cluster_id; item < total_items; item += num_clusters, ++it) {
            int m_pair, n_blk, split;
            decode_item2(p, item, m_pair, n_blk, split);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int n0 = (n_blk * nmul + (int)pidx) * G2_BN;
            const int c0 = half * 4;
            const int row0w = m_pair * C::kTileM + (int)rank * G2_BM + q * 32;     // first row of this warp's 32-row slab
            const int colw = n0 + c0 * 32;                                         // first of this warp's 128 columns
            epi_tma_issue_load<EPI>(p, ec, row0w, colw, lane);
            if (epi_has_load<EPI>(p)) {
                if (it == 0) epi_l2_prefetch<EPI>(p, row0w, colw, lane);
                const int nitem = item + num_clusters;
                if (nitem < total_items) {
                    int m2, n2, s2;
                    decode_item2(p, nitem, m2, n2, s2);
                    epi_l2_prefetch<EPI>(p, m2 * C::kTileM + (int)rank * G2_BM + q * 32, (n2 * nmul + (int)pidx) * G2_BN + c0 * 32, lane);
                }
            }
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) {
                const int c = colw + lane * 4;
                if (c + 3 < p.N) bq = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                else {
                    if (c < p.N) bq.x = __ldg(p.bias + c);
                    if (c + 1 < p.N) bq.y = __ldg(p.bias + c + 1);
                    if (c + 2 < p.N) bq.z = __ldg(p.bias + c + 2);
                }
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * G2_BN;
            uint32_t rA[32], rB[32];
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            __syncwarp();
            tmem_ld_32x32(taddr + c0 * 32, rA);
            // chunk 0 (registers A) while chunk 1 loads into B, and so on: explicit ping-pong, no register copies
            tmem_ld_wait();
            __syncwarp();
            tmem_ld_32x32(taddr + (c0 + 1) * 32, rB);
            epi_tma_apply<EPI>(p, ec, row0w, colw, lane, rA, bq, 0, colw + 32);
            tmem_ld_wait();
            __syncwarp();
            tmem_ld_32x32(taddr + (c0 + 2) * 32, rA);
            epi_tma_apply<EPI>(p, ec, row0w, colw + 32, lane, rB, bq, 1, colw + 64);
            tmem_ld_wait();
            __syncwarp();
            tmem_ld_32x32(taddr + (c0 + 3) * 32, rB);
            epi_tma_apply<EPI>(p, ec, row0w, colw + 64, lane, rA, bq, 2, colw + 96);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_rank(&tempty_bar[as], crank & ~1u); else mbar_arrive(&tempty_bar[as]); }   // accumulator stage drained
            epi_tma_apply<EPI>(p, ec, row0w, colw + 96, lane, rB, bq, 3, -1);
        }
        if (lane == 0) bulk_wait_all<0>();       // this warp's TMA stores have left shared memory and are globally performed
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer may still be reading this CTA's smem / arriving on its barriers until here
    tc_fence_after();
    if (warp == 1) { __syncwarp(); if (PAIR) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

template <bool A_MN, bool B_MN, bool PAIR, int EPI = -1>
static int launch_gemm2(const ttts_gemm_args& a, const GemmParams& p_in, int grid, cudaStream_t stream) {
    using C = G2Cfg<PAIR>;
    GemmParams p = p_in;
    CUtensorMap tmA, tmB;
    int rc;
    static int use3d = -1;
    if (use3d < 0) { const char* e = getenv("TTTS_GEMM_NO3D"); use3d = (e && e[0] == '1') ? 0 : 1; }
    p.a3d = (A_MN && !PAIR && use3d && a.M % 64 == 0) ? 1 : 0;
    p.b3d = (B_MN && !PAIR && use3d && a.N % 64 == 0) ? 1 : 0;
    if (A_MN) rc = p.a3d ? make_tmap_mn3d(&tmA, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, G2_BK, G2_BM / 64)
                         : make_tmap_2d(&tmA, a.A, 2, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, 64, G2_BK, true);
    else      rc = make_tmap_2d(&tmA, a.A, 2, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, G2_BK, p.quad ? G2_BM / 2 : G2_BM, true);   // quad: 64-row halves
    if (rc) return rc;
    if (B_MN) rc = p.b3d ? make_tmap_mn3d(&tmB, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, G2_BK, C::kBRows / 64)
                         : make_tmap_2d(&tmB, a.B, 2, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, 64, G2_BK, true);
    else      rc = make_tmap_2d(&tmB, a.B, 2, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, G2_BK, C::kBRows, true);
    if (rc) return rc;
    // epilogue tensor maps: [32 rows x 64 B] boxes in the 64-byte swizzle (32 bf16 or 16 fp32 columns), true extents so ragged edges clip
    CUtensorMap tmOut, tmAux, tmAuxOut;
    const bool out_f32 = (a.epi == TTTS_EPI_RESID || a.epi == TTTS_EPI_F32_ADD || a.epi == TTTS_EPI_F32);
    rc = make_tmap_2d(&tmOut, a.out, out_f32 ? 4 : 2, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldo, out_f32 ? 16 : 32, 32, 2);
    if (rc) return rc;
    tmAux = tmOut; tmAuxOut = tmOut;
    if (a.epi == TTTS_EPI_RESID) rc = make_tmap_2d(&tmAux, a.aux, 4, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldaux, 16, 32, 2);
    else if (a.epi == TTTS_EPI_DGELU) rc = make_tmap_2d(&tmAux, a.aux, 2, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldaux, 32, 32, 2);
    if (rc) return rc;
    if (a.epi == TTTS_EPI_GELU && a.aux_out) rc = make_tmap_2d(&tmAuxOut, a.aux_out, 2, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldaux_out, 32, 32, 2);
    if (rc) return rc;
    auto kern = gemm2_bf16_kernel<A_MN, B_MN, PAIR, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        TTTS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(G2_THREADS); cfg.dynamicSmemBytes = C::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? (p.quad ? 4 : 2) : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    prof_gemm_begin(stream, 2.0 * (double)a.M * (double)a.N * (double)a.K);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmAux, tmAuxOut, p);
    prof_gemm_end(stream);
    if (e != cudaSuccess) return fail_cuda(e, "gemm2_bf16_kernel launch");
    TTTS_LAUNCH_CHECK("gemm2_bf16_kernel");
    return TTTS_OK;
}

// TTTS_GEMM_QUAD=1: clusters of 4 CTAs (two pairs) with the A tile multicast to both pairs
static bool use_quad() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_GEMM_QUAD"); on = (e && e[0] == '1') ? 1 : 0; }
    return on != 0;
}
// how many 4-CTA clusters of the pair kernel the device can hold at once (0: do not use quad mode)
static int quad_clusters() {
    static int n = -1;
    if (n < 0) {
        using C = G2Cfg<true>;
        auto kern = gemm2_bf16_kernel<false, true, true, -1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(4 * (num_sms() / 4)); cfg.blockDim = dim3(G2_THREADS); cfg.dynamicSmemBytes = C::kSmemBytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int c = 0;
        if (cudaOccupancyMaxActiveClusters(&c, kern, &cfg) != cudaSuccess) { cudaGetLastError(); c = 0; }
        n = (c * 4 >= num_sms() * 85 / 100) ? c : 0;         // B200: 33 clusters of 4 = 132 of 148 SMs (GPC granularity)
        if (getenv("TTTS_GEMM_QUAD_VERBOSE")) fprintf(stderr, "[ttts] quad clusters: occupancy query says %d -> using %d\n", c, n);
    }
    return n;
}

template <bool PAIR>
static int gemm2_impl(const ttts_gemm_args& a, cudaStream_t stream) {
    using C = G2Cfg<PAIR>;
    GemmParams p;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.num_m_blocks = (a.M + C::kTileM - 1) / C::kTileM;
    p.num_n_blocks = (a.N + G2_BN - 1) / G2_BN;
    const int clusters = num_sms() / C::kCtas;
    p.group_m = clusters / p.num_n_blocks;
    if (p.group_m < 1) p.group_m = 1;
    if (p.group_m > p.num_m_blocks) p.group_m = p.num_m_blocks;
    p.num_k_blocks = (a.K + G2_BK - 1) / G2_BK;
    int split = a.split_k < 1 ? 1 : a.split_k;
    if (split > p.num_k_blocks) split = p.num_k_blocks;
    p.kb_per_split = (p.num_k_blocks + split - 1) / split;
    p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    p.epi = a.epi;
    p.out = a.out; p.ldo = a.ldo; p.bias = a.bias;
    p.aux = a.aux; p.ldaux = a.ldaux; p.aux_out = a.aux_out; p.ldaux_out = a.ldaux_out;
    p.drop_thresh16 = a.drop_thresh16; p.drop_scale = a.drop_scale; p.drop_seed = a.drop_seed;
    p.a3d = p.b3d = 0;
    p.quad = 0;
    static int l2pf = -1;
    if (l2pf < 0) { const char* e = getenv("TTTS_GEMM_L2PF"); l2pf = (e && e[0] == '1') ? 1 : 0; }
    p.l2pf = l2pf;
    int items = p.num_m_blocks * p.num_n_blocks * p.split_k;
    int grid = C::kCtas * (items < clusters ? items : clusters);
    if (PAIR && use_quad() && p.num_n_blocks >= 2) {
        // clusters of two pairs: items cover two neighbouring n-blocks (an odd last block leaves the second pair idle on an out-of-range block)
        const int q = quad_clusters();
        if (q > 0) {
            p.quad = 1;
            p.num_n_blocks = (p.num_n_blocks + 1) / 2;
            p.group_m = q / p.num_n_blocks;
            if (p.group_m < 1) p.group_m = 1;
            if (p.group_m > p.num_m_blocks) p.group_m = p.num_m_blocks;
            items = p.num_m_blocks * p.num_n_blocks * p.split_k;
            grid = 4 * (items < q ? items : q);
        }
    }
    // the step's hot (operand majors, epilogue) combinations run kernels with the epilogue fixed at compile time (TTTS_GEMM_DYN_EPI=1: off)
    static int dyn = -1;
    if (dyn < 0) { const char* e = getenv("TTTS_GEMM_DYN_EPI"); dyn = (e && e[0] == '1') ? 1 : 0; }
    if constexpr (PAIR) if (!dyn && !p.quad) {
        if (!a.a_mn && a.b_mn) {
            if (a.epi == TTTS_EPI_BF16) return launch_gemm2<false, true, PAIR, TTTS_EPI_BF16>(a, p, grid, stream);
            if (a.epi == TTTS_EPI_RESID) return launch_gemm2<false, true, PAIR, TTTS_EPI_RESID>(a, p, grid, stream);
            if (a.epi == TTTS_EPI_GELU) return launch_gemm2<false, true, PAIR, TTTS_EPI_GELU>(a, p, grid, stream);
        } else if (!a.a_mn && !a.b_mn) {
            if (a.epi == TTTS_EPI_BF16) return launch_gemm2<false, false, PAIR, TTTS_EPI_BF16>(a, p, grid, stream);
            if (a.epi == TTTS_EPI_DGELU) return launch_gemm2<false, false, PAIR, TTTS_EPI_DGELU>(a, p, grid, stream);
        } else if (a.a_mn && a.b_mn) {
            if (a.epi == TTTS_EPI_F32_ADD) return launch_gemm2<true, true, PAIR, TTTS_EPI_F32_ADD>(a, p, grid, stream);
        }
    }
    if (!a.a_mn && !a.b_mn) return launch_gemm2<false, false, PAIR>(a, p, grid, stream);
    if (!a.a_mn && a.b_mn) return launch_gemm2<false, true, PAIR>(a, p, grid, stream);
    if (a.a_mn && a.b_mn) return launch_gemm2<true, true, PAIR>(a, p, grid, stream);
    return launch_gemm2<true, false, PAIR>(a, p, grid, stream);
}

// caller (gemm_bf16) has validated the arguments
int gemm2_bf16(const ttts_gemm_args& a, bool pair, cudaStream_t stream) {
    return pair ? gemm2_impl<true>(a, stream) : gemm2_impl<false>(a, stream);
}

int pick_split_k2(int M, int N, int K, bool pair) {
    const int tm = pair ? 2 * G2_BM : G2_BM;
    const int tiles = ((M + tm - 1) / tm) * ((N + G2_BN - 1) / G2_BN);
    const int kblocks = (K + G2_BK - 1) / G2_BK;
    const int clusters = num_sms() / (pair ? 2 : 1);
    int best = 1; double best_eff = -1.0;
    for (int s = 1; s <= 32 && s <= kblocks; ++s) {
        if (kblocks / s < 8 && s > 1) break;
        const long items = (long)tiles * s;
        const long waves = (items + clusters - 1) / clusters;
        const double eff = (double)items / (double)(waves * clusters);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

}  // namespace ttts
