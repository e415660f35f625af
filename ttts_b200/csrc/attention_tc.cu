// Causal flash attention on the 5th-gen tensor cores (tcgen05.mma + TMEM + TMA), head_dim 64, forward and backward.
// Replaces HF GPT2Attention (HF: modeling_gpt2.py:144-226) like attention.cu, which stays as the legacy mma.sync path
// (TTTS_ATTN_LEGACY=1) and as a cross-check in the tests.
//
// Every operand tile is [128 rows x 64 bf16] = 128 rows of 128 B in the 128B-swizzle layout, loaded by ONE TMA box straight
// out of the packed c_attn output [B*T, 3d] (or the [B*T, d] dO buffer).  The same bytes serve as a K-major operand (rows =
// M/N index, the 64 head dims = K) and as an MN-major operand (rows = K index, the 64 head dims = M/N), so no transposes:
//
//   forward   S  = Q K^T        A=Q  (K-maj)  B=K (K-maj)   128x128x64  -> TMEM
//             P  = softmax tile (4 warps, one query row per thread = one TMEM lane: no shuffles), bf16 -> smem
//             O += P V          A=P  (K-maj)  B=V (MN-maj)  128x64x128  -> TMEM scratch, folded into registers
//   backward (CTA owns a key block, loops over query blocks):
//             S  = Q K^T , dP = dO V^T                       128x128x64 each -> TMEM
//             P = exp(S*scale - lse), dS = P (dP - delta) scale   (thread = query row) -> bf16 smem, both
//             dV += P^T dO      A=P  (MN-maj) B=dO (MN-maj) 128x64x128
//             dK += dS^T Q      A=dS (MN-maj) B=Q  (MN-maj) 128x64x128
//             dQ  = dS K        A=dS (K-maj)  B=K  (MN-maj) 128x64x128 -> TMEM -> red.global.add.f32 into an fp32 dQ buffer
//
// Warp roles (576 threads): warp 0 = TMA loader, warp 1 = MMA issuer + TMEM alloc, warps 2-17 = softmax / gradient math: four warps
// per TMEM lane quadrant, each owning a quarter of the tile's columns (32 keys, 16 of the 64 output dims).  Earlier versions (one,
// then two math warps per sub-partition, one CTA per (block, head) item) are described in profiles/r1_notes.md; what is kept here is
// the persistent kernel pair in two flavours: kMode 0 = "version 4" as profiled in r1h, kMode 1 / 2 = "version 5" (default).
#include <stdlib.h>
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int AT_BM = 128;            // queries per tile
constexpr int AT_BN = 128;            // keys per tile
constexpr int AT_TILE = 128 * 128;    // bytes of a [128 x 64] bf16 tile
constexpr float kLog2eF = 1.4426950408889634f;

TTTS_DEVICE void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
TTTS_DEVICE float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// K-major descriptor for k-step kk (16 elements of K) inside a [rows x 64] tile (or the 2-atom [rows x 128] P/dS buffers)
TTTS_DEVICE uint64_t desc_kmajor(uint32_t base, int kk) { return make_smem_desc_sw128(base + (kk >> 2) * AT_TILE + (kk & 3) * 32, 16, 1024); }
// MN-major descriptor for k-step kk (16 rows of K); MN extent 64 (single atom) or 128 (two atoms, AT_TILE apart)
TTTS_DEVICE uint64_t desc_mnmajor(uint32_t base, int kk) { return make_smem_desc_sw128(base + kk * 2048, AT_TILE, 1024); }

// write 8 bf16 (16 B) of row r, 16B-chunk c16 (0..15 over 128 columns) into a 2-atom K-major/MN-major swizzled buffer
TTTS_DEVICE void st_tile_chunk(uint32_t base, int r, int c16, uint4 v) {
    const uint32_t addr = base + (c16 >> 3) * AT_TILE + r * 128 + ((((c16 & 7) ^ r) & 7) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

TTTS_DEVICE uint32_t tile_chunk_addr(uint32_t base, int r, int c16) { return base + (c16 >> 3) * AT_TILE + r * 128 + ((((c16 & 7) ^ r) & 7) << 4); }
TTTS_DEVICE void sts_v4(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
// shared-space scalar accesses by 32-bit address: through a generic pointer derived from the aligned dynamic-smem base the compiler emits
// generic LD / ST (r1n profile: the row-max exchange's LD.E was the third-hottest stall site of the forward kernel)
TTTS_DEVICE void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
TTTS_DEVICE float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }

// attention_tail.cu
int attn_tail_rows(int T);
int attn_tail_fwd(const bf16* qkv, bf16* out, float* lse, int B, int T, int H, int Tm, uint32_t thresh16, float drop_scale, uint64_t seed, cudaStream_t st);
int attn_tail_bwd(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, bf16* dqkv, float* dq_acc, int B, int T, int H, int Tm,
                  uint32_t thresh16, float drop_scale, uint64_t seed, cudaStream_t st);
bool attn_tail_split();

constexpr int AT3_THREADS = 576;      // warp 0 TMA, warp 1 MMA, 16 math warps (four per TMEM lane quadrant, 32 columns each)

// ------------------------------------------------------------------------------------------------------------
// Version 4 = version 3 made PERSISTENT: one CTA per SM walks a list of (query block, head) items, heavy items first.  TMEM is
// allocated once, the barrier phases run on across items, the TMA warp prefetches the next item's Q / K / V while the current item is
// still being processed, and the MMA warp issues S of the next item's first key block before the last P V of the current one: the
// per-CTA prologue / epilogue that cost ~1 000 (forward) - 2 600 (backward) cycles per 128x128 block pair at T = 1 156 (5.5 block pairs
// per CTA on average, profiles/r1_notes.md) is overlapped instead of exposed.
// ------------------------------------------------------------------------------------------------------------
// Exact n / d for n * d < 2^32 by one multiply-high (m = ceil(2^32 / d); d == 1 handled apart).  The persistent kernels decompose an item
// index three times per item in each of the 18 warps; the generic integer division (I2F + MUFU.RCP + ~20 instructions) queued behind the
// softmax warps' MUFU.EX2 stream (profiles/r1h_attn_bwd4_ncu_full.txt: 3.7 % of all stall samples on one I2F).
struct FastDiv { uint32_t d, m; };
TTTS_DEVICE FastDiv fastdiv_make(uint32_t d) { FastDiv f; f.d = d; f.m = (uint32_t)((0x100000000ULL + d - 1) / d); return f; }
TTTS_DEVICE uint32_t fastdiv(uint32_t n, const FastDiv f) { return f.d == 1 ? n : __umulhi(n, f.m); }
TTTS_DEVICE uint32_t fastmod(uint32_t n, const FastDiv f) { return n - fastdiv(n, f) * f.d; }

struct Fwd4Smem {
    static constexpr int kKvStages = 3;
    static constexpr int oQ = 0;                                        // [2]
    static constexpr int oKV = 2 * AT_TILE;                             // [stages][K | V]
    static constexpr int oP = oKV + kKvStages * 2 * AT_TILE;            // [2][2 atoms]
    static constexpr int oBar = oP + 2 * 2 * AT_TILE;
    static constexpr int oXch = oBar + 256;                             // row max [2][4][128] + row sum [4][128] floats
    static constexpr int kBytes = oXch + (2 * 4 * 128 + 4 * 128) * 4 + 1024;
};

// kMode 0: the r1e kernel as measured in profiles/r1h_attn_fwd4_ncu_full.txt (dropout decided at run time, after all exponentials, one
// 512-thread barrier per key block).  kMode 1 (dropout) / 2 (none) = version 5: the r1h profile showed the 16 softmax warps spending 27 % of
// their time in a MUFU-bound exponential phase (32 MUFU.EX2 per warp, 4 warps per sub-partition in lock step) followed by 26 % in a pure
// integer phase (the dropout hash, behind a uniform branch the scheduler cannot move code across).  With the dropout decision a template
// parameter and the hash of each 4-key group written next to that group's exponentials, both sit in one basic block and the IMAD / LOP3
// work issues in the shadow of the MUFU pipe.  The row-max exchange only concerns the four warps that share a lane quadrant (same rows,
// different columns), so it uses one 128-thread named barrier per quadrant instead of one for all 16 warps.  Same arithmetic, same bits.
template <int kMode>
__global__ void __maxnreg__(96)
attn_fwd_tc4_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out, int T, int Tm, int H, int BH,
                    float scale, DropCfg drop) {
    // T = positions per sequence (row / lse / dropout-row strides); Tm <= T = the positions this kernel attends over (queries AND keys):
    // Tm = T, or T - T mod 128 when the ragged tail rows are left to attention_tail.cu
    using S = Fwd4Smem;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::oBar);
    uint64_t* q_full = bars;                       // [2]
    uint64_t* q_empty = bars + 2;                  // [2] commit: every S MMA of the item has read Q
    uint64_t* kv_full = bars + 4;                  // [3]
    uint64_t* kv_empty = bars + 7;                 // [3]
    uint64_t* s_full = bars + 10;                  // [2]
    uint64_t* s_empty = bars + 12;                 // [2]  16
    uint64_t* p_full = bars + 14;                  // [2]  16
    uint64_t* p_empty = bars + 16;                 // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = H * 64;
    const int nq = (Tm + AT_BM - 1) / AT_BM;
    const int total = nq * BH;
    // Items are numbered head-major (all query blocks of a head are neighbours, late = heavy blocks first) and dealt out in rounds of
    // gridDim.x: in round n this CTA takes position (blockIdx.x + n) mod gridDim.x.  CTAs running at the same time therefore work on
    // ~gridDim.x / nq neighbouring heads (their K/V stay in L2), and the rotation walks every CTA through all block weights.
    const int rounds = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const FastDiv fd_grid = fastdiv_make(gridDim.x), fd_nq = fastdiv_make((uint32_t)nq), fd_H = fastdiv_make((uint32_t)H);
    auto item_at = [&](int n) { if (n >= rounds) return -1; const int idx = n * (int)gridDim.x + (int)fastmod(blockIdx.x + n, fd_grid); return idx < total ? idx : -1; };
    auto item_qb = [&](int idx) { return nq - 1 - (int)fastmod((uint32_t)idx, fd_nq); };
    auto item_bh = [&](int idx) { return (int)fastdiv((uint32_t)idx, fd_nq); };
    auto item_nkv = [&](int qb) { return min(qb + 1, (Tm + AT_BN - 1) / AT_BN); };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
        for (int s = 0; s < S::kKvStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 16);
            mbar_init(&p_full[s], 16); mbar_init(&p_empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tS = tmem_base;              // [2] x 128 cols
    const uint32_t tO = tmem_base + 256;        // 64 cols
    pdl_launch_dependents();                    // PDL (common.cuh): the prologue above overlapped the previous kernel's tail
    pdl_wait();

    if (warp == 0) {
        // ---------------- TMA: Q per item (double-buffered), K/V ring across items ----------------
        uint32_t kvc = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int qb = item_qb(idx), bh = item_bh(idx), b = (int)fastdiv((uint32_t)bh, fd_H), h = bh - b * H;
            const int row_base = b * T, nkv = item_nkv(qb);
            mbar_wait_relaxed(&q_empty[n & 1], ((n >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&q_full[n & 1], AT_TILE);
                tma_load_2d(smem + S::oQ + (n & 1) * AT_TILE, &tmQKV, &q_full[n & 1], h * 64, row_base + qb * AT_BM);
            }
            __syncwarp();
            for (int j = 0; j < nkv; ++j, ++kvc) {
                const int st = kvc % S::kKvStages; const uint32_t ph = (kvc / S::kKvStages) & 1;
                mbar_wait_relaxed(&kv_empty[st], ph ^ 1);
                uint8_t* sk = smem + S::oKV + st * 2 * AT_TILE;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE);
                    tma_load_2d(sk, &tmQKV, &kv_full[st], d + h * 64, row_base + j * AT_BN);
                    tma_load_2d(sk + AT_TILE, &tmQKV, &kv_full[st], 2 * d + h * 64, row_base + j * AT_BN);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA: S runs exactly one key block ahead of P V, across item boundaries ----------------
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
        const uint32_t smem_base = smem_u32(smem);
        // S cursor
        int nS = 0, jS = 0, idxS = item_at(0), nkvS = idxS >= 0 ? item_nkv(item_qb(idxS)) : 0;
        uint32_t gS = 0;
        auto issue_next_s = [&]() {
            if (idxS < 0) return;
            if (jS == 0) mbar_wait(&q_full[nS & 1], (nS >> 1) & 1);
            const int st = gS % S::kKvStages;
            mbar_wait(&kv_full[st], (gS / S::kKvStages) & 1);
            mbar_wait(&s_empty[gS & 1], ((gS >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint64_t dQ = desc_kmajor(smem_base + S::oQ + (nS & 1) * AT_TILE, 0);
            const uint64_t dK = desc_kmajor(smem_base + S::oKV + st * 2 * AT_TILE, 0);
            const bool last = (jS == nkvS - 1);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tS + (gS & 1) * 128, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[gS & 1]);
                if (last) umma_commit(&q_empty[nS & 1]);
            }
            __syncwarp();
            ++gS; ++jS;
            if (last) { ++nS; idxS = item_at(nS); jS = 0; nkvS = idxS >= 0 ? item_nkv(item_qb(idxS)) : 0; }
        };
        issue_next_s();
        uint32_t g = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int nkv = item_nkv(item_qb(idx));
            for (int j = 0; j < nkv; ++j, ++g) {
                issue_next_s();
                const int st = g % S::kKvStages;
                mbar_wait(&p_full[g & 1], (g >> 1) & 1);       // P in smem, any rescale of O done, (j == 0) previous item's O read out
                tc_fence_after();
                const uint64_t dP = desc_kmajor(smem_base + S::oP + (g & 1) * 2 * AT_TILE, 0);
                const uint64_t dV = desc_mnmajor(smem_base + S::oKV + st * 2 * AT_TILE + AT_TILE, 0);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16(tO, dP + (uint64_t)((k >> 2) * (AT_TILE >> 4) + (k & 3) * 2), dV + (uint64_t)(k * 128), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                    // one commit serves both "the P buffer is free" and "O holds key block g": the softmax warps wait on p_empty[g & 1] before
                    // they rescale or read O (a separate barrier for the second meaning completed a phase per key block that nobody waited
                    // for in most blocks -- legal, but compute-sanitizer synccheck calls it a missing wait; profiles/r2c_sanitizer_synccheck.log)
                    umma_commit(&p_empty[g & 1]);
                    umma_commit(&kv_empty[st]);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int qtr = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * kLog2eF;
        const uint32_t xmax = smem_u32(smem + S::oXch);                       // [buf][qtr][row] floats, shared-space address
        const uint32_t xsum = xmax + 2 * 4 * 128 * 4;                         // [qtr][row]
        const uint32_t addc = attn_drop_addc(drop.thresh16);
        uint32_t p_addr[4];                                                   // this thread's four 16-byte chunks of a P tile (buffer 0)
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) p_addr[gg] = tile_chunk_addr(smem_u32(smem + S::oP), r, qtr * 4 + gg);
        uint32_t g = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int qb = item_qb(idx), bh = item_bh(idx), b = (int)fastdiv((uint32_t)bh, fd_H), h = bh - b * H;
            const int row_base = b * T, nkv = item_nkv(qb);
            const int qi = qb * AT_BM + r;
            float m_used = -INFINITY, l_run = 0.f;
            const AttnDropRow rk = attn_drop_row(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi);
            for (int j = 0; j < nkv; ++j, ++g) {
                const int k0 = j * AT_BN;
                const bool need_mask = (j == qb) || (k0 + AT_BN > Tm);
                const int kc0 = k0 + qtr * 32;
                uint32_t v[32];
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                __syncwarp();
                tmem_ld_32x32(tS + (g & 1) * 128 + lane_off + qtr * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[g & 1]);
                float mx0 = -INFINITY, mx1 = -INFINITY;
                if (need_mask) {
                    const int nv = min(qi, Tm - 1) - kc0 + 1;        // keys kc0 .. kc0 + nv - 1 exist for this query row (causal + sequence end)
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = (i < nv) ? v[i] : 0xff800000u;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) { mx0 = fmaxf(mx0, __uint_as_float(v[i])); mx1 = fmaxf(mx1, __uint_as_float(v[i + 1])); }
                sts_f32(xmax + (((g & 1) * 4 + qtr) * 128 + r) * 4, fmaxf(mx0, mx1));
                named_bar_sync(2 + quad, 128);
                const uint32_t xm = xmax + ((g & 1) * 4 * 128 + r) * 4;
                const float m_new = fmaxf(fmaxf(fmaxf(lds_f32(xm), lds_f32(xm + 512)), fmaxf(lds_f32(xm + 1024), lds_f32(xm + 1536))), m_used);
                if (j == 0) {
                    m_used = m_new;
                } else {
                    const bool grow = (m_new - m_used) * sl2 > 8.f;
                    if (__any_sync(0xffffffffu, grow)) {
                        const float f = grow ? ex2_fast((m_used - m_new) * sl2) : 1.f;
                        mbar_wait(&p_empty[(g - 1) & 1], ((g - 1) >> 1) & 1);      // P V of key block g - 1 has completed (see the MMA warp)
                        tc_fence_after();
                        uint32_t o[16];
                        __syncwarp();
                        tmem_ld_32x16(tO + lane_off + qtr * 16, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                        tmem_st_32x16(tO + lane_off + qtr * 16, o);
                        tmem_st_wait();
                        l_run *= f;
                        if (grow) m_used = m_new;
                    }
                }
                const float msc = m_used * sl2;
                float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
                uint32_t pk[16];
                // ONE basic block per key block: exponentials (MUFU + FMA pipe), row sums (FMA pipe), the dropout hash (IMAD.WIDE on the FMA
                // pipe, LOP3 on the ALU pipe) and the keep masks, applied to the PACKED bf16x2 probabilities (add + PRMT + AND per two keys;
                // common.cuh attn_drop_mask2).  r1n profile of the previous form (fp32 selects: shift + ISETP + SEL per key, a dead VIADD per
                // multiply, the P-tile addresses recomputed per block): 17.9 warp instructions per probability, ALU pipe (half rate) busiest.
                // r2c profile of the single-basic-block form: the instruction count halved (17.9 -> 8.9 per probability) but the kernel only
                // gained 10 %: ptxas emits the 32 exponentials of a block back to back, the four warps of a scheduler reach that phase together
                // (row-max barrier), so the MUFU pipe (one warp instruction per 8 cycles and scheduler) runs for ~1 000 cycles with the FMA / ALU
                // pipes idle, then the integer phase runs with the MUFU pipe idle.  The exponent offset of 4-key group i is therefore made to
                // DEPEND on the packed probabilities of group i - 2 (OR-ing in a word that is zero at run time, thresh16 < 2^16, but not provably
                // so): two interleaved dependency chains, so at most 8 exponentials can be adjacent and the hash / row-sum / pack / mask work of
                // the neighbouring groups has to be scheduled between them -- a software pipeline ptxas cannot undo.
                {
                    const uint32_t g0 = (uint32_t)kc0 >> 2;
                    const uint32_t zero = drop.thresh16 >> 16;
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        float mo = msc;
                        if (i4 >= 2) mo = __uint_as_float(__float_as_uint(msc) | (pk[2 * (i4 - 2)] & zero));
                        const float p0 = ex2_fast(fmaf(__uint_as_float(v[4 * i4]), sl2, -mo));
                        const float p1 = ex2_fast(fmaf(__uint_as_float(v[4 * i4 + 1]), sl2, -mo));
                        const float p2 = ex2_fast(fmaf(__uint_as_float(v[4 * i4 + 2]), sl2, -mo));
                        const float p3 = ex2_fast(fmaf(__uint_as_float(v[4 * i4 + 3]), sl2, -mo));
                        rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
                        if (kMode == 1) {
                            uint32_t w0, w1;
                            attn_drop_words(rk, g0 + i4, w0, w1);
                            pk[2 * i4] = pack_bf16(p0, p1) & attn_drop_mask2(w0, addc);
                            pk[2 * i4 + 1] = pack_bf16(p2, p3) & attn_drop_mask2(w1, addc);
                        } else {
                            pk[2 * i4] = pack_bf16(p0, p1);
                            pk[2 * i4 + 1] = pack_bf16(p2, p3);
                        }
                    }
                }
                l_run += (rs0 + rs1) + (rs2 + rs3);
#pragma unroll
                for (int e = 0; e < 16; ++e) asm volatile("" : "+r"(pk[e]));       // packed BEFORE the wait: the chain must not sink below the wait loop
                mbar_wait(&p_empty[g & 1], ((g >> 1) & 1) ^ 1);
                {
                    const uint32_t pb = (g & 1) * 2 * AT_TILE;
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) sts_v4(p_addr[gg] + pb, make_uint4(pk[4 * gg], pk[4 * gg + 1], pk[4 * gg + 2], pk[4 * gg + 3]));
                }
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g & 1]);
            }
            // item epilogue: O / l (dropout's 1/(1-p) folded in), lse.  The next item's first P V (which overwrites O) cannot be issued
            // before every softmax warp has passed this point and arrived on that block's p_full.
            mbar_wait(&p_empty[(g - 1) & 1], ((g - 1) >> 1) & 1);                  // the item's last P V has completed
            tc_fence_after();
            uint32_t o[16];
            __syncwarp();
            tmem_ld_32x16(tO + lane_off + qtr * 16, o);
            tmem_ld_wait();
            tc_fence_before();
            sts_f32(xsum + (qtr * 128 + r) * 4, l_run);
            named_bar_sync(2 + quad, 128);
            const float l_tot = (lds_f32(xsum + r * 4) + lds_f32(xsum + (128 + r) * 4)) + (lds_f32(xsum + (256 + r) * 4) + lds_f32(xsum + (384 + r) * 4));
            if (qi < Tm) {
                const float inv = l_tot > 0.f ? drop.scale / l_tot : 0.f;
                if (qtr == 0) lse_out[(size_t)bh * T + qi] = m_used * scale + logf(l_tot);
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(row_base + qi) * d + h * 64 + qtr * 16);
#pragma unroll
                for (int gg = 0; gg < 2; ++gg)
                    dst[gg] = make_uint4(pack_bf16(__uint_as_float(o[8 * gg]) * inv, __uint_as_float(o[8 * gg + 1]) * inv),
                                         pack_bf16(__uint_as_float(o[8 * gg + 2]) * inv, __uint_as_float(o[8 * gg + 3]) * inv),
                                         pack_bf16(__uint_as_float(o[8 * gg + 4]) * inv, __uint_as_float(o[8 * gg + 5]) * inv),
                                         pack_bf16(__uint_as_float(o[8 * gg + 6]) * inv, __uint_as_float(o[8 * gg + 7]) * inv));
            }
            // xsum is rewritten only after the next item's row-max barriers: no extra sync needed (every item has >= 1 key block)
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// Softmax-gradient math of one thread's 32 keys of a [128 x 128] block: P (dropout applied) and dS / c as packed bf16x2 words.
//     pe = exp2(S * scale * log2e - lse2) ; nd = pe * (-delta / c) ; kept key: dS / c = fma(pe, dP, nd) ; dropped key: nd
// The keep decision is taken on the PACKED words (common.cuh attn_drop_mask2).  kMasked: diagonal / edge blocks, keys >= n_ok do not exist
// (a separate instantiation: folded into one body the compiler predicates the per-key compare + select into EVERY block, r2c profile).
// Software pipeline as in the forward kernel: the exponent offset of group i depends on the packed P of group i - 2 through a word that
// is zero at run time, so the exponentials cannot be batched and the integer / FMA work of the neighbouring groups fills the MUFU shadow.
template <int kMode, bool kMasked>
TTTS_DEVICE void bwd_math(const int cc, const uint32_t (&sv)[16], const uint32_t (&gv)[16], uint32_t (&pp)[16], uint32_t (&dd)[16], float sl2, float lse2,
                          float ndl, const AttnDropRow rk, uint32_t g0, uint32_t addc, int n_ok, uint32_t zero) {
    // cc = 0 / 1: the thread's first / second 16 keys (groups 0-3 / 4-7 of the block; the dependency chain runs on across the two calls)
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        const int i4 = cc * 4 + j4;
        float lo = lse2;
        if (i4 >= 2) lo = __uint_as_float(__float_as_uint(lse2) | (pp[2 * (i4 - 2)] & zero));
        float pe[4], nd[4], dk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            pe[e] = ex2_fast(fmaf(__uint_as_float(sv[j4 * 4 + e]), sl2, -lo));
            if (kMasked) pe[e] = (i4 * 4 + e < n_ok) ? pe[e] : 0.f;
            nd[e] = pe[e] * ndl;
            dk[e] = fmaf(pe[e], __uint_as_float(gv[j4 * 4 + e]), nd[e]);
        }
        if (kMode == 1) {
            uint32_t w0, w1;
            attn_drop_words(rk, g0 + i4, w0, w1);
            const uint32_t m0 = attn_drop_mask2(w0, addc), m1 = attn_drop_mask2(w1, addc);
            pp[2 * i4] = pack_bf16(pe[0], pe[1]) & m0;
            pp[2 * i4 + 1] = pack_bf16(pe[2], pe[3]) & m1;
            dd[2 * i4] = (pack_bf16(dk[0], dk[1]) & m0) | (pack_bf16(nd[0], nd[1]) & ~m0);
            dd[2 * i4 + 1] = (pack_bf16(dk[2], dk[3]) & m1) | (pack_bf16(nd[2], nd[3]) & ~m1);
        } else {
            pp[2 * i4] = pack_bf16(pe[0], pe[1]);
            pp[2 * i4 + 1] = pack_bf16(pe[2], pe[3]);
            dd[2 * i4] = pack_bf16(dk[0], dk[1]);
            dd[2 * i4 + 1] = pack_bf16(dk[2], dk[3]);
        }
    }
}

// backward, persistent (see the forward above).  Items = (key block, head), heavy (early) key blocks first; K/V double-buffered so the
// next item's tiles arrive while the current one finishes; S/dP run one query block ahead of the gradient GEMMs across item boundaries.
struct Bwd4Smem {
    static constexpr int oKV = 0;                             // [2][K | V]
    static constexpr int oQdO = 2 * 2 * AT_TILE;              // [2 stages][Q | dO]
    static constexpr int oP = oQdO + 2 * 2 * AT_TILE;         // 2 atoms
    static constexpr int oDS = oP + 2 * AT_TILE;              // 2 atoms
    static constexpr int oBar = oDS + 2 * AT_TILE;
    static constexpr int kBytes = oBar + 256 + 1024;
};

// kMode as in the forward kernel: 0 = r1e code (run-time dropout branch after the exponentials of each 16-column chunk), 1 / 2 = dropout
// on / off fixed at compile time with each 4-key group's hash next to its exponentials.
template <int kMode>
__global__ void __maxnreg__(96)
attn_bwd_tc4_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const float* __restrict__ lse,
                    const float* __restrict__ delta, bf16* __restrict__ dqkv, float* __restrict__ dq_acc, int T, int Tm, int H, int BH, float scale,
                    DropCfg drop) {
    // T / Tm as in the forward kernel: strides from T, queries and keys 0 .. Tm - 1
    using S = Bwd4Smem;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::oBar);
    uint64_t* kv_full = bars;                 // [2]
    uint64_t* kv_empty = bars + 2;            // [2] commit after the item's last gradient MMA
    uint64_t* qdo_full = bars + 4;            // [2]
    uint64_t* qdo_empty = bars + 6;           // [2]
    uint64_t* sdp_full = bars + 8;            // commit
    uint64_t* sdp_empty = bars + 9;           // 16
    uint64_t* pds_full = bars + 10;           // 16
    uint64_t* pds_empty = bars + 11;          // commit
    uint64_t* dq_full = bars + 12;            // commit
    uint64_t* dq_empty = bars + 13;           // 16
    uint64_t* dkv_empty = bars + 14;          // 16: dK / dV of the finished item have been read out of TMEM
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = H * 64, ld3 = 3 * d;
    const int nq = (Tm + AT_BM - 1) / AT_BM;
    const int total = nq * BH;
    // head-major items dealt out in rotating rounds (see the forward kernel): concurrent CTAs share a few heads, so Q / dO / K / V and the
    // fp32 dQ accumulator rows they red.add into stay in L2 (with key-block-major items the accumulator traffic went to HBM)
    const int rounds = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const FastDiv fd_grid = fastdiv_make(gridDim.x), fd_nq = fastdiv_make((uint32_t)nq), fd_H = fastdiv_make((uint32_t)H);
    auto item_at = [&](int n) { if (n >= rounds) return -1; const int idx = n * (int)gridDim.x + (int)fastmod(blockIdx.x + n, fd_grid); return idx < total ? idx : -1; };
    auto item_jb = [&](int idx) { return (int)fastmod((uint32_t)idx, fd_nq); };
    auto item_bh = [&](int idx) { return (int)fastdiv((uint32_t)idx, fd_nq); };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
            mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
        }
        mbar_init(sdp_full, 1); mbar_init(sdp_empty, 16);
        mbar_init(pds_full, 16); mbar_init(pds_empty, 1);
        mbar_init(dq_full, 1); mbar_init(dq_empty, 16);
        mbar_init(dkv_empty, 16);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        uint32_t c = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int jb = item_jb(idx), bh = item_bh(idx), b = (int)fastdiv((uint32_t)bh, fd_H), h = bh - b * H;
            const int row_base = b * T, k0 = jb * AT_BN, nit = nq - jb;
            mbar_wait_relaxed(&kv_empty[n & 1], ((n >> 1) & 1) ^ 1);
            uint8_t* sk = smem + S::oKV + (n & 1) * 2 * AT_TILE;
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[n & 1], 2 * AT_TILE);
                tma_load_2d(sk, &tmQKV, &kv_full[n & 1], d + h * 64, row_base + k0);
                tma_load_2d(sk + AT_TILE, &tmQKV, &kv_full[n & 1], 2 * d + h * 64, row_base + k0);
            }
            __syncwarp();
            for (int it = 0; it < nit; ++it, ++c) {
                const int st = c & 1, i = jb + it;
                mbar_wait_relaxed(&qdo_empty[st], ((c >> 1) & 1) ^ 1);
                uint8_t* sq = smem + S::oQdO + st * 2 * AT_TILE;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&qdo_full[st], 2 * AT_TILE);
                    tma_load_2d(sq, &tmQKV, &qdo_full[st], h * 64, row_base + i * AT_BM);
                    tma_load_2d(sq + AT_TILE, &tmDO, &qdo_full[st], h * 64, row_base + i * AT_BM);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);    // S, dP
        constexpr uint32_t idesc_g = make_idesc_bf16(128, 64, true, true);       // dV, dK
        constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);      // dQ
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t dP_mn = desc_mnmajor(smem_base + S::oP, 0), dDS_mn = desc_mnmajor(smem_base + S::oDS, 0);
        const uint64_t dDS_k = desc_kmajor(smem_base + S::oDS, 0);
        // S / dP cursor: one query block ahead of the gradient MMAs
        int n1 = 0, it1 = 0, idx1 = item_at(0), nit1 = idx1 >= 0 ? nq - item_jb(idx1) : 0;
        uint32_t c1 = 0;
        auto issue_next_sdp = [&]() {
            if (idx1 < 0) return;
            if (it1 == 0) mbar_wait(&kv_full[n1 & 1], (n1 >> 1) & 1);
            mbar_wait(&qdo_full[c1 & 1], (c1 >> 1) & 1);
            mbar_wait(sdp_empty, (c1 & 1) ^ 1);                 // S / dP of iteration c1 - 1 are in registers
            tc_fence_after();
            const uint32_t sKV = smem_base + S::oKV + (n1 & 1) * 2 * AT_TILE;
            const uint32_t sQ = smem_base + S::oQdO + (c1 & 1) * 2 * AT_TILE;
            const uint64_t dK_k = desc_kmajor(sKV, 0), dV_k = desc_kmajor(sKV + AT_TILE, 0);
            const uint64_t dQ_k = desc_kmajor(sQ, 0), dDO_k = desc_kmajor(sQ + AT_TILE, 0);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tS, dQ_k + 2 * k, dK_k + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tDP, dDO_k + 2 * k, dV_k + 2 * k, idesc_s, k > 0 ? 1u : 0u);
                umma_commit(sdp_full);
            }
            __syncwarp();
            ++c1; ++it1;
            if (it1 == nit1) { ++n1; idx1 = item_at(n1); it1 = 0; nit1 = idx1 >= 0 ? nq - item_jb(idx1) : 0; }
        };
        issue_next_sdp();
        uint32_t c = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int nit = nq - item_jb(idx);
            const uint32_t sKV = smem_base + S::oKV + (n & 1) * 2 * AT_TILE;
            const uint64_t dK_mn = desc_mnmajor(sKV, 0);
            for (int it = 0; it < nit; ++it, ++c) {
                issue_next_sdp();
                mbar_wait(pds_full, c & 1);
                mbar_wait(dq_empty, (c & 1) ^ 1);                   // dQ of iteration c - 1 has been read out
                if (it == 0 && n > 0) mbar_wait(dkv_empty, (n - 1) & 1);     // previous item's dK / dV have been read out
                tc_fence_after();
                const uint32_t sQ = smem_base + S::oQdO + (c & 1) * 2 * AT_TILE;
                const uint64_t dQ_mn = desc_mnmajor(sQ, 0), dDO_mn = desc_mnmajor(sQ + AT_TILE, 0);
                const bool last = (it == nit - 1);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16(tDV, dP_mn + (uint64_t)(k * 128), dDO_mn + (uint64_t)(k * 128), idesc_g, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16(tDK, dDS_mn + (uint64_t)(k * 128), dQ_mn + (uint64_t)(k * 128), idesc_g, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16(tDQ, dDS_k + (uint64_t)((k >> 2) * (AT_TILE >> 4) + (k & 3) * 2), dK_mn + (uint64_t)(k * 128), idesc_q, k > 0 ? 1u : 0u);
                    umma_commit(dq_full);
                    umma_commit(pds_empty);
                    umma_commit(&qdo_empty[c & 1]);
                    if (last) umma_commit(&kv_empty[n & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        const int quad = warp & 3;
        const int qtr = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * kLog2eF;
        const uint32_t sP = smem_u32(smem + S::oP), sDS = smem_u32(smem + S::oDS);
        const uint32_t addc = attn_drop_addc(drop.thresh16);
        const float inv_c = 1.0f / drop.scale;                // drop.scale = 1 without dropout
        const float out_scale = scale * drop.scale;           // dK (and dQ, in attn_dq_convert) carry the factor c the math warps leave out
        uint32_t t_off[4];                                    // this thread's four 16-byte chunks inside a [128 x 128] bf16 operand buffer
#pragma unroll
        for (int g = 0; g < 4; ++g) t_off[g] = tile_chunk_addr(0u, r, qtr * 4 + g);

        // dQ tile of global iteration cc -> fp32 accumulator rows starting at dst (nullptr: row beyond T)
        auto dq_out = [&](uint32_t cc, float* dst) {
            mbar_wait(dq_full, cc & 1);
            tc_fence_after();
            uint32_t v[16];
            __syncwarp();
            tmem_ld_32x16(tDQ + lane_off + qtr * 16, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);
            if (dst != nullptr) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g), "f"(__uint_as_float(v[4 * g])),
                                 "f"(__uint_as_float(v[4 * g + 1])), "f"(__uint_as_float(v[4 * g + 2])), "f"(__uint_as_float(v[4 * g + 3])) : "memory");
            }
        };

        // lse / delta of this thread's query row are fetched ONE query block ahead (across item boundaries too): loaded right before use
        // they cost an L2 round trip per block pair (r1h profile: 4.5 % of all stall samples on the first use of lse)
        // dK (x softmax scale), dV (x dropout scale) of item idx_: TMEM -> registers (then the accumulators are free) -> bf16 rows of dqkv
        auto dkv_out = [&](int idx_) {
            const int jb_ = item_jb(idx_), bh_ = item_bh(idx_), b_ = (int)fastdiv((uint32_t)bh_, fd_H), h_ = bh_ - b_ * H;
            const int kj = jb_ * AT_BN + r;
            bf16* dkp = dqkv + (size_t)(b_ * T + min(kj, Tm - 1)) * ld3 + d + h_ * 64 + qtr * 16;
            bf16* dvp = dkp + d;
            uint32_t a[16], v[16];
            __syncwarp();
            tmem_ld_32x16(tDK + lane_off + qtr * 16, a);
            tmem_ld_32x16(tDV + lane_off + qtr * 16, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dkv_empty);
            if (kj < Tm) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    reinterpret_cast<uint4*>(dkp)[g] =
                        make_uint4(pack_bf16(__uint_as_float(a[8 * g]) * out_scale, __uint_as_float(a[8 * g + 1]) * out_scale),
                                   pack_bf16(__uint_as_float(a[8 * g + 2]) * out_scale, __uint_as_float(a[8 * g + 3]) * out_scale),
                                   pack_bf16(__uint_as_float(a[8 * g + 4]) * out_scale, __uint_as_float(a[8 * g + 5]) * out_scale),
                                   pack_bf16(__uint_as_float(a[8 * g + 6]) * out_scale, __uint_as_float(a[8 * g + 7]) * out_scale));
                    reinterpret_cast<uint4*>(dvp)[g] =
                        make_uint4(pack_bf16(__uint_as_float(v[8 * g]) * drop.scale, __uint_as_float(v[8 * g + 1]) * drop.scale),
                                   pack_bf16(__uint_as_float(v[8 * g + 2]) * drop.scale, __uint_as_float(v[8 * g + 3]) * drop.scale),
                                   pack_bf16(__uint_as_float(v[8 * g + 4]) * drop.scale, __uint_as_float(v[8 * g + 5]) * drop.scale),
                                   pack_bf16(__uint_as_float(v[8 * g + 6]) * drop.scale, __uint_as_float(v[8 * g + 7]) * drop.scale));
                }
            }
        };
        // Version 5 defers an item's tail -- the dQ tile of its last query block and the dK / dV read-out, both of which need the item's
        // last gradient MMAs to have completed -- to after the FIRST block of the next item has been handed to the MMA warp: the wait
        // that was exposed once per item (r1h / r1n profiles: 6.6 % of all stall samples on that one try_wait) now overlaps a whole
        // block of softmax-gradient math.  The MMA warp's dkv_empty wait before the next item's first gradient MMA is unchanged.
        constexpr bool kDefer = true;
        float* prev_dst = nullptr;
        int idx_prev = -1;
        float lse_nx = 0.f, dlt_nx = 0.f;
        auto fetch_row_stats = [&](int bh_, int i_) {
            const int qi_ = i_ * AT_BM + r;
            if (qi_ < Tm) { lse_nx = __ldg(lse + (size_t)bh_ * T + qi_); dlt_nx = __ldg(delta + (size_t)bh_ * T + qi_); }
            else { lse_nx = 0.f; dlt_nx = 0.f; }
        };
        { const int idx0 = item_at(0); if (idx0 >= 0) fetch_row_stats(item_bh(idx0), item_jb(idx0)); }
        uint32_t c = 0;
        for (int n = 0, idx; (idx = item_at(n)) >= 0; ++n) {
            const int jb = item_jb(idx), bh = item_bh(idx), b = (int)fastdiv((uint32_t)bh, fd_H), h = bh - b * H;
            const int row_base = b * T, k0 = jb * AT_BN, nit = nq - jb;
            const int kc0 = k0 + qtr * 32;
            const int idx_nx = item_at(n + 1);
            const int bh_nx = idx_nx >= 0 ? item_bh(idx_nx) : 0, jb_nx = idx_nx >= 0 ? item_jb(idx_nx) : 0;
            if (!kDefer) prev_dst = nullptr;
            for (int it = 0; it < nit; ++it, ++c) {
                const int i = jb + it;
                const uint32_t ph = c & 1;
                const int qi = i * AT_BM + r;
                const bool q_ok = qi < Tm;
                // pinned: without it the compiler rotates this multiply into the previous iteration, right behind the load it was meant to
                // be a whole block away from (r1h profile: 4.5 % of the stall samples on that FMUL)
                float lse_cur = lse_nx, dlt_cur = dlt_nx;
                asm volatile("" : "+f"(lse_cur), "+f"(dlt_cur));
                const float lse2 = lse_cur * kLog2eF;
                const float dlt = dlt_cur;
                if (it + 1 < nit) fetch_row_stats(bh, i + 1);
                else if (idx_nx >= 0) fetch_row_stats(bh_nx, jb_nx);
                const bool need_mask = (i == jb) || (k0 + AT_BN > Tm) || (i * AT_BM + AT_BM > Tm);
                const AttnDropRow rk = attn_drop_row(drop.seed, (uint64_t)bh * (uint64_t)T + (uint64_t)qi);
                float* cur_dst = q_ok ? dq_acc + (size_t)(row_base + qi) * d + h * 64 + qtr * 16 : nullptr;
                uint32_t pp[16], dd[16];
                mbar_wait(sdp_full, ph);
                tc_fence_after();
                // The gradient in the scaling that costs two FMA-pipe operations per element:
                //     dS / c = P o (mask o dP - delta / c)       (c = dropout scale 1 / (1 - p); the factor c is applied to dK and dQ on the way out)
                //     nd = pe * (-delta / c) ;  kept key: fma(pe, dP, nd) ;  dropped key: nd
                // and the keep decision is taken on the PACKED bf16x2 words: P &= mask, dS = select(mask, pack(kept), pack(dropped)) -- two
                // 16-bit lanes per LOP3 (common.cuh attn_drop_mask2) instead of shift + ISETP + 2 SEL per key on fp32 values.
                const float ndl = -dlt * inv_c;
                {
                    // keys of this block that exist for this query row (causal + sequence end); only consulted on diagonal / edge blocks
                    const int n_ok = q_ok ? min(qi, Tm - 1) - kc0 + 1 : 0;
                    const uint32_t zero = drop.thresh16 >> 16;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t sv[16], gv[16];
                        __syncwarp();
                        tmem_ld_32x16(tS + lane_off + qtr * 32 + cc * 16, sv);
                        tmem_ld_32x16(tDP + lane_off + qtr * 32 + cc * 16, gv);
                        tmem_ld_wait();
                        if (cc == 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sdp_empty);
                        }
                        if (need_mask) bwd_math<kMode, true>(cc, sv, gv, pp, dd, sl2, lse2, ndl, rk, (uint32_t)kc0 >> 2, addc, n_ok, zero);
                        else bwd_math<kMode, false>(cc, sv, gv, pp, dd, sl2, lse2, ndl, rk, (uint32_t)kc0 >> 2, addc, 32, zero);
                    }
                }
                mbar_wait(pds_empty, ph ^ 1);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    sts_v4(sP + t_off[g], make_uint4(pp[4 * g], pp[4 * g + 1], pp[4 * g + 2], pp[4 * g + 3]));
                    sts_v4(sDS + t_off[g], make_uint4(dd[4 * g], dd[4 * g + 1], dd[4 * g + 2], dd[4 * g + 3]));
                }
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pds_full);
                if (it > 0) dq_out(c - 1, prev_dst);
                else if (kDefer && n > 0) { dq_out(c - 1, prev_dst); dkv_out(idx_prev); }      // the previous item's tail, under this block's MMAs
                prev_dst = cur_dst;
            }
            if (!kDefer) {
                dq_out(c - 1, prev_dst);                               // last query block of the item; all gradient MMAs are complete after this
                dkv_out(idx);
            }
            idx_prev = idx;
        }
        if (kDefer && c > 0) { dq_out(c - 1, prev_dst); dkv_out(idx_prev); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

// fp32 dQ accumulation buffer [B*T, d] -> bf16 dqkv[:, 0:d]
__global__ void attn_dq_convert_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv, size_t rows, int d, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t n4 = rows * (size_t)d / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const size_t row = e / d;
        const int c = (int)(e - row * d);
        const float4 v = reinterpret_cast<const float4*>(dq_acc)[i];
        *reinterpret_cast<uint2*>(dqkv + row * 3 * d + c) = make_uint2(pack_bf16(v.x * scale, v.y * scale), pack_bf16(v.z * scale, v.w * scale));
    }
}

// TTTS_ATTN_TAIL=1: leave a ragged last query tile of <= 16 rows to attention_tail.cu.  OFF by default -- measured (r2ab / r2ac, B = 32,
// H = 16, T = 1156): the tile kernels do get faster without their tenth, four-row query tile (forward 0.362 -> 0.311 ms, backward
// 0.62 -> 0.53 ms), but four query rows still meet EVERY key: the tail kernels stream K and V (151 MB) once more in the forward and K, V plus
// a read-modify-write of all dK / dV rows (453 MB, 70 us at the HBM roof) in the backward, and came out at 61 / 200 us -- a net loss
// (forward 0.382 vs 0.362 ms, backward 0.814 vs 0.685 ms).  The traffic only disappears if the tail rows are handled while the key / value
// tiles are resident in shared memory, i.e. inside the tile kernels.
bool attn_tail_split() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_ATTN_TAIL"); on = (e && e[0] == '1') ? 1 : 0; }
    return on != 0;
}

bool attn_use_tc() {
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("TTTS_ATTN_LEGACY"); legacy = (e && e[0] == '1') ? 1 : 0; }
    return !legacy;
}

int attn_fwd_tc(const bf16* qkv, bf16* o, float* lse, int B, int T, int H, DropCfg drop, cudaStream_t st) {
    const int d = H * 64;
    CUtensorMap tm;
    int rc = make_tmap_2d(&tm, qkv, 2, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, 128, true);
    if (rc) return rc;
    static bool attr4 = false;
    if (!attr4) {
        TTTS_CUDA(cudaFuncSetAttribute(attn_fwd_tc4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd4Smem::kBytes));
        TTTS_CUDA(cudaFuncSetAttribute(attn_fwd_tc4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd4Smem::kBytes));
        attr4 = true;
    }
    // opt-in (see attn_tail_split): a ragged tail of <= 16 rows (T = 1156 = 9 x 128 + 4 in training) is left to attention_tail.cu, the tile kernel stops at Tm
    const int tail = attn_tail_split() ? attn_tail_rows(T) : 0;
    const int Tm = T - tail;
    const int nq = (Tm + AT_BM - 1) / AT_BM;
    const int items = nq * B * H;
    const int nblk = items < num_sms() ? items : num_sms();
    TTTS_CHECK_ARG((uint64_t)(items + nblk) * (uint64_t)(nblk > H ? nblk : H) < (1ull << 32) && nq <= 4096, "attention: too many (block, head) items");
    if (drop.thresh16) TTTS_CUDA(launch_pdl(attn_fwd_tc4_kernel<1>, dim3(nblk), dim3(AT3_THREADS), Fwd4Smem::kBytes, st, tm, o, lse, T, Tm, H, B * H, 0.125f, drop));
    else TTTS_CUDA(launch_pdl(attn_fwd_tc4_kernel<2>, dim3(nblk), dim3(AT3_THREADS), Fwd4Smem::kBytes, st, tm, o, lse, T, Tm, H, B * H, 0.125f, drop));
    TTTS_LAUNCH_CHECK("attn_fwd_tc");
    if (tail) TTTS_RUN(attn_tail_fwd(qkv, o, lse, B, T, H, Tm, drop.thresh16, drop.scale, drop.seed, st));
    return TTTS_OK;
}

// defined in attention.cu
int attn_delta(const bf16* o, const bf16* dout, float* delta, int B, int T, int H, cudaStream_t st);

int attn_bwd_tc(const bf16* qkv, const bf16* o, const bf16* dout, const float* lse, float* delta, bf16* dqkv, int B, int T, int H, DropCfg drop,
                cudaStream_t st) {
    const int d = H * 64;
    float* dq_acc = delta + (((size_t)B * H * T + 63) / 64) * 64;       // scratch layout: delta | dq_acc[B*T, d]
    TTTS_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * T * d * sizeof(float), st));
    int rc = attn_delta(o, dout, delta, B, T, H, st);
    if (rc) return rc;
    CUtensorMap tmQ, tmDO;
    rc = make_tmap_2d(&tmQ, qkv, 2, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d, 64, 128, true);
    if (rc) return rc;
    rc = make_tmap_2d(&tmDO, dout, 2, (uint64_t)d, (uint64_t)B * T, (uint64_t)d, 64, 128, true);
    if (rc) return rc;
    static bool attr4 = false;
    if (!attr4) {
        TTTS_CUDA(cudaFuncSetAttribute(attn_bwd_tc4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd4Smem::kBytes));
        TTTS_CUDA(cudaFuncSetAttribute(attn_bwd_tc4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd4Smem::kBytes));
        attr4 = true;
    }
    const int tail = attn_tail_split() ? attn_tail_rows(T) : 0;
    const int Tm = T - tail;
    const int nkb = (Tm + AT_BN - 1) / AT_BN;
    const int items = nkb * B * H;
    const int nblk = items < num_sms() ? items : num_sms();
    TTTS_CHECK_ARG((uint64_t)(items + nblk) * (uint64_t)(nblk > H ? nblk : H) < (1ull << 32) && nkb <= 4096, "attention: too many (block, head) items");
    if (drop.thresh16) TTTS_CUDA(launch_pdl(attn_bwd_tc4_kernel<1>, dim3(nblk), dim3(AT3_THREADS), Bwd4Smem::kBytes, st, tmQ, tmDO, lse, delta, dqkv, dq_acc, T, Tm, H, B * H, 0.125f, drop));
    else TTTS_CUDA(launch_pdl(attn_bwd_tc4_kernel<2>, dim3(nblk), dim3(AT3_THREADS), Bwd4Smem::kBytes, st, tmQ, tmDO, lse, delta, dqkv, dq_acc, T, Tm, H, B * H, 0.125f, drop));
    TTTS_LAUNCH_CHECK("attn_bwd_tc");
    // the tail rows' dQ (into dq_acc) and their share of every dK / dV row (added to what the tile kernel has just written)
    if (tail) TTTS_RUN(attn_tail_bwd(qkv, dout, lse, delta, dqkv, dq_acc, B, T, H, Tm, drop.thresh16, drop.scale, drop.seed, st));
    const size_t n4 = (size_t)B * T * d / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    TTTS_CUDA(launch_pdl(attn_dq_convert_kernel, dim3(blocks), dim3(256), 0, st, dq_acc, dqkv, (size_t)B * T, d, 0.125f * drop.scale));     // dQ = scale * c * (dS / c) K
    TTTS_LAUNCH_CHECK("attn_dq_convert");
    return TTTS_OK;
}

}  // namespace ttts
