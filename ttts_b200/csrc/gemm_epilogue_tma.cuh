// Warp-collective GEMM epilogue that moves every tile through the TMA engine (round-1 "r1h" profile: the register/LSU epilogue spent
// 580 warp instructions per 32x32 chunk -- address arithmetic, bounds predicates, two staging transposes -- and, for the residual
// variant, spilled its prefetched rows to local memory, so the K = 1024 GEMMs ran at the epilogue's pace: 468 TFLOP/s for the attention
// c_proj, 931 for c_fc+GELU, 845 for the GELU' dgrad against 1 400 - 1 640 for the K >= 3072 shapes, profiles/r1h_gemm_step_table.txt).
//
// A "chunk" is what one tcgen05.ld.32x32b.x32 delivers: 32 rows (one per lane) x 32 accumulator columns.  Each epilogue warp owns two
// 2 KB staging units in shared memory; a unit is a [32 rows x 64 B] box in the 64-byte TMA swizzle, i.e. 32 bf16 or 16 fp32 columns:
//   outputs   : lane writes its row (4 x st.shared.v4, conflict-free under the swizzle) -> fence.proxy.async -> lane 0 issues ONE
//               cp.async.bulk.tensor store (or cp.reduce...add for the split-K weight gradients) -> commit_group
//   aux inputs: the fp32 residual chunk (RESID) / bf16 pre-activation chunk (DGELU) arrive by cp.async.bulk.tensor loads on a per-warp
//               mbarrier, issued one chunk ahead, so no register holds prefetched data
// Ragged rows (M % 128) need no special path: the TMA unit clips stores and zero-fills loads along the outer dimension.  A ragged
// COLUMN tail (N % 32: the two heads) goes through the per-thread path of gemm_epilogue.cuh: a TMA store whose in-bounds width is not
// a multiple of 16 bytes also overwrites the rest of that 16-byte unit (measured: the pad columns of a [*, 1026] bf16 output were hit).
#pragma once
#include "gemm_epilogue.cuh"

namespace ttts {

constexpr int EU_BYTES = 2048;                 // one staging unit
constexpr int EPI_WARP_BYTES = 2 * EU_BYTES;   // per epilogue warp

TTTS_DEVICE void st_sh_v4(uint32_t a, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
TTTS_DEVICE uint4 ld_sh_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }

TTTS_DEVICE void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
TTTS_DEVICE void tma_reduce_add_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
TTTS_DEVICE void tma_load_2d_s(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
TTTS_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> TTTS_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> TTTS_DEVICE void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// 16-byte piece j (0..3) of row r inside a unit: CU_TENSOR_MAP_SWIZZLE_64B XORs address bits [7,9) into bits [4,6)
TTTS_DEVICE uint32_t eu_addr(uint32_t U, int r, int j) { return U + r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }
TTTS_DEVICE void eu_put_row(uint32_t U, int lane, const uint4 (&q)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) st_sh_v4(eu_addr(U, lane, j), q[j]);
}
TTTS_DEVICE void eu_get_row(uint32_t U, int lane, uint4 (&q)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = ld_sh_v4(eu_addr(U, lane, j));
}

TTTS_DEVICE void pack16(const float* v, uint4 (&q)[4]) {     // 32 floats -> 32 bf16 = 64 B
#pragma unroll
    for (int j = 0; j < 4; ++j)
        q[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                          pack_bf16(v[8 * j + 6], v[8 * j + 7]));
}
TTTS_DEVICE void f32x16(const float* v, uint4 (&q)[4]) {     // 16 floats = 64 B
#pragma unroll
    for (int j = 0; j < 4; ++j)
        q[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
}

struct EpiTmaCtx {
    const CUtensorMap* tm_out;
    const CUtensorMap* tm_aux;       // RESID: fp32 residual ; DGELU: bf16 pre-activation
    const CUtensorMap* tm_aux_out;   // GELU: bf16 pre-activation output
    uint32_t U;                      // shared-memory address of this warp's two staging units
    uint64_t* ldbar;                 // this warp's aux-load barrier (count 1)
    uint32_t ld_phase;               // completions consumed so far
    uint32_t nstore;                 // BF16: unit toggle
};

template <int EPI>
TTTS_DEVICE bool epi_has_load(const GemmParams& p) {
    const int epi = EPI >= 0 ? EPI : p.epi;
    return epi == TTTS_EPI_RESID || epi == TTTS_EPI_DGELU;
}

// L2 prefetch of the aux input this warp will need for a whole tile (32 rows x 128 columns), one row per lane, issued one tile ahead
// so that the chunk-by-chunk TMA loads below hit L2 instead of waiting for HBM inside the epilogue's critical path.  Plain
// prefetch.global.L2 (LSU path): the bulk form goes through the TMA unit's queue, where 256 extra requests per tile delayed the main
// loop's operand loads (measured: every aux-loading GEMM got 4 - 11 % slower with cp.async.bulk.prefetch.L2).
template <int EPI>
TTTS_DEVICE void epi_l2_prefetch(const GemmParams& p, const int row0, const int colw, const int lane) {
    const int epi = EPI >= 0 ? EPI : p.epi;
    if (epi != TTTS_EPI_RESID && epi != TTTS_EPI_DGELU) return;
    if (!p.l2pf || row0 + lane >= p.M || colw + 128 > p.N) return;
    const int eb = epi == TTTS_EPI_RESID ? 4 : 2;
    const uint8_t* g = reinterpret_cast<const uint8_t*>(p.aux) + ((size_t)(row0 + lane) * p.ldaux + colw) * eb;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i * 128 < 128 * eb) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + i * 128) : "memory");
}

// Issue the aux-input load of chunk (row0, col0) into the staging units.  Warp-collective (all lanes call; lane 0 issues).  For DGELU
// the caller guarantees every lane has read the previous chunk out of unit 0; for RESID the units were last read by this warp's own
// stores, which lane 0 waits for here.
template <int EPI>
TTTS_DEVICE void epi_tma_issue_load(const GemmParams& p, EpiTmaCtx& c, const int row0, const int col0, const int lane) {
    const int epi = EPI >= 0 ? EPI : p.epi;
    if (epi != TTTS_EPI_RESID && epi != TTTS_EPI_DGELU) return;
    if (row0 >= p.M || col0 + 32 > p.N) return;                   // warp-uniform; epi_tma_apply takes no TMA load for the same chunks
    if (lane == 0) {
        if (epi == TTTS_EPI_RESID) {
            bulk_wait_read<0>();
            mbar_arrive_expect_tx(c.ldbar, 2 * EU_BYTES);
            tma_load_2d_s(c.U, c.tm_aux, c.ldbar, col0, row0);
            tma_load_2d_s(c.U + EU_BYTES, c.tm_aux, c.ldbar, col0 + 16, row0);
        } else {
            mbar_arrive_expect_tx(c.ldbar, EU_BYTES);
            tma_load_2d_s(c.U, c.tm_aux, c.ldbar, col0, row0);
        }
    }
    __syncwarp();
}

// Warp-collective.  row0 = first row of this warp's 32-row slab, col0 = first column of the chunk, cidx = chunk index (0..3) inside
// this warp's 128 columns (selects the lanes holding the chunk's bias), next_col0 = first column of the chunk this warp handles next
// in the same tile (-1: none) so that its aux load can be issued as early as the staging units allow.
template <int EPI>
TTTS_DEVICE void epi_tma_apply(const GemmParams& p, EpiTmaCtx& c, const int row0, const int col0, const int lane, const uint32_t (&r)[32],
                               const float4 bq, const int cidx, const int next_col0) {
    const int epi = EPI >= 0 ? EPI : p.epi;
    if (row0 >= p.M || col0 >= p.N) return;                       // warp-uniform
    const int row = row0 + lane;
    if (col0 + 32 > p.N) {                                        // ragged column tail (heads only): per-thread path, global bias / aux
        EpiAux dummy;
        epi_apply(p, row, col0, r, nullptr, dummy, EPI);
        return;
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (epi != TTTS_EPI_F32_ADD && p.bias != nullptr) {
        // bias of column (cidx * 32 + j) of this warp's 128 lives in lane cidx * 8 + j / 4, component j % 4
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const int src = cidx * 8 + j4;
            v[4 * j4] += __shfl_sync(0xffffffffu, bq.x, src);
            v[4 * j4 + 1] += __shfl_sync(0xffffffffu, bq.y, src);
            v[4 * j4 + 2] += __shfl_sync(0xffffffffu, bq.z, src);
            v[4 * j4 + 3] += __shfl_sync(0xffffffffu, bq.w, src);
        }
    }
    uint4 q[4];
    switch (epi) {
    case TTTS_EPI_BF16: {
        const uint32_t U = c.U + (c.nstore & 1u) * EU_BYTES;
        ++c.nstore;
        pack16(v, q);
        if (lane == 0) bulk_wait_read<1>();                       // the store issued two chunks ago has read this unit
        __syncwarp();
        eu_put_row(U, lane, q);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(c.tm_out, U, col0, row0); bulk_commit(); }
    } break;
    case TTTS_EPI_GELU: {
        // pre = bf16(acc + bias) ; h = gelu_new(pre) in packed bf16x2 (common.cuh)
        uint4 qp[4];
        pack16(v, qp);
        {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(qp);
            uint32_t* wh = reinterpret_cast<uint32_t*>(q);
#pragma unroll
            for (int t = 0; t < 16; ++t) wh[t] = gelu_new_bf2(wp[t]);
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        eu_put_row(c.U, lane, q);
        if (p.aux_out) eu_put_row(c.U + EU_BYTES, lane, qp);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(c.tm_out, c.U, col0, row0);
            if (p.aux_out) tma_store_2d(c.tm_aux_out, c.U + EU_BYTES, col0, row0);
            bulk_commit();
        }
    } break;
    case TTTS_EPI_RESID: {
        // x_out = x_in + dropout(bf16(acc + bias))     (HF: modeling_gpt2.py:224,282 / 242,307); the residual chunk is added in place
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
        if (p.drop_thresh16) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const uint64_t e4 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(col0 + 4 * j4)) >> 2;
                const uint64_t bits = dropout_bits4(p.drop_seed, e4);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * j4 + j] = dropout_keep(bits, j, p.drop_thresh16) ? v[4 * j4 + j] * p.drop_scale : 0.f;
            }
        }
        mbar_wait(c.ldbar, c.ld_phase & 1u);
        ++c.ld_phase;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const uint32_t U = c.U + hh * EU_BYTES;
            eu_get_row(U, lane, q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                q[j].x = __float_as_uint(__uint_as_float(q[j].x) + v[16 * hh + 4 * j]);
                q[j].y = __float_as_uint(__uint_as_float(q[j].y) + v[16 * hh + 4 * j + 1]);
                q[j].z = __float_as_uint(__uint_as_float(q[j].z) + v[16 * hh + 4 * j + 2]);
                q[j].w = __float_as_uint(__uint_as_float(q[j].w) + v[16 * hh + 4 * j + 3]);
            }
            eu_put_row(U, lane, q);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(c.tm_out, c.U, col0, row0);
            tma_store_2d(c.tm_out, c.U + EU_BYTES, col0 + 16, row0);
            bulk_commit();
        }
        if (next_col0 >= 0) epi_tma_issue_load<EPI>(p, c, row0, next_col0, lane);
    } break;
    case TTTS_EPI_DGELU: {
        mbar_wait(c.ldbar, c.ld_phase & 1u);
        ++c.ld_phase;
        eu_get_row(c.U, lane, q);
        const uint32_t w[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
        __syncwarp();                                             // every lane has its pre-activation row: unit 0 may be refilled
        if (next_col0 >= 0) epi_tma_issue_load<EPI>(p, c, row0, next_col0, lane);
        // out = bf16(acc) * gelu'(pre), both factors packed bf16x2 (the reference's autocast backward rounds the dgrad output and every
        // elementwise factor to bf16 as well)
        pack16(v, q);
        {
            uint32_t* wo = reinterpret_cast<uint32_t*>(q);
#pragma unroll
            for (int t = 0; t < 16; ++t) wo[t] = bf2_mul(wo[t], gelu_new_grad_bf2(w[t]));
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        eu_put_row(c.U + EU_BYTES, lane, q);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(c.tm_out, c.U + EU_BYTES, col0, row0); bulk_commit(); }
    } break;
    default: {  // TTTS_EPI_F32_ADD (split-K reduction through the TMA unit) / TTTS_EPI_F32
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        f32x16(v, q);
        eu_put_row(c.U, lane, q);
        f32x16(v + 16, q);
        eu_put_row(c.U + EU_BYTES, lane, q);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (epi == TTTS_EPI_F32_ADD) {
                tma_reduce_add_2d(c.tm_out, c.U, col0, row0);
                tma_reduce_add_2d(c.tm_out, c.U + EU_BYTES, col0 + 16, row0);
            } else {
                tma_store_2d(c.tm_out, c.U, col0, row0);
                tma_store_2d(c.tm_out, c.U + EU_BYTES, col0 + 16, row0);
            }
            bulk_commit();
        }
    } break;
    }
}

}  // namespace ttts
