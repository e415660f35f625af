// Warp-collective, shared-memory-staged GEMM epilogue: the thread-per-row register tile that tcgen05.ld produces is transposed
// through a small per-warp staging buffer so that every global load / store / red instruction touches 8 rows x 64 contiguous
// bytes (8 LSU wavefronts) instead of 32 rows x 16 B (32 wavefronts).  profiles/r1_notes.md: the un-staged epilogue was
// LSU-wavefront-bound (GELU: ~8k cycles per 128x256 tile, as long as the K=1024 main loop).
//
// A "chunk" is 32 rows (one per lane) x 32 columns.  Staging buffer per warp: 32 rows x 80 B (64 B payload + 16 B pad).
#pragma once
#include "gemm_epilogue.cuh"

namespace ttts {

constexpr int ST_ROW = 80;                   // bytes per staged row
constexpr int ST_BYTES = 32 * ST_ROW;        // per warp

TTTS_DEVICE void st_sh_v4(uint32_t a, uint4 v) { asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
TTTS_DEVICE uint4 ld_sh_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }

// registers of the prefetched auxiliary rows, held in the COALESCED mapping: piece i of this lane = row (lane>>2)+8i, 16-byte column slot (lane&3)
struct EpiAuxC { uint4 q[8]; };      // RESID: two 64-byte rounds x 4 pieces ; DGELU: one round (q[0..3])

// coalesced load of a [32 rows x 64 B] block starting at byte address gbase (row stride ld_bytes); rows >= nrows are skipped
TTTS_DEVICE void gload64(const uint8_t* gbase, size_t ld_bytes, int nrows, int lane, uint4* dst) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        dst[i] = (r < nrows) ? *reinterpret_cast<const uint4*>(gbase + (size_t)r * ld_bytes + (lane & 3) * 16) : make_uint4(0, 0, 0, 0);
    }
}
// pieces (coalesced mapping) -> staging
TTTS_DEVICE void stage_put_pieces(uint32_t S, int lane, const uint4* src) {
#pragma unroll
    for (int i = 0; i < 4; ++i) st_sh_v4(S + ((lane >> 2) + 8 * i) * ST_ROW + (lane & 3) * 16, src[i]);
}
// staging -> coalesced global store / red of a [32 x 64 B] block
TTTS_DEVICE void stage_store64(uint32_t S, uint8_t* gbase, size_t ld_bytes, int nrows, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        const uint4 v = ld_sh_v4(S + r * ST_ROW + (lane & 3) * 16);
        if (r < nrows) *reinterpret_cast<uint4*>(gbase + (size_t)r * ld_bytes + (lane & 3) * 16) = v;
    }
}
TTTS_DEVICE void stage_red64_f32(uint32_t S, float* gbase, size_t ld_elems, int nrows, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        const uint4 v = ld_sh_v4(S + r * ST_ROW + (lane & 3) * 16);
        if (r < nrows)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + (size_t)r * ld_elems + (lane & 3) * 4), "f"(__uint_as_float(v.x)),
                         "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
    }
}
// my row (lane) <-> staging, 64 bytes
TTTS_DEVICE void stage_put_row(uint32_t S, int lane, const uint4 (&q)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) st_sh_v4(S + lane * ST_ROW + j * 16, q[j]);
}
TTTS_DEVICE void stage_get_row(uint32_t S, int lane, uint4 (&q)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = ld_sh_v4(S + lane * ST_ROW + j * 16);
}

TTTS_DEVICE void pack16(const float* v, uint4 (&q)[4]) {     // 32 floats -> 32 bf16 = 64 B
#pragma unroll
    for (int j = 0; j < 4; ++j)
        q[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                          pack_bf16(v[8 * j + 6], v[8 * j + 7]));
}

// issue the global loads the NEXT chunk needs (coalesced mapping); a no-op for epilogues without an auxiliary input
// EPI >= 0: epilogue fixed at compile time (hot shapes get their own kernel: no switch, no registers for the other variants); -1: p.epi
template <int EPI>
TTTS_DEVICE void epi_prefetch_c(const GemmParams& p, int row0, int col0, int lane, EpiAuxC& x) {
    const int epi = EPI >= 0 ? EPI : p.epi;
    if (epi != TTTS_EPI_RESID && epi != TTTS_EPI_DGELU) return;
    if (row0 >= p.M || col0 + 32 > p.N) return;
    const int nrows = p.M - row0;
    if (epi == TTTS_EPI_RESID) {
        const uint8_t* g = reinterpret_cast<const uint8_t*>(reinterpret_cast<const float*>(p.aux) + (size_t)row0 * p.ldaux + col0);
        gload64(g, (size_t)p.ldaux * 4, nrows, lane, x.q);
        gload64(g + 64, (size_t)p.ldaux * 4, nrows, lane, x.q + 4);
    } else {
        const uint8_t* g = reinterpret_cast<const uint8_t*>(reinterpret_cast<const bf16*>(p.aux) + (size_t)row0 * p.ldaux + col0);
        gload64(g, (size_t)p.ldaux * 2, nrows, lane, x.q);
    }
}

// Warp-collective: all 32 lanes must call it (it contains __syncwarp).  row0 = first row of this warp's 32-row slab.
template <int EPI>
TTTS_DEVICE void epi_apply_staged(const GemmParams& p, const int row0, const int col0, const int lane, const uint32_t (&r)[32], const float* sbias,
                                  const EpiAuxC& x, const uint32_t S) {
    if (row0 >= p.M || col0 >= p.N) return;                     // warp-uniform
    const int row = row0 + lane;
    const int epi = EPI >= 0 ? EPI : p.epi;
    if (col0 + 32 > p.N) {                                        // ragged column tail (heads only): simple per-thread path
        EpiAux dummy;
        epi_apply(p, row, col0, r, sbias, dummy, EPI);
        return;
    }
    const int nrows = p.M - row0;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (epi != TTTS_EPI_F32_ADD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    uint4 q[4];
    switch (epi) {
    case TTTS_EPI_BF16: {
        pack16(v, q);
        stage_put_row(S, lane, q);
        __syncwarp();
        stage_store64(S, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(p.out) + (size_t)row0 * p.ldo + col0), (size_t)p.ldo * 2, nrows, lane);
        __syncwarp();
    } break;
    case TTTS_EPI_GELU: {
        // pre = bf16(acc + bias) ; h = gelu_new(pre) in packed bf16x2 (common.cuh): 3 instructions per element instead of 13
        uint4 qp[4];
        pack16(v, qp);
        {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(qp);
            uint32_t* wh = reinterpret_cast<uint32_t*>(q);
#pragma unroll
            for (int t = 0; t < 16; ++t) wh[t] = gelu_new_bf2(wp[t]);
        }
        stage_put_row(S, lane, q);
        __syncwarp();
        stage_store64(S, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(p.out) + (size_t)row0 * p.ldo + col0), (size_t)p.ldo * 2, nrows, lane);
        __syncwarp();
        if (p.aux_out) {
            stage_put_row(S, lane, qp);
            __syncwarp();
            stage_store64(S, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(p.aux_out) + (size_t)row0 * p.ldaux_out + col0), (size_t)p.ldaux_out * 2, nrows, lane);
            __syncwarp();
        }
    } break;
    case TTTS_EPI_RESID: {
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
        float* obase = reinterpret_cast<float*>(p.out) + (size_t)row0 * p.ldo + col0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                          // two rounds of 16 fp32 columns (64 B per row)
            stage_put_pieces(S, lane, x.q + 4 * hh);              // residual rows, loaded coalesced earlier
            __syncwarp();
            stage_get_row(S, lane, q);
            const uint32_t xs[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                q[j] = make_uint4(__float_as_uint(__uint_as_float(xs[4 * j]) + v[16 * hh + 4 * j]), __float_as_uint(__uint_as_float(xs[4 * j + 1]) + v[16 * hh + 4 * j + 1]),
                                  __float_as_uint(__uint_as_float(xs[4 * j + 2]) + v[16 * hh + 4 * j + 2]), __float_as_uint(__uint_as_float(xs[4 * j + 3]) + v[16 * hh + 4 * j + 3]));
            __syncwarp();
            stage_put_row(S, lane, q);
            __syncwarp();
            stage_store64(S, reinterpret_cast<uint8_t*>(obase + 16 * hh), (size_t)p.ldo * 4, nrows, lane);
            __syncwarp();
        }
    } break;
    case TTTS_EPI_DGELU: {
        stage_put_pieces(S, lane, x.q);
        __syncwarp();
        stage_get_row(S, lane, q);
        __syncwarp();
        const uint32_t w[16] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w, q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w};
        // out = bf16(acc) * gelu'(pre), both factors packed bf16x2 (the reference's autocast backward rounds the dgrad output and every
        // elementwise factor to bf16 as well)
        pack16(v, q);
        {
            uint32_t* wo = reinterpret_cast<uint32_t*>(q);
#pragma unroll
            for (int t = 0; t < 16; ++t) wo[t] = bf2_mul(wo[t], gelu_new_grad_bf2(w[t]));
        }
        stage_put_row(S, lane, q);
        __syncwarp();
        stage_store64(S, reinterpret_cast<uint8_t*>(reinterpret_cast<bf16*>(p.out) + (size_t)row0 * p.ldo + col0), (size_t)p.ldo * 2, nrows, lane);
        __syncwarp();
    } break;
    case TTTS_EPI_F32_ADD: {
        float* obase = reinterpret_cast<float*>(p.out) + (size_t)row0 * p.ldo + col0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                q[j] = make_uint4(__float_as_uint(v[16 * hh + 4 * j]), __float_as_uint(v[16 * hh + 4 * j + 1]), __float_as_uint(v[16 * hh + 4 * j + 2]),
                                  __float_as_uint(v[16 * hh + 4 * j + 3]));
            stage_put_row(S, lane, q);
            __syncwarp();
            stage_red64_f32(S, obase + 16 * hh, (size_t)p.ldo, nrows, lane);
            __syncwarp();
        }
    } break;
    default: {  // TTTS_EPI_F32
        float* obase = reinterpret_cast<float*>(p.out) + (size_t)row0 * p.ldo + col0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                q[j] = make_uint4(__float_as_uint(v[16 * hh + 4 * j]), __float_as_uint(v[16 * hh + 4 * j + 1]), __float_as_uint(v[16 * hh + 4 * j + 2]),
                                  __float_as_uint(v[16 * hh + 4 * j + 3]));
            stage_put_row(S, lane, q);
            __syncwarp();
            stage_store64(S, reinterpret_cast<uint8_t*>(obase + 16 * hh), (size_t)p.ldo * 4, nrows, lane);
            __syncwarp();
        }
    } break;
    }
}

}  // namespace ttts
