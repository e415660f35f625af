// Fused GEMM epilogues shared by the 1-CTA and 2-CTA tcgen05 kernels: one thread owns one output row and 32 consecutive
// columns (the tcgen05.ld 32x32b layout).  Split in two phases so the kernels can software-pipeline it:
//   epi_prefetch : issue the global loads this chunk needs (fp32 residual row / bf16 pre-activation row)
//   epi_apply    : bias (from shared memory or global) + activation / residual / dropout + 128-bit stores
#pragma once
#include "common.cuh"
#include "../../include/ttts_b200.h"

namespace ttts {

struct GemmParams {
    int M, N, K;
    int num_m_blocks, num_n_blocks, group_m;
    int num_k_blocks, kb_per_split, split_k;
    int epi;
    void* out; int ldo;
    const float* bias;
    const void* aux; int ldaux;
    void* aux_out; int ldaux_out;
    uint32_t drop_thresh16; float drop_scale; uint64_t drop_seed;
    int a3d, b3d;               // MN-major operand fetched with ONE 3-D TMA box per stage (extent % 64 == 0)
    int l2pf;                   // TMA epilogue: L2-prefetch the next tile's residual / pre-activation slab (TTTS_GEMM_L2PF=1; measured slower: default off)
    int quad;                   // CTA-pair kernel in clusters of 4: two pairs on neighbouring n-blocks share the A tile by TMA multicast
};

struct EpiAux { uint4 q[8]; };     // RESID: 32 fp32 (8 x float4) ; DGELU: 32 bf16 (first 4 x uint4)

TTTS_DEVICE bool epi_chunk_full(const GemmParams& p, int row, int col0) { return row < p.M && col0 + 32 <= p.N; }

TTTS_DEVICE void epi_prefetch(const GemmParams& p, const int row, const int col0, EpiAux& x) {
    if (!epi_chunk_full(p, row, col0)) return;
    if (p.epi == TTTS_EPI_RESID) {
        const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ldaux + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) x.q[j] = s[j];
    } else if (p.epi == TTTS_EPI_DGELU) {
        const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.aux) + (size_t)row * p.ldaux + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) x.q[j] = s[j];
    }
}

// sbias: 32 floats for this chunk's columns in shared memory (zeros when there is no bias), or nullptr -> read p.bias
// epi_static >= 0: the epilogue is a compile-time constant of the calling kernel (the switch below folds away)
TTTS_DEVICE void epi_apply(const GemmParams& p, const int row, const int col0, const uint32_t (&r)[32], const float* sbias, const EpiAux& x,
                           const int epi_static = -1) {
    if (row >= p.M || col0 >= p.N) return;
    const int epi = epi_static >= 0 ? epi_static : p.epi;
    const bool full = (col0 + 32 <= p.N);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (epi != TTTS_EPI_F32_ADD) {
        if (sbias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
        } else if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
        }
    }
    switch (epi) {
    case TTTS_EPI_BF16: {
        bf16* o = reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ldo + col0;
        if (full) {
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o4[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
        }
    } break;
    case TTTS_EPI_GELU: {
        // pre = bf16(acc + bias) ; h = bf16(gelu_new(pre))   (reference: bf16 autocast, HF: modeling_gpt2.py:239-240)
        bf16* o = reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ldo + col0;
        bf16* a = p.aux_out ? reinterpret_cast<bf16*>(p.aux_out) + (size_t)row * p.ldaux_out + col0 : nullptr;
        float h[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = bf16_round(v[j]); h[j] = gelu_new_fast(v[j]); }
        if (full) {
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o4[j] = make_uint4(pack_bf16(h[8 * j], h[8 * j + 1]), pack_bf16(h[8 * j + 2], h[8 * j + 3]),
                                   pack_bf16(h[8 * j + 4], h[8 * j + 5]), pack_bf16(h[8 * j + 6], h[8 * j + 7]));
            if (a) {
                uint4* a4 = reinterpret_cast<uint4*>(a);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    a4[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                       pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) { o[j] = __float2bfloat16_rn(h[j]); if (a) a[j] = __float2bfloat16_rn(v[j]); }
        }
    } break;
    case TTTS_EPI_RESID: {
        // x_out = x_in + dropout(bf16(acc + bias))     (HF: modeling_gpt2.py:224,282 / 242,307)
        float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
        if (p.drop_thresh16) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const uint64_t e4 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(col0 + 4 * j4)) >> 2;
                const uint64_t bits = dropout_bits4(p.drop_seed, e4);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[4 * j4 + j] = dropout_keep(bits, j, p.drop_thresh16) ? v[4 * j4 + j] * p.drop_scale : 0.f;
            }
        }
        if (full) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 xq = x.q[j];
                o4[j] = make_float4(__uint_as_float(xq.x) + v[4 * j], __uint_as_float(xq.y) + v[4 * j + 1],
                                    __uint_as_float(xq.z) + v[4 * j + 2], __uint_as_float(xq.w) + v[4 * j + 3]);
            }
        } else {
            const float* xin = reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ldaux + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = xin[j] + v[j];
        }
    } break;
    case TTTS_EPI_DGELU: {
        bf16* o = reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ldo + col0;
        if (full) {
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 pp = x.q[j];
                const uint32_t w[4] = {pp.x, pp.y, pp.z, pp.w};
                uint32_t ow[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float g0 = v[8 * j + 2 * t] * gelu_new_grad_fast(bf16_lo(w[t]));
                    const float g1 = v[8 * j + 2 * t + 1] * gelu_new_grad_fast(bf16_hi(w[t]));
                    ow[t] = pack_bf16(g0, g1);
                }
                o4[j] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        } else {
            const bf16* pre = reinterpret_cast<const bf16*>(p.aux) + (size_t)row * p.ldaux + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j] * gelu_new_grad_fast(__bfloat162float(pre[j])));
        }
    } break;
    case TTTS_EPI_F32_ADD: {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
        if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                             "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(o + j, v[j]);
        }
    } break;
    default: {  // TTTS_EPI_F32
        float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
        if (full) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = v[j];
        }
    } break;
    }
}

// un-pipelined convenience wrapper (1-CTA kernel)
TTTS_DEVICE void gemm_epilogue_chunk(const GemmParams& p, const int row, const int col0, const uint32_t (&r)[32]) {
    EpiAux x;
    epi_prefetch(p, row, col0, x);
    epi_apply(p, row, col0, r, nullptr, x);
}

}  // namespace ttts
