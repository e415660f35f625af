// Shared device helpers for the ttts_b200 sm_100a kernels: mbarrier / TMA / tcgen05 PTX
// wrappers, warp reductions, bf16 packing, counter-based dropout hash.
// Everything here is inline PTX for sm_100a (no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef TTTS_DEVICE
#define TTTS_DEVICE __device__ __forceinline__
#endif

namespace ttts {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------------------
TTTS_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

TTTS_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
TTTS_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

TTTS_DEVICE uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
TTTS_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
TTTS_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
TTTS_DEVICE float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// gelu_new (tanh approximation), HF activations.py:59-66:
//   0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715 x^3)))
TTTS_DEVICE float gelu_new_f(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (x + k1 * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
}
TTTS_DEVICE float gelu_new_grad_f(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (x + k1 * x * x * x);
    float t = tanhf(u);
    float du = k0 * (1.0f + 3.0f * k1 * x * x);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

// Same functions on the SFU (tanh.approx.f32, rel. error ~2^-11): used in the GEMM epilogues, whose results are rounded to
// bf16 (rel. 2^-9) anyway; the exact versions above stay for fp32 consumers.
TTTS_DEVICE float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
TTTS_DEVICE float gelu_new_fast(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float u = k0 * x * fmaf(k1 * x, x, 1.0f);
    return 0.5f * x * (1.0f + tanh_fast(u));
}
TTTS_DEVICE float gelu_new_grad_fast(float x) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float x2 = x * x;
    const float t = tanh_fast(k0 * x * fmaf(k1, x2, 1.0f));
    const float du = k0 * fmaf(3.0f * k1, x2, 1.0f);
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

// Packed bf16x2 forms for the GEMM epilogues (two elements per instruction, tanh on the SFU at two per MUFU op).  The reference runs
// gelu_new as a chain of bf16 elementwise kernels under autocast (every intermediate rounded to bf16, HF: activations.py:59-66), so bf16
// intermediates are its own precision; constants are the bf16 roundings of k0 = sqrt(2/pi), k0*0.044715, 3*k0*0.044715.
TTTS_DEVICE uint32_t bf2_mul(uint32_t a, uint32_t b) { uint32_t d; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
TTTS_DEVICE uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
TTTS_DEVICE uint32_t bf2_tanh(uint32_t a) { uint32_t d; asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a)); return d; }
constexpr uint32_t kBf2K0 = 0x3F4C3F4Cu, kBf2K0K1 = 0x3D123D12u, kBf2K0K1x3 = 0x3DDB3DDBu, kBf2Half = 0x3F003F00u, kBf2One = 0x3F803F80u;
// gelu_new of two packed bf16 values
TTTS_DEVICE uint32_t gelu_new_bf2(uint32_t x) {
    const uint32_t x2 = bf2_mul(x, x);
    const uint32_t u = bf2_mul(x, bf2_fma(x2, kBf2K0K1, kBf2K0));      // k0 x (1 + k1 x^2)
    const uint32_t t = bf2_tanh(u);
    const uint32_t hx = bf2_mul(x, kBf2Half);
    return bf2_fma(hx, t, hx);                                          // 0.5 x (1 + t)
}
// d gelu_new / dx of two packed bf16 values
TTTS_DEVICE uint32_t gelu_new_grad_bf2(uint32_t x) {
    const uint32_t x2 = bf2_mul(x, x);
    const uint32_t u = bf2_mul(x, bf2_fma(x2, kBf2K0K1, kBf2K0));
    const uint32_t t = bf2_tanh(u);
    const uint32_t du = bf2_fma(x2, kBf2K0K1x3, kBf2K0);              // k0 (1 + 3 k1 x^2)
    const uint32_t a = bf2_fma(t, kBf2Half, kBf2Half);                  // 0.5 (1 + t)
    const uint32_t omt2 = bf2_fma(t ^ 0x80008000u, t, kBf2One);         // 1 - t^2
    const uint32_t b = bf2_mul(bf2_mul(x, kBf2Half), omt2);             // 0.5 x (1 - t^2)
    return bf2_fma(b, du, a);
}

}  // namespace ttts
#include "attn_drop_hash.h"      // mix64, dropout_keep, attn_drop_* (also compiled by the host emulation, tests/emu)
namespace ttts {

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
TTTS_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TTTS_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TTTS_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

TTTS_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TTTS_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TTTS_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a kernel bug must not hang the GPU box (that is a "strike"); after 2 s of wall clock (%globaltimer)
// we trap so the launch fails loudly instead.
TTTS_DEVICE uint64_t global_timer_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
TTTS_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {          // try_wait itself suspends the thread for a HW-defined window
        if ((++spins & 0x3ffu) == 0) {
            const uint64_t t = global_timer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) { asm volatile("trap;"); }
        }
    }
}

// the same wait for a warp that is far ahead of its consumers (TMA producers waiting for a free stage): sleeps between polls so that its
// spin loop does not take issue slots from the math warps of its scheduler (r1n attention profile: 15 % of all stall samples sat on the
// branches of such loops)
TTTS_DEVICE void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
#ifndef TTTS_HOST_EMU
        __nanosleep(100);
#endif
        if ((++spins & 0x3ffu) == 0) {
            const uint64_t t = global_timer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) { asm volatile("trap;"); }
        }
    }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) 2D load into shared memory, completion on an mbarrier
// ----------------------------------------------------------------------------------------
TTTS_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
TTTS_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

TTTS_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
TTTS_DEVICE void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
TTTS_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TTTS_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TTTS_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One elected lane of a fully converged warp.  The single-thread instructions (tcgen05.mma / commit, TMA) take their operands from
// UNIFORM registers: issued from inside `if (lane == 0)` the compiler must move every operand there with ELECT + R2UR.BROADCAST inside
// a BRA.U.ANY retry loop (~21 SASS instructions and ~200 cycles per MMA: ncu showed that loop, not the tensor pipe, bounding the GEMM and
// the attention kernels).  With the whole warp running the role loop and only the instruction itself predicated on elect_one(), the
// descriptors are computed on the uniform datapath and an MMA costs a handful of instructions.
TTTS_DEVICE bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, single CTA.
TTTS_DEVICE void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
TTTS_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ----------------------------------------------------------------------------------------
// thread-block clusters / CTA pairs (cta_group::2)
// ----------------------------------------------------------------------------------------
TTTS_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TTTS_DEVICE void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
TTTS_DEVICE void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
TTTS_DEVICE void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
// In a CTA pair the shared::cluster address of the same offset in the EVEN (leader) CTA is the local address with bit 24
// cleared (the convention CUTLASS' Sm100MmaPeerBitMask encodes).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
// arrive on the barrier at the same offset in the leader CTA (valid from either CTA of the pair)
TTTS_DEVICE void mbar_arrive_leader(uint64_t* bar) {
    // default semantics (.release at CTA scope): the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of
    // every arrive (ncu: the peer's producer spent its time there and the CTA-pair GEMM ran at half speed, profiles/r1_notes.md)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's smem, completion bytes are signalled on the LEADER CTA's barrier
TTTS_DEVICE void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// Same, multicast: the box lands at the same smem offset in every CTA of `mask` (cluster ranks), and each destination's bytes are
// signalled on the barrier at this offset in the LEADER of that destination's CTA pair.
TTTS_DEVICE void tma_load_2d_2sm_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// arrive on the barrier at this offset in cluster rank `target`
TTTS_DEVICE void mbar_arrive_rank(uint64_t* bar, uint32_t target) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(target));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// commit with an explicit multicast mask (cluster ranks whose barrier at this offset gets one arrival)
TTTS_DEVICE void umma_commit_2sm_mask(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
TTTS_DEVICE void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {  // one whole warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
TTTS_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA); issued by the leader CTA only
TTTS_DEVICE void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit: arrive (once) on the barrier at this offset in BOTH CTAs of the pair when the issued MMAs have completed
TTTS_DEVICE void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
TTTS_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
TTTS_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16-column forms (one TMEM lane per thread, 16 consecutive fp32 columns) + the matching store
TTTS_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
TTTS_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
TTTS_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format, version=1, SWIZZLE_128B).
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1   bits [61,64) layout (2 = SW128)
TTTS_DEVICE uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// cp.async (Ampere-style) for the non-TMA kernels
// ----------------------------------------------------------------------------------------
TTTS_DEVICE void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    uint32_t sz = pred ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
TTTS_DEVICE void cp_async4(void* smem_dst, const void* gsrc, bool pred) {      // 4 bytes, zero fill when !pred (gsrc must still be a valid address)
    uint32_t sz = pred ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while
// its predecessor in the stream is still running; pdl_wait() blocks until that predecessor has completed and its writes are visible, so
// everything before it (barrier init, TMEM allocation, tensor-map prefetch -- nothing that touches global memory the predecessor writes
// or reads) overlaps the predecessor's tail.  pdl_launch_dependents() lets the successor be scheduled as soon as every CTA of THIS grid
// has executed it or exited.  Both are no-ops for kernels launched without the attribute.
TTTS_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
TTTS_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
TTTS_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
TTTS_DEVICE void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------------
// mma.sync m16n8k16 bf16 + ldmatrix (attention kernels)
// ----------------------------------------------------------------------------------------
TTTS_DEVICE void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr));
}
TTTS_DEVICE void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr));
}
TTTS_DEVICE void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

}  // namespace ttts
