// Counter-based dropout hashes of the GPT train step (element dropout and attention-probability dropout), shared by the sm_100a kernels
// (through common.cuh) and by the host emulation of the plain kernels (tests/emu: included after cuda_emu.h with TTTS_HOST_EMU defined).
// Needs TTTS_DEVICE and <stdint.h>; everything lives in namespace ttts.
#pragma once

namespace ttts {

// Counter-based dropout generator: one 64-bit mix -> 4 keep-decisions of 16 bits each.
// keep iff u16 >= thresh16 where thresh16 = round(p*65536).
TTTS_DEVICE uint64_t mix64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdULL;
    z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL;
    z ^= z >> 33;
    return z;
}
// element index e (global, per site) -> keep? ; 4 consecutive elements share one mix.
TTTS_DEVICE uint64_t dropout_bits4(uint64_t seed, uint64_t e4) { return mix64(seed + e4 * 0x9E3779B97F4A7C15ULL); }
TTTS_DEVICE bool dropout_keep(uint64_t bits, int j, uint32_t thresh16) {
    return ((uint32_t)(bits >> (16 * j)) & 0xffffu) >= thresh16;
}

// Attention-probability dropout (B*H*T*T decisions per layer: the hash must cost ~2 instructions per element, not ~5 like mix64).
// Row key = mix64(seed, b*h*T + query) once per row; then per group of 4 consecutive keys three rounds of a 32x32->64 multiply-fold
// (one IMAD.WIDE + one LOP3 each) give two 32-bit words = four 15-bit uniform fields (bits 0-14 and 16-30 of each word; bits 15 / 31
// are cleared by the same LOP3 that folds the last round):
//     key 4g + 0 -> w0 bits 0-14     key 4g + 1 -> w0 bits 16-30     key 4g + 2 -> w1 bits 0-14     key 4g + 3 -> w1 bits 16-30
//     keep <=> field >= t15,  t15 = thresh16 >> 1  (the host rounds p to a multiple of 2^-15, thresh16 is even)
// Round 2 layout (r2c): the 15-bit fields exist so that BOTH decisions of a word come out of one add -- field + (0x8000 - t15) carries
// into bit 15 / 31 exactly when the key is kept, no carry crosses a field -- and one PRMT with sign replication turns the two flag bits
// into a 0xFFFF / 0x0000 mask per half-word, which is ANDed onto the packed bf16x2 probabilities: 3 ALU instructions per 2 keys instead
// of 2 x (shift, ISETP, SEL) on fp32 values.  The r1 profiles had the attention math warps bound by the ALU pipe (half rate on sm_100).
// mul.wide.u32 in PTX: written as (uint64_t)a * b the compiler adds a dead `+ 0` to every high word (one VIADD per product).
// Statistical checks (keep rate, key/row/diagonal correlations, 2-D spectrum) are in tests/test_oracle_golden.py::test_attn_dropout_hash.
struct AttnDropRow { uint32_t k0, k1; };
TTTS_DEVICE AttnDropRow attn_drop_row(uint64_t seed, uint64_t row) {
    const uint64_t z = mix64(seed + row * 0x9E3779B97F4A7C15ULL);
    AttnDropRow k; k.k0 = (uint32_t)z; k.k1 = (uint32_t)(z >> 32);
    return k;
}
TTTS_DEVICE void mul_wide_u32(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
#ifdef TTTS_HOST_EMU
    const uint64_t m = (uint64_t)a * b; lo = (uint32_t)m; hi = (uint32_t)(m >> 32);
#else
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#endif
}
TTTS_DEVICE void attn_drop_words(const AttnDropRow k, uint32_t g, uint32_t& w0, uint32_t& w1) {
    const uint32_t a = g * 0x9E3779B1u + k.k0;
    uint32_t l1, h1, l2, h2, l3, h3;
    mul_wide_u32(a, 0x85EBCA6Bu, l1, h1);
    const uint32_t x = l1 ^ h1 ^ k.k1;
    mul_wide_u32(x, 0xC2B2AE35u, l2, h2);
    const uint32_t y = l2 ^ h2;
    mul_wide_u32(y, 0x27D4EB2Fu, l3, h3);
    w0 = (h3 ^ l2) & 0x7FFF7FFFu;
    w1 = (l3 ^ h2) & 0x7FFF7FFFu;
}
// add constant for attn_drop_mask2: both 15-bit fields of a word + (0x8000 - t15) set bit 15 / 31 <=> kept
TTTS_DEVICE uint32_t attn_drop_addc(uint32_t thresh16) { return (0x8000u - (thresh16 >> 1)) * 0x00010001u; }
// 0xFFFF in the low / high half-word where the word's low / high key is kept: AND it onto pack_bf16(p[2j], p[2j + 1])
TTTS_DEVICE uint32_t attn_drop_mask2(uint32_t w, uint32_t addc) {
#ifdef TTTS_HOST_EMU
    const uint32_t z = w + addc;
    return ((z & 0x8000u) ? 0xFFFFu : 0u) | ((z & 0x80000000u) ? 0xFFFF0000u : 0u);
#else
    uint32_t m;
    asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(m) : "r"(w + addc), "r"(0u));     // bytes 0,1 <- sign of byte 1 ; bytes 2,3 <- sign of byte 3
    return m;
#endif
}
// keep decision of key 4g + j (j = 0..3) from the group's words; t32 = thresh16 << 16 (generic form: legacy kernels, mask dump)
TTTS_DEVICE bool attn_drop_keep(uint32_t w0, uint32_t w1, int j, uint32_t t32) {
    const uint32_t w = (j & 2) ? w1 : w0;
    return (((j & 1) ? (w >> 16) : w) & 0x7FFFu) >= (t32 >> 17);
}
// slow generic form (legacy kernels, mask dump): one element
TTTS_DEVICE bool attn_drop_keep1(uint64_t seed, uint64_t row, int kj, uint32_t thresh16) {
    uint32_t w0, w1;
    attn_drop_words(attn_drop_row(seed, row), (uint32_t)kj >> 2, w0, w1);
    return attn_drop_keep(w0, w1, kj & 3, thresh16 << 16);
}

}  // namespace ttts
