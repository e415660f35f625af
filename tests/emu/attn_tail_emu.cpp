// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/attention_tail.cu (kernels + launch code, unchanged source) on the CUDA emulation layer
// cuda_emu.h.  Built by tests/test_emu_attn_tail_cpu.py with g++; the entry points take HOST pointers.
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/attention_tail.cu"

extern "C" {
int emu_attn_tail_rows(int T) { return ttts::attn_tail_rows(T); }
int emu_attn_tail_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int Tm, uint32_t thresh16, float drop_scale, uint64_t seed) {
    return ttts::attn_tail_fwd((const ttts::bf16*)qkv, (ttts::bf16*)out, lse, B, T, H, Tm, thresh16, drop_scale, seed, nullptr);
}
int emu_attn_tail_bwd(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv, float* dq_acc, int B, int T, int H, int Tm,
                      uint32_t thresh16, float drop_scale, uint64_t seed) {
    return ttts::attn_tail_bwd((const ttts::bf16*)qkv, (const ttts::bf16*)dout, lse, delta, (ttts::bf16*)dqkv, dq_acc, B, T, H, Tm, thresh16, drop_scale,
                               seed, nullptr);
}
const char* emu_last_error() { return ttts::g_err; }
}
