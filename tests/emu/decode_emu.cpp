// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/gpt_decode.cu (kernels + launch sequence, unchanged source) on the CUDA emulation
// layer cuda_emu.h.  Built by tests/test_emu_decode_cpu.py with g++; exposes the same entry points the C ABI has, on HOST pointers.
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/gpt_decode.cu"

extern "C" {
int emu_gpt_decode_step(const ttts_gpt_decode* a) { return ttts::gpt_decode_step(a, nullptr); }
int emu_kv_fill_layer(const void* qkv, int B, int T, int d, int H, int n_pos, void* kcache, void* vcache, int T_max) {
    return ttts::gpt_kv_fill_layer((const ttts::bf16*)qkv, B, T, d, H, n_pos, (ttts::bf16*)kcache, (ttts::bf16*)vcache, T_max, nullptr);
}
long long emu_param_off(const ttts_gpt_config* c, int tensor, int layer) { return ttts::gpt_param_off(*c, tensor, layer); }
long long emu_param_count(const ttts_gpt_config* c) { return ttts::check_cfg(*c) == TTTS_OK ? ttts::make_layout(*c).total : -1; }
long long emu_kv_bytes(const ttts_gpt_config* c, int B, int T_max) { return ttts::gpt_kv_bytes(c->layers, B, c->heads, T_max); }
long long emu_decode_workspace_bytes(const ttts_gpt_config* c, int B) { return ttts::gpt_decode_workspace_bytes(B, c->model_dim); }
const char* emu_last_error() { return ttts::g_err; }
unsigned long long emu_launches() { return ttts_emu::launches; }
}
