// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/conv1d_bwd.cu (unchanged source) on the CUDA emulation layer cuda_emu.h.
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/conv1d_bwd.cu"

extern "C" {
int emu_conv1d_bwd_input(const float* dy, const float* w, const float* x, float* dx, int B, int Cin, int Tin, int Cout, int K, int stride, int dil,
                         int pad, int pre_lrelu, int accumulate) {
    return ttts::conv1d_bwd_input(dy, w, x, dx, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, accumulate, nullptr);
}
int emu_conv1d_bwd_weight(const float* dy, const float* x, float* dw, float* db, int B, int Cin, int Tin, int Cout, int K, int stride, int dil, int pad,
                          int pre_lrelu) {
    return ttts::conv1d_bwd_weight(dy, x, dw, db, B, Cin, Tin, Cout, K, stride, dil, pad, pre_lrelu, nullptr);
}
int emu_bias_grad(const float* dy, float* db, int B, int C, int T) {
    return ttts::launch_plain(ttts::conv1d_bgrad_kernel, dim3(C), dim3(256), 0, nullptr, dy, db, B, C, T);
}
const char* emu_last_error() { return ttts::g_err; }
}
