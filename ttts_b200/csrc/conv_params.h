// Parameter block and tile constants shared by the fp32 convolution kernels (conv1d.cu, conv1d_split.cu).  Included after common.cuh /
// host_util.h, or after the CPU emulation prelude (tests/emu/cuda_emu.h).
#pragma once

namespace ttts {

constexpr int IG_P = 64, IG_R = 16;       // implicit-GEMM tile: 64 output positions x 16 reduction rows per chunk
constexpr int IG_STAGES = 4;              // cp.async ring depth of the pipelined kernels

struct ConvParams {
    const float* x; const float* w; const float* bias; float* y;
    int B, Cin, Tin, Cout, Tout, K, stride, dil, pad;
    int pre_lrelu;            // leaky_relu(0.1) on the input
    const float* resid;       // [B, Cout_eff, Tout] added to the result (may alias y)
    float out_scale;          // y = (conv + resid) * out_scale
    int accumulate;           // y += ... instead of y = ...
    const float* mask;        // [B, Tout] multiplied in (or null)
    int post;                 // 0 none, 1 GLU (Cout = 2*C: y[C] = a * sigmoid(b) (+resid)), 2 Mish, 3 WN gate (Cout = 2*C with cond)
    const float* cond;        // post==3: [B, 2*C] per-batch conditioning added before tanh/sigmoid (or null)
    int cond_ld;
};

TTTS_DEVICE float mish_f(float x) {
    const float sp = x > 20.f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}

}  // namespace ttts
