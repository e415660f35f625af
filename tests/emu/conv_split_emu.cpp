// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/conv1d_split.cu (unchanged source) on the CUDA emulation layer cuda_emu.h.
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/conv1d_split.cu"

extern "C" int emu_conv1d_split(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Tin, int Cout, int K, int stride,
                                int dil, int pad, int pre_lrelu, const float* resid, float out_scale, int accumulate, const float* mask, int post,
                                const float* cond, int cond_ld, int groups) {
    using namespace ttts;
    const int Tout = (Tin + 2 * pad - dil * (K - 1) - 1) / stride + 1;
    const bool gated = (post == 1 || post == 3);
    ConvParams p;
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.B = B; p.Cin = Cin; p.Tin = Tin; p.Cout = Cout; p.Tout = Tout; p.K = K; p.stride = stride;
    p.dil = dil; p.pad = pad; p.pre_lrelu = pre_lrelu; p.resid = resid; p.out_scale = out_scale; p.accumulate = accumulate; p.mask = mask;
    p.post = post; p.cond = cond; p.cond_ld = cond_ld;
    const long long Ptot = (long long)B * Tout;
    const int ceff = gated ? Cout / 2 : Cout;
    dim3 grid((unsigned)((Ptot + IG_P - 1) / IG_P), (ceff + (gated ? 15 : 31)) / (gated ? 16 : 32));     // as ttts_conv1d_f32 does (conv1d.cu)
    return conv1d_split_try(p, grid, groups, nullptr);
}
