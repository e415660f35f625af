// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/diffusion_kernels.cu (unchanged source, its extern "C" entry points on host pointers).
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/diffusion_kernels.cu"
extern "C" const char* emu_last_error() { return ttts::g_err; }
