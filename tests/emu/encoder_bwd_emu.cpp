// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/encoder_bwd.cu (unchanged source; its extern "C" entry points are used directly, on host
// pointers) on the CUDA emulation layer cuda_emu.h.
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/encoder_bwd.cu"
extern "C" const char* emu_last_error() { return ttts::g_err; }
