// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/conv1d_grouped.cu (unchanged source, its extern "C" entry points on host pointers).
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/conv1d_grouped.cu"
extern "C" const char* emu_last_error() { return ttts::g_err; }
