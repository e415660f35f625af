// TEST INFRASTRUCTURE: host build of ttts_b200/csrc/gan_losses.cu (unchanged source, its extern "C" entry points on host pointers).
#define TTTS_HOST_EMU 1
#include "../../ttts_b200/csrc/gan_losses.cu"
