#include "ddp_common.cuh"
extern "C" const char *ddp_version(void) { return "ddp_b200 0.1 (sm_100a)"; }
