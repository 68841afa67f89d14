// ABI version of libmlb200 (bumped on any signature change in include/mlb200.h).
#include "common.cuh"
MLB_API int mlb_abi_version(void) { return 11; }
