// placeholder until the tcgen05 kernels land
#include "../../include/saragan_b200.h"
#include "common.cuh"
int sg_tc_fprop(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, float, int, void*, int64_t, cudaStream_t) { return 1; }
int sg_tc_wgrad(const void*, const void*, float*, float*, int, int, int, int, int, int, float, void*, int64_t, cudaStream_t) { return 1; }
int64_t sg_tc_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
