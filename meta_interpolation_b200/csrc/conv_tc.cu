// placeholder, replaced below
#include "mi_common.cuh"
bool mi_tc_fprop_eligible(const float*, int, const float*, int, const float*, int, int, int, int, int, int, int) { return false; }
bool mi_tc_wgrad_eligible(const float*, int, const float*, int, int, int, int, int, int, int) { return false; }
int mi_tc_fprop(const float*, int, const float*, int, const float*, float*, int, const float*, int, int, float, int, int, int, int, int, int, int, int, float, cudaStream_t) { return MI_ERR_UNSUPPORTED; }
int mi_tc_wgrad_partials(const float*, int, const float*, int, int, int, int, int, int, int, int, float*, float*, int, cudaStream_t) { return MI_ERR_UNSUPPORTED; }
extern "C" int mi_tc_available(void) { return 0; }
