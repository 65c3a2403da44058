// tcgen05 / TMEM / TMA implicit-GEMM path (placeholder until the kernel lands): -1 = "not handled".
#include "icf_common.cuh"
int icf_tc_conv_forward(const icf_conv_args*, cudaStream_t) { return -1; }
int icf_tc_conv_wgrad(const icf_wgrad_args*, cudaStream_t) { return -1; }
