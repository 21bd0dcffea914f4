// tc_inst.h — one launcher per (mode, CTA-pair) instantiation of tc_topk_impl.cuh.
#pragma once
#include "kernels.h"

namespace b200knn {
#define B200KNN_TC_LAUNCHER(name)                                                                   \
  cudaError_t name(const TcParams& p, int grid, int cap, cudaStream_t stream, float* dump, int32_t* diag, \
                   int flags, const char** why)
B200KNN_TC_LAUNCHER(launch_tc_bf16_pair);
B200KNN_TC_LAUNCHER(launch_tc_bf16_single);
B200KNN_TC_LAUNCHER(launch_tc_f16_pair);
B200KNN_TC_LAUNCHER(launch_tc_f16_single);
B200KNN_TC_LAUNCHER(launch_tc_f16x2_pair);
B200KNN_TC_LAUNCHER(launch_tc_f16x2_single);
B200KNN_TC_LAUNCHER(launch_tc_bf16x3);
B200KNN_TC_LAUNCHER(launch_tc_tf32x3);
}  // namespace b200knn
