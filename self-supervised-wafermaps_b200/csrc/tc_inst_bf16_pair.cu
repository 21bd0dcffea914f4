// tc_inst_bf16_pair.cu — tc_topk_kernel<B200KNN_MODE_BF16, *, *, *, PAIR=true> (see tc_topk_impl.cuh).
#include "tc_inst.h"
#include "tc_topk_impl.cuh"

namespace b200knn {
B200KNN_TC_LAUNCHER(launch_tc_bf16_pair) {
  return launch_variant<B200KNN_MODE_BF16, true>(p, grid, cap, stream, dump, diag, flags, why);
}
}  // namespace b200knn
