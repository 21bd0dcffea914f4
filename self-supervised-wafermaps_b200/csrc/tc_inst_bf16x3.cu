// tc_inst_bf16x3.cu — tc_topk_kernel<B200KNN_MODE_BF16X3, *, *, *, PAIR=false> (see tc_topk_impl.cuh).
#include "tc_inst.h"
#include "tc_topk_impl.cuh"

namespace b200knn {
B200KNN_TC_LAUNCHER(launch_tc_bf16x3) {
  return launch_variant<B200KNN_MODE_BF16X3, false>(p, grid, cap, stream, dump, diag, flags, why);
}
}  // namespace b200knn
