// tc_inst_f16x2_pair.cu — tc_topk_kernel<B200KNN_MODE_F16X2, *, *, *, PAIR=true> (see tc_topk_impl.cuh).
#include "tc_inst.h"
#include "tc_topk_impl.cuh"

namespace b200knn {
B200KNN_TC_LAUNCHER(launch_tc_f16x2_pair) {
  return launch_variant<B200KNN_MODE_F16X2, true>(p, grid, cap, stream, dump, diag, flags, why);
}
}  // namespace b200knn
