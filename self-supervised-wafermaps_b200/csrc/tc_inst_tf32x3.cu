// tc_inst_tf32x3.cu — tc_topk_kernel<B200KNN_MODE_TF32X3, *, *, *, PAIR=false> (see tc_topk_impl.cuh).
#include "tc_inst.h"
#include "tc_topk_impl.cuh"

namespace b200knn {
B200KNN_TC_LAUNCHER(launch_tc_tf32x3) {
  return launch_variant<B200KNN_MODE_TF32X3, false>(p, grid, cap, stream, dump, diag, flags, why);
}
}  // namespace b200knn
