// tc_topk.cu — dispatch of the tcgen05 similarity + top-k kernel (tc_topk_impl.cuh) over its
// (mode, CTA-pair) instantiations, each compiled in its own translation unit (tc_inst_*.cu).
#include <cstdlib>

#include "../../include/b200knn.h"
#include "kernels.h"
#include "tc_inst.h"

namespace b200knn {

// Bank rows per accumulator tile.  The resident-query modes can run 256-wide tiles (2 TMEM
// buffers) up to D_pad = 512; B200KNN_TILE_N=128 forces 128-wide tiles (4 TMEM buffers) for A/B runs.
int tc_tile_n(int mode, int dim) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("B200KNN_TILE_N");
    forced = (e != nullptr && atoi(e) == 128) ? 128 : 0;
  }
  if (forced == 128) return 128;
  if (mode == B200KNN_MODE_BF16 || mode == B200KNN_MODE_F16X2 || mode == B200KNN_MODE_F16)
    return ((dim + 63) / 64 * 64) <= 512 ? 256 : 128;
  return 128;
}

// The resident-query modes (BF16, F16, F16X2) run as CTA pairs (cta_group::2, 256 query rows per
// work item) unless the whole batch fits one 128-row tile: a pair would then spend half of its
// MMA cycles on padding rows, and at B <= 128 the call is bound by streaming the bank, which
// 148 single CTAs (M = 128) do at twice the rows per MMA cycle of 74 pairs (M = 256).
// B200KNN_NO_PAIR=1 forces single CTAs (A/B experiments).
bool tc_use_pair(int mode, int64_t B) {
  static int no_pair = -1;
  if (no_pair < 0) {
    const char* e = getenv("B200KNN_NO_PAIR");
    no_pair = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (mode != B200KNN_MODE_BF16 && mode != B200KNN_MODE_F16X2 && mode != B200KNN_MODE_F16) return false;
  return no_pair == 0 && B > 128;
}

cudaError_t launch_tc(const TcParams& p, int grid, int cap, cudaStream_t stream, float* dump,
                      int32_t* diag, int flags, const char** why) {
  *why = "";
  const bool pair = tc_use_pair(p.mode, p.B);
  switch (p.mode) {
    case B200KNN_MODE_BF16:
      return pair ? launch_tc_bf16_pair(p, grid, cap, stream, dump, diag, flags, why)
                  : launch_tc_bf16_single(p, grid, cap, stream, dump, diag, flags, why);
    case B200KNN_MODE_F16:
      return pair ? launch_tc_f16_pair(p, grid, cap, stream, dump, diag, flags, why)
                  : launch_tc_f16_single(p, grid, cap, stream, dump, diag, flags, why);
    case B200KNN_MODE_F16X2:
      return pair ? launch_tc_f16x2_pair(p, grid, cap, stream, dump, diag, flags, why)
                  : launch_tc_f16x2_single(p, grid, cap, stream, dump, diag, flags, why);
    case B200KNN_MODE_BF16X3: return launch_tc_bf16x3(p, grid, cap, stream, dump, diag, flags, why);
    case B200KNN_MODE_TF32X3: return launch_tc_tf32x3(p, grid, cap, stream, dump, diag, flags, why);
    default: *why = "unknown mode"; return cudaErrorNotSupported;
  }
}

}  // namespace b200knn
