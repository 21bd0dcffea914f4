// plan.h — host-side work decomposition shared by the similarity kernels.
//
// A work item is (query tile of TILE_M rows) x (bank split of split_rows rows).
// Items are ordered split-major so CTAs that run concurrently read the same
// bank rows (one HBM read per wave, the rest L2 hits).  Each resident CTA owns
// TILE_M candidate lists in the caller's workspace; per-split results go to a
// (splits, B, k) partial buffer and are merged by merge_kernel when splits > 1.
#pragma once
#include <cstddef>
#include <cstdint>

namespace b200knn {

struct TopkPlan {
  int tile_m = 128;
  int tile_n = 128;
  int64_t n_qtiles = 0;
  int splits = 1;
  int64_t split_rows = 0;  // multiple of tile_n
  int64_t n_items = 0;
  int grid = 0;
  int cap = 0;            // list capacity (keys per row)
  // chunk-major order (tensor-core kernel, one split, several tiles per worker): a worker keeps
  // `slots` tiles open and scans the bank in `chunks` pieces of `chunk_rows` rows (a multiple of
  // tile_n) that stay in L2 while every open tile of every worker passes over them
  int chunks = 1;
  int slots = 1;
  int64_t chunk_rows = 0;
  size_t lists_bytes = 0;   // grid * slots * tile_m * cap * 8
  size_t state_bytes = 0;   // chunks>1 ? B * 8 (parked threshold + list fill per row) : 0
  size_t partial_bytes = 0; // splits>1 ? splits * B * k * 8 : 0
  size_t total_bytes = 0;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ctas: number of concurrently resident CTAs the kernel runs (SM count x CTAs/SM).
// Chunking parameters: bytes one bank row occupies in the streamed operand arrays, and the bytes
// of bank that should stay L2-resident during a chunk phase (0: never chunk).
constexpr int kMaxSlots = 16;
inline TopkPlan make_plan(int64_t B, int64_t N, int k, int cap, int tile_m, int tile_n, int ctas,
                          int64_t bank_row_bytes = 0, int64_t l2_chunk_bytes = 0) {
  TopkPlan p;
  p.tile_m = tile_m;
  p.tile_n = tile_n;
  p.cap = cap;
  p.n_qtiles = (B + tile_m - 1) / tile_m;
  // cost(S) ~ waves(S) * (rows per split + warm-up), warm-up ~ the rows it takes
  // a fresh list to reach a selective threshold (a few multiples of k).
  const double warm = 40.0 * k;
  const int64_t n_tiles = (N + tile_n - 1) / tile_n;
  double best = 1e300;
  int best_s = 1;
  for (int s = 1; s <= 1024 && s <= n_tiles; ++s) {
    const int64_t tiles_per = (n_tiles + s - 1) / s;
    const int64_t s_eff = (n_tiles + tiles_per - 1) / tiles_per;
    if (s_eff != s) continue;
    const int64_t items = p.n_qtiles * s;
    const int64_t waves = (items + ctas - 1) / ctas;
    const double cost = double(waves) * (double(tiles_per) * tile_n + warm);
    if (cost < best * 0.98) {  // prefer fewer splits unless clearly better
      best = cost;
      best_s = s;
    }
  }
  const int64_t tiles_per = (n_tiles + best_s - 1) / best_s;
  p.splits = int((n_tiles + tiles_per - 1) / tiles_per);
  p.split_rows = tiles_per * tile_n;
  p.n_items = p.n_qtiles * p.splits;
  p.grid = int(p.n_items < ctas ? p.n_items : ctas);
  if (p.grid < 1) p.grid = 1;
  p.chunk_rows = p.split_rows;
  // Chunk when one split is scanned by workers that each own several tiles and the bank does not
  // fit the L2 budget: otherwise every (worker, tile) pass re-streams the bank and the workers
  // drift apart until L2 no longer shares their reads (measured: 56 GB of DRAM reads for a
  // 0.83 GB bank at 8 tiles per worker, profiles/r02_call5_ncu_full_tc_topk_f16_pair_q151k.txt).
  if (p.splits == 1 && l2_chunk_bytes > 0 && bank_row_bytes > 0 && p.n_qtiles >= 2 * int64_t(p.grid) &&
      N * bank_row_bytes > 2 * l2_chunk_bytes) {
    int64_t rows = l2_chunk_bytes / bank_row_bytes / tile_n * tile_n;
    if (rows < 4 * tile_n) rows = 4 * tile_n;
    const int64_t n_chunks = (N + rows - 1) / rows;
    // even chunks (multiples of tile_n)
    rows = ((N + n_chunks - 1) / n_chunks + tile_n - 1) / tile_n * tile_n;
    p.chunks = int((N + rows - 1) / rows);
    p.chunk_rows = rows;
    const int64_t per_worker = (p.n_qtiles + p.grid - 1) / p.grid;
    p.slots = int(per_worker < kMaxSlots ? per_worker : kMaxSlots);
    if (p.chunks <= 1) {
      p.chunks = 1;
      p.slots = 1;
      p.chunk_rows = p.split_rows;
    }
  }
  p.lists_bytes = align_up(size_t(p.grid) * p.slots * tile_m * cap * 8, 256);
  p.state_bytes = p.chunks > 1 ? align_up(size_t(B) * 8, 256) : 0;
  p.partial_bytes = p.splits > 1 ? align_up(size_t(p.splits) * B * k * 8, 256) : 0;
  p.total_bytes = p.lists_bytes + p.state_bytes + p.partial_bytes;
  return p;
}

}  // namespace b200knn
