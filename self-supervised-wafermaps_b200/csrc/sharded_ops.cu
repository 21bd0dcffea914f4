// sharded_ops.cu — the small device-side steps that keep the bank-row-sharded fp32 mode free of
// host synchronisation and NCCL round trips (b200knn/sharded.py):
//   route_scatter : the query owner splits its merged approximate candidates by the shard that
//                   owns each candidate's bank row and stores them straight into that shard's
//                   inbox over NVLink peer memory (replaces route_keys + all_to_all_single);
//   compact_rows  : stable compaction of the row numbers whose status word has a masked bit set
//                   (rows a cascade level could not certify) into a fixed-capacity list + count,
//                   so the next level runs on a device-chosen subset without a host read;
//   scatter_rows  : write the sub-batch results back to those rows.
// The reference has no distributed kNN (scripts/WM811k_benchmark.py:54 `distributed = False`);
// these serve the scale-out of its single-device bank (src/ssl_wafermap/models/knn.py:80).
#include "common.cuh"
#include "kernels.h"

namespace b200knn {
namespace {

constexpr int kMaxPeers = 8;
struct PeerPtrs {
  uint64_t* p[kMaxPeers];
};

// One warp per owned query row r: keys (n, k) sorted by approximate similarity; shard g receives
// the keys whose bank row lies in [g*rows_per_shard, (g+1)*rows_per_shard), compacted to the front
// of row (row_offset + r) of its inbox (Q_pad, k).  Inboxes are zeroed by their owners before
// the barrier that precedes this kernel, so only the non-empty prefix is written: 8*k bytes per
// query leave the GPU, spread over the G peers.
__global__ void __launch_bounds__(128)
    route_scatter_kernel(const uint64_t* __restrict__ keys, int64_t n, int k, int64_t rows_per_shard, int G,
                         PeerPtrs inbox, int64_t row_offset) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * 4 + warp;
  if (r >= n) return;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint64_t* row = keys + r * k;
  int cnt[kMaxPeers];
#pragma unroll
  for (int g = 0; g < kMaxPeers; ++g) cnt[g] = 0;
  for (int j0 = 0; j0 < k; j0 += 32) {
    const uint64_t key = j0 + lane < k ? row[j0 + lane] : 0ull;
    const int owner = key != 0 ? int(key_idx(key) / rows_per_shard) : -1;
#pragma unroll
    for (int g = 0; g < kMaxPeers; ++g) {
      if (g < G) {  // warp-uniform
        const unsigned bm = __ballot_sync(kFull, owner == g);
        if (owner == g) inbox.p[g][(row_offset + r) * k + cnt[g] + __popc(bm & lt_mask)] = key;
        cnt[g] += __popc(bm);
      }
    }
  }
}

// rows_out[0 .. min(count, cap)) = ascending row numbers i with (status[i*ld] & mask) != 0,
// rows_out[min(count, cap) .. cap) = 0, *count_out = number of such rows (may exceed cap: the
// caller treats that as an overflow).  One block; n is at most a few hundred thousand.
__global__ void __launch_bounds__(1024)
    compact_rows_kernel(const int64_t* __restrict__ status, int64_t ld, int64_t n, int64_t mask,
                        int64_t* __restrict__ rows_out, int cap, int32_t* __restrict__ count_out) {
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int64_t i0 = 0; i0 < n; i0 += 1024) {
    const int64_t i = i0 + threadIdx.x;
    const bool f = i < n && (status[i * ld] & mask) != 0;
    const unsigned bm = __ballot_sync(kFull, f);
    if (lane == 0) warp_cnt[warp] = __popc(bm);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = warp_cnt[w];
      before += w < warp ? c : 0;
      total += c;
    }
    const int pos = base_s + before + __popc(bm & ((1u << lane) - 1u));
    if (f && pos < cap) rows_out[pos] = i;
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
  const int total = base_s;
  for (int j = (total < cap ? total : cap) + threadIdx.x; j < cap; j += 1024) rows_out[j] = 0;
  if (threadIdx.x == 0) *count_out = total;
}

// dst[rows[i], 0:width) = src[i, 0:width) for i < min(n, *count)
__global__ void __launch_bounds__(256)
    scatter_rows_kernel(int64_t* __restrict__ dst, int64_t dst_ld, const int64_t* __restrict__ src, int64_t src_ld,
                        const int64_t* __restrict__ rows, int n, const int32_t* __restrict__ count, int width) {
  const int m = min(n, *count);
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < int64_t(m) * width;
       t += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(t / width), c = int(t - int64_t(i) * width);
    dst[rows[i] * dst_ld + c] = src[int64_t(i) * src_ld + c];
  }
}

// dst[g][dst_offset + i] = src[i] for every peer g: the query owner publishes its rows' thresholds
struct PeerF32 {
  float* p[kMaxPeers];
};
__global__ void __launch_bounds__(256)
    broadcast_f32_kernel(const float* __restrict__ src, int64_t n, PeerF32 dst, int G, int64_t dst_offset) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
#pragma unroll
  for (int g = 0; g < kMaxPeers; ++g)
    if (g < G) dst.p[g][dst_offset + i] = v;
}

}  // namespace

cudaError_t launch_broadcast_f32(const float* src, int64_t n, float* const* dst, int G, int64_t dst_offset,
                                 cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (G < 1 || G > kMaxPeers) return cudaErrorInvalidValue;
  PeerF32 pp;
  for (int g = 0; g < kMaxPeers; ++g) pp.p[g] = g < G ? dst[g] : nullptr;
  broadcast_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(src, n, pp, G, dst_offset);
  return cudaGetLastError();
}

cudaError_t launch_route_scatter(const uint64_t* keys, int64_t n, int k, int64_t rows_per_shard, int G,
                                 uint64_t* const* inbox, int64_t row_offset, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (G < 1 || G > kMaxPeers) return cudaErrorInvalidValue;
  PeerPtrs pp;
  for (int g = 0; g < kMaxPeers; ++g) pp.p[g] = g < G ? inbox[g] : nullptr;
  route_scatter_kernel<<<unsigned((n + 3) / 4), 128, 0, stream>>>(keys, n, k, rows_per_shard, G, pp, row_offset);
  return cudaGetLastError();
}

cudaError_t launch_compact_rows(const int64_t* status, int64_t ld, int64_t n, int64_t mask, int64_t* rows_out,
                                int cap, int32_t* count_out, cudaStream_t stream) {
  compact_rows_kernel<<<1, 1024, 0, stream>>>(status, ld, n, mask, rows_out, cap, count_out);
  return cudaGetLastError();
}

cudaError_t launch_scatter_rows(int64_t* dst, int64_t dst_ld, const int64_t* src, int64_t src_ld,
                                const int64_t* rows, int n, const int32_t* count, int width,
                                cudaStream_t stream) {
  if (n == 0 || width == 0) return cudaSuccess;
  int64_t blocks = (int64_t(n) * width + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scatter_rows_kernel<<<unsigned(blocks), 256, 0, stream>>>(dst, dst_ld, src, src_ld, rows, n, count, width);
  return cudaGetLastError();
}

}  // namespace b200knn
