// select.cu — candidate-list merge (the exchange step of the sharded mode and
// of bank splits inside one GPU) and key decoding.
//
// Inputs are lists of 64-bit selection keys sorted descending under the
// canonical (sim desc, idx asc) order, so merging G lists is itself a top-k of
// their union under one integer compare; the result is independent of G and
// of how the bank was partitioned (SURVEY.md §8e invariant).
#include "common.cuh"
#include "kernels.h"

namespace b200knn {
namespace {

// one warp per query row; HBM-bound: reads G*k_in*8 B, writes k_out*8 B per row
template <int ITEMS>
__global__ void __launch_bounds__(128) merge_kernel(const uint64_t* __restrict__ in, int G,
                                                    int64_t B, int k_in, int k_out,
                                                    uint64_t* __restrict__ out) {
  constexpr int CAP = ITEMS * 32;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int64_t total = int64_t(G) * k_in;
  auto fetch = [&](int64_t t) -> uint64_t {
    if (t >= total) return 0ull;
    const int64_t g = t / k_in, j = t - g * k_in;
    return in[(g * B + row) * k_in + j];
  };
  uint64_t v[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) v[r] = fetch(r * 32 + lane);
  warp_sort_desc<ITEMS>(v, lane);
  int64_t next = CAP;
  while (next < total) {
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = r * 32 + lane;
      if (i >= k_out) v[r] = fetch(next + (i - k_out));
    }
    next += CAP - k_out;
    warp_sort_desc<ITEMS>(v, lane);
  }
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = r * 32 + lane;
    if (i < k_out) out[row * k_out + i] = v[r];
  }
}

__global__ void decode_kernel(const uint64_t* __restrict__ keys, int64_t n, float* __restrict__ sims,
                              int64_t* __restrict__ idx) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = keys[i];
  if (sims) sims[i] = key_sim(key);
  if (idx) idx[i] = key_idx(key);
}

template <int ITEMS>
cudaError_t launch_merge_t(const uint64_t* in, int G, int64_t B, int k_in, int k_out, uint64_t* out,
                           cudaStream_t stream) {
  const int warps = 4;
  const int64_t blocks = (B + warps - 1) / warps;
  merge_kernel<ITEMS><<<unsigned(blocks), warps * 32, 0, stream>>>(in, G, B, k_in, k_out, out);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_merge(const uint64_t* in, int G, int64_t B, int k_in, int k_out, uint64_t* out,
                         cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  const int cap = list_capacity(k_out);
  switch (cap) {
    case 64: return launch_merge_t<2>(in, G, B, k_in, k_out, out, stream);
    case 128: return launch_merge_t<4>(in, G, B, k_in, k_out, out, stream);
    case 256: return launch_merge_t<8>(in, G, B, k_in, k_out, out, stream);
    case 512: return launch_merge_t<16>(in, G, B, k_in, k_out, out, stream);
    case 1024: return launch_merge_t<32>(in, G, B, k_in, k_out, out, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_decode(const uint64_t* keys, int64_t n, float* sims, int64_t* idx,
                          cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int threads = 256;
  decode_kernel<<<unsigned((n + threads - 1) / threads), threads, 0, stream>>>(keys, n, sims, idx);
  return cudaGetLastError();
}

}  // namespace b200knn
