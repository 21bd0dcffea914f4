// rescore.cu — exact re-scoring of tensor-core candidates ("fp32" mode).
//
// The tensor-core pass (tc_topk.cu) returns, per query, the k_in > k best bank
// rows under APPROXIMATE similarities (bf16 or 3xTF32 products, fp32 TMEM
// accumulation).  This kernel recomputes the similarity of every candidate
// exactly as MODE_EXACT defines it — an fmaf chain over d = 0..D-1 from +0.0f on
// the caller's fp32 values — re-ranks them under the canonical (sim desc, idx
// asc) order and keeps the best k.  The result is bitwise the MODE_EXACT result
// whenever the candidate set contains the true top-k, which is CERTIFIED per row:
//     exact_sim(rank k)  >  approx_sim(rank k_in) + E,
//     E = err_coef * ||q|| * max_n ||bank_n||      (Cauchy-Schwarz bound on the
//         approximation error of any non-candidate row)
// Rows that fail the certificate are flagged; the host shim re-runs those rows in
// MODE_EXACT, so the mode as a whole reproduces lightly's fp32 knn_predict
// neighbours (reference call site src/ssl_wafermap/models/knn.py:91-98) bit for
// bit against the oracle (oracle/seqfma.c).
//
// HBM/L2-bound gather: per query k_in rows of D fp32 (k_in*D*4 B, x2 in TF32X3
// where x = hi + lo is reassembled), 8*k B out.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/b200knn.h"
#include "common.cuh"
#include "kernels.h"

namespace b200knn {
namespace {

constexpr int kWarps = 4;
constexpr int kChunk = 64;            // d-columns staged per step
constexpr int kTileLd = kChunk + 1;   // padded: lane c walks row c conflict-free

__device__ __forceinline__ float ldq(const void* p, int dtype, int64_t i) {
  if (dtype == B200KNN_F32) return static_cast<const float*>(p)[i];
  if (dtype == B200KNN_F16) return __half2float(static_cast<const __half*>(p)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}

template <int ITEMS>
__global__ void __launch_bounds__(kWarps * 32)
    rescore_kernel(RescoreParams p) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float* qs = reinterpret_cast<float*>(rs_smem);                       // [dim_pad]
  float* tiles = qs + p.dim_pad;                                        // [kWarps][32][kTileLd]
  uint64_t* keys = reinterpret_cast<uint64_t*>(tiles + kWarps * 32 * kTileLd);  // [ITEMS*32]
  float* red = reinterpret_cast<float*>(keys + ITEMS * 32);             // [kWarps]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x;
  const uint64_t* cand = p.cand + b * p.k_in;

  float qq = 0.0f;
  for (int d = threadIdx.x; d < p.dim_pad; d += kWarps * 32) {
    const float v = d < p.dim ? ldq(p.q, p.q_dtype, b * p.q_ld + d) : 0.0f;
    qs[d] = v;
    qq = fmaf(v, v, qq);
  }
  for (int i = threadIdx.x; i < ITEMS * 32; i += kWarps * 32) keys[i] = 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);
  if (lane == 0) red[warp] = qq;
  __syncthreads();

  float* tile = tiles + warp * 32 * kTileLd;
  for (int c0 = warp * 32; c0 < p.k_in; c0 += kWarps * 32) {
    const int c = c0 + lane;
    const uint64_t key = c < p.k_in ? cand[c] : 0ull;
    const int64_t row = key != 0 ? key_idx(key) - p.idx_offset : -1;
    float acc = 0.0f;
    for (int d0 = 0; d0 < p.dim_pad; d0 += kChunk) {
      // cooperative, coalesced: lane l fetches columns d0+2l, d0+2l+1 of each of the 32 rows
      for (int cc = 0; cc < 32; ++cc) {
        const int64_t r = __shfl_sync(kFull, row, cc);
        float2 v = make_float2(0.0f, 0.0f);
        if (r >= 0) {
          v = *reinterpret_cast<const float2*>(p.rows_a + r * p.dim_pad + d0 + 2 * lane);
          if (p.rows_b != nullptr) {
            const float2 w = *reinterpret_cast<const float2*>(p.rows_b + r * p.dim_pad + d0 + 2 * lane);
            v.x += w.x;  // hi + lo is exact: lo = x - rna_tf32(x)
            v.y += w.y;
          }
        }
        tile[cc * kTileLd + 2 * lane] = v.x;
        tile[cc * kTileLd + 2 * lane + 1] = v.y;
      }
      __syncwarp();
#pragma unroll 8
      for (int dd = 0; dd < kChunk; ++dd) acc = __fmaf_rn(qs[d0 + dd], tile[lane * kTileLd + dd], acc);
      __syncwarp();
    }
    if (c < p.k_in) keys[c] = row >= 0 ? make_key(acc, uint32_t(row + p.idx_offset)) : 0ull;
  }
  __syncthreads();

  if (warp == 0) {
    uint64_t v[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) v[r] = keys[r * 32 + lane];
    warp_sort_desc<ITEMS>(v, lane);
    uint64_t kth = 0;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = r * 32 + lane;
      if (i < p.k_out) p.out[b * p.k_out + i] = v[r];
      kth = (i == p.k_out - 1) ? v[r] : kth;
    }
    kth = __shfl_sync(kFull, kth, (p.k_out - 1) & 31);
    if (lane == 0) {
      const uint64_t last = cand[p.k_in - 1];  // worst candidate under the approximate order
      int ok = 1;
      if (last == 0) {
        // an empty slot certifies the row only when every bank row was a candidate (k_in >= N);
        // otherwise the candidate pass ran under an admission threshold that starved this row
        ok = p.all_rows ? 1 : 0;
      } else {
        const float qn = sqrtf(red[0] + red[1] + red[2] + red[3]) * 1.001f;
        const float e = p.err_coef * qn * (*p.bank_max_norm);
        ok = (kth != 0) && (key_sim(kth) > key_sim(last) + e);
      }
      p.uncertified[b] = ok ? 0 : 1;
      if (!ok) atomicAdd(p.n_uncertified, 1);
    }
  }
}

// max over rows of ||row||_2 (fp32), inflated by 0.1 % for the rounding of the reduction
__global__ void __launch_bounds__(256)
    row_norm_max_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                        int dim_pad, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  float best = 0.0f;
  for (int64_t r = warp; r < n; r += n_warps) {
    float s = 0.0f;
    for (int d = lane; d < dim_pad; d += 32) {
      float v = a[r * dim_pad + d];
      if (b != nullptr) v += b[r * dim_pad + d];
      s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    best = fmaxf(best, s);
  }
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best) * 1.001f));
}

template <int ITEMS>
cudaError_t launch_rescore_t(const RescoreParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (size_t(p.dim_pad) + kWarps * 32 * kTileLd) +
                      sizeof(uint64_t) * ITEMS * 32 + sizeof(float) * kWarps;
  cudaError_t e = cudaFuncSetAttribute(rescore_kernel<ITEMS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  rescore_kernel<ITEMS><<<unsigned(p.B), kWarps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_rescore(const RescoreParams& p, cudaStream_t stream) {
  if (p.B == 0) return cudaSuccess;
  const int items = (p.k_in + 31) / 32;
  if (items <= 2) return launch_rescore_t<2>(p, stream);
  if (items <= 4) return launch_rescore_t<4>(p, stream);
  if (items <= 8) return launch_rescore_t<8>(p, stream);
  if (items <= 16) return launch_rescore_t<16>(p, stream);
  if (items <= 32) return launch_rescore_t<32>(p, stream);
  return cudaErrorInvalidValue;
}

cudaError_t launch_row_norm_max(const float* a, const float* b, int64_t n, int dim_pad, float* out,
                                cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), stream);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  const int blocks = int(n < 148 * 8 * 8 ? (n + 7) / 8 : 148 * 8);
  row_norm_max_kernel<<<blocks, 256, 0, stream>>>(a, b, n, dim_pad, out);
  return cudaGetLastError();
}

}  // namespace b200knn
