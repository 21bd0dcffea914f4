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
//
// Two implementations of the same arithmetic:
//  * rescore_dot_kernel + rescore_select_kernel (fp32 queries, one fp32 row array, caller
//    workspace): persistent warps, each owning a ring of shared-memory stages that the
//    TMA engine fills with 32 candidate rows x CHUNK columns (one cp.async.bulk per row, the
//    query chunk by cp.async, all completing on the stage's mbarrier).  Lane c then walks
//    row c with conflict-free LDS.128 (row pitch CHUNK+4 floats) — the fma chain of one
//    candidate is sequential by definition, so the parallelism is 32 candidates per warp
//    and the copies of the next stage are in flight while the chain runs.  The exact keys go
//    to the workspace; a second kernel sorts them per query and evaluates the certificate.
//  * rescore_kernel (any query dtype, optional lo array, no workspace): one block per
//    query, rows staged by plain loads.  Kept for the TF32X3 operands and small calls.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "../../include/b200knn.h"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace b200knn {
namespace {

// Certificate of one row (see the header): kth = exact key at rank k_out, last = approximate key
// at rank k_in, qn = ||q|| (inflated by the caller for its rounding).
__device__ __forceinline__ int certified(const RescoreParams& p, uint64_t kth, uint64_t last, float qn) {
  if (last == 0) {
    // an empty slot certifies the row only when every bank row was a candidate (k_in >= N);
    // otherwise the candidate pass ran under an admission threshold that starved this row
    return p.all_rows ? 1 : 0;
  }
  const float m = *p.bank_max_norm;
  if (p.max_abs > 0.0f && !(qn < p.max_abs && m < p.max_abs)) return 0;  // operand range of the candidate pass
  const float e = p.err_coef * qn * m + p.err_abs * (qn + m);
  return (kth != 0) && (key_sim(kth) > key_sim(last) + e);
}

// slots per query in the workspace of the pipelined variant: k_in rounded up to 64 (work units of
// up to 64 candidate slots)
inline int ws_slots(int k_in) { return (k_in + 63) / 64 * 64; }

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
      const float qn = sqrtf(red[0] + red[1] + red[2] + red[3]) * 1.001f;
      const int ok = certified(p, kth, last, qn);
      p.uncertified[b] = ok ? 0 : 1;
      if (!ok) atomicAdd(p.n_uncertified, 1);
    }
  }
}


// ------------------------------------------------------------------ pipelined variant
// One work unit = 32*R consecutive candidate slots of one query (lane l owns slots l, l+32, ...);
// a step = one unit x one CHUNK of columns.  Each warp owns STAGES stages and a contiguous range
// of units.  R = 2 (two independent fma chains per lane) is an experiment that did not pay, see
// launch_rescore.
template <int CHUNK, int R>
struct DotStage {
  static constexpr int kPitch = CHUNK + 4;  // floats; lane c reads float4 j of row c: bank group (c + j) mod 8
  float rows[32 * R * kPitch];
  float q[CHUNK];
  uint32_t klo[32 * R];  // low key word (0xFFFFFFFF - idx) of each candidate, 0 = empty slot
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// tmp: (B, n_groups * 32 * R) exact keys, slot c of query b at tmp[b * n_groups*32*R + c]
template <int CHUNK, int STAGES, int R>
__global__ void __launch_bounds__(384)
    rescore_dot_kernel(RescoreParams p, uint64_t* __restrict__ tmp, int n_groups, int64_t n_units,
                       int q_vec16) {
  using Stage = DotStage<CHUNK, R>;
  constexpr int kSlots = 32 * R;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int n_warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Stage* stages = reinterpret_cast<Stage*>(rs_smem) + size_t(warp) * STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs_smem + sizeof(Stage) * size_t(n_warps) * STAGES) +
                   warp * STAGES;
  if (lane == 0) {
    // per phase: lane 0's arrive.expect_tx + one cp.async "noinc" arrival per lane (query chunk)
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(ptx::smem_u32(&bars[s]), 33);
    ptx::fence_barrier_init();
  }
  __syncwarp();

  const int64_t wg = int64_t(blockIdx.x) * n_warps + warp, tw = int64_t(gridDim.x) * n_warps;
  const int64_t u_begin = n_units * wg / tw, u_end = n_units * (wg + 1) / tw;
  if (u_begin >= u_end) return;
  const int n_chunks = (p.dim_pad + CHUNK - 1) / CHUNK;
  const float* q32 = static_cast<const float*>(p.q);

  struct Keys {
    uint64_t k[R];
  };
  auto load_keys = [&](int64_t b, int g) -> Keys {
    Keys ks;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = g * kSlots + r * 32 + lane;
      ks.k[r] = (b < p.B && g < n_groups && c < p.k_in) ? p.cand[b * p.k_in + c] : 0ull;
    }
    return ks;
  };

  // producer cursor: unit (pb, pg), chunk pc; this lane's keys of that unit, of the next unit of the
  // same query and of the first unit of the next query (candidate lists are filled from the front:
  // the first empty unit of a query ends it, and the cursor jumps to the next query)
  int64_t pu = u_begin, pb = u_begin / n_groups;
  int pg = int(u_begin - pb * n_groups), pc = 0, ps = 0;
  Keys p_key = load_keys(pb, pg);
  Keys p_key_next = load_keys(pb, pg + 1);
  Keys p_key_nq = load_keys(pb + 1, 0);

  // One pipeline step = (unit, chunk).  The first unit of a query without candidates (routed lists
  // of the sharded mode use ~1/G of their slots) takes ONE step — no copies, just the barrier
  // handshake — and ends the query: producer and consumer, which sees the same all-empty key
  // words, both jump to the next query.
  auto produce = [&]() {
    Stage& st = stages[ps];
    const uint32_t bar = ptx::smem_u32(&bars[ps]);
    const int d0 = pc * CHUNK;
    const int len = min(CHUNK, p.dim_pad - d0);
    int64_t row[R];
    unsigned n_valid = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row[r] = p_key.k[r] != 0 ? key_idx(p_key.k[r]) - p.idx_offset : -1;
      n_valid += __popc(__ballot_sync(kFull, row[r] >= 0));
    }
    if (lane == 0) ptx::mbar_expect_tx(bar, n_valid * uint32_t(len) * 4u);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (row[r] >= 0)
        bulk_g2s(ptx::smem_u32(st.rows + (r * 32 + lane) * Stage::kPitch), p.rows_a + row[r] * p.dim_pad + d0,
                 uint32_t(len) * 4u, bar);
    const float* qrow = q32 + pb * p.q_ld;
    if (n_valid == 0u) {
    } else if (q_vec16) {
      for (int i = lane; i < len / 4; i += 32) {
        const int d = d0 + 4 * i;
        const int nb = max(0, min(16, (p.dim - d) * 4));
        cp_async16_zfill(ptx::smem_u32(st.q + 4 * i), qrow + (nb > 0 ? d : 0), uint32_t(nb));
      }
    } else {
      for (int i = lane; i < len; i += 32) {
        const int d = d0 + i;
        cp_async4_zfill(ptx::smem_u32(st.q + i), qrow + (d < p.dim ? d : 0), d < p.dim ? 4u : 0u);
      }
    }
    cp_async_arrive_noinc(bar);
    if (pc == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) st.klo[r * 32 + lane] = row[r] >= 0 ? uint32_t(p_key.k[r]) : 0u;
    }
    if (++ps == STAGES) ps = 0;
    if (n_valid == 0u || ++pc == n_chunks) {
      pc = 0;
      if (n_valid == 0u || pg + 1 == n_groups) {  // on to the next query
        pu += n_groups - pg;
        ++pb;
        pg = 0;
        p_key = p_key_nq;
        p_key_nq = load_keys(pb + 1, 0);
      } else {
        ++pu;
        ++pg;
        p_key = p_key_next;
      }
      p_key_next = load_keys(pb, pg + 1);
    }
  };

  for (int t = 0; t < STAGES && pu < u_end; ++t) produce();

  int cs = 0, cc = 0;
  uint32_t phase = 0;
  uint32_t klo[R];
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    klo[r] = 0u;
    acc[r] = 0.0f;
  }
  int64_t cu = u_begin;
  int cg = int(u_begin % n_groups);
  bool live = true;  // the unit holds at least one candidate
  while (cu < u_end) {
    Stage& st = stages[cs];
    ptx::mbar_wait(ptx::smem_u32(&bars[cs]), phase, nullptr, 7);
    if (cc == 0) {
      bool any = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] = 0.0f;
        klo[r] = st.klo[r * 32 + lane];
        any = any || klo[r] != 0u;
      }
      live = __any_sync(kFull, any);
    }
    if (live) {
      const int len4 = min(CHUNK, p.dim_pad - cc * CHUNK) / 4;
      const float4* qr = reinterpret_cast<const float4*>(st.q);
      const float4* xr[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xr[r] = reinterpret_cast<const float4*>(st.rows + (r * 32 + lane) * Stage::kPitch);
      // len4 is a multiple of 16 (dim_pad is a multiple of 64).  The loads of a block of 4 float4 per
      // row are issued before its fma chains start, and the next block's loads while they run; the
      // R chains of a lane are independent and interleave in the pipeline.
      constexpr int kBlk = 4;
      float4 x[R][kBlk], w[kBlk];
#pragma unroll
      for (int i = 0; i < kBlk; ++i) {
        w[i] = qr[i];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r][i] = xr[r][i];
      }
#pragma unroll 1
      for (int j = 0; j < len4; j += kBlk) {
        float4 xn[R][kBlk], wn[kBlk];
        const int jn = (j + kBlk < len4) ? j + kBlk : j;
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
          wn[i] = qr[jn + i];
#pragma unroll
          for (int r = 0; r < R; ++r) xn[r][i] = xr[r][jn + i];
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = __fmaf_rn(w[i].x, x[r][i].x, acc[r]);
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = __fmaf_rn(w[i].y, x[r][i].y, acc[r]);
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = __fmaf_rn(w[i].z, x[r][i].z, acc[r]);
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = __fmaf_rn(w[i].w, x[r][i].w, acc[r]);
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
          w[i] = wn[i];
#pragma unroll
          for (int r = 0; r < R; ++r) x[r][i] = xn[r][i];
        }
      }
    }
    if (!live) {  // the rest of this query is empty (the select kernel masks by the candidate keys)
      cc = 0;
      cu += n_groups - cg;
      cg = 0;
    } else if (++cc == n_chunks) {
      cc = 0;
#pragma unroll
      for (int r = 0; r < R; ++r)
        tmp[cu * kSlots + r * 32 + lane] =
            klo[r] != 0 ? ((uint64_t(f32_to_orderable(acc[r])) << 32) | uint64_t(klo[r])) : 0ull;
      ++cu;
      if (++cg == n_groups) cg = 0;
    }
    // the stage is free: order this warp's generic-proxy reads before the async-proxy refill
    __syncwarp();
    ptx::fence_proxy_async();
    if (pu < u_end) produce();
    if (++cs == STAGES) {
      cs = 0;
      phase ^= 1;
    }
  }
}

// Sort the first `groups` 32-key groups of v descending with the smallest network that covers
// them (the other groups are empty and stay where they are: zeros sort last anyway).
template <int ITEMS>
__device__ __forceinline__ void sort_groups(uint64_t (&v)[ITEMS], int groups, int lane) {
  if (ITEMS > 2 && groups <= 2) {
    uint64_t w[2] = {v[0], v[1]};
    warp_sort_desc<2>(w, lane);
    v[0] = w[0];
    v[1] = w[1];
  } else if (ITEMS > 4 && groups <= 4) {
    uint64_t w[4] = {v[0], v[1], v[2], v[3]};
    warp_sort_desc<4>(w, lane);
#pragma unroll
    for (int r = 0; r < 4; ++r) v[r] = w[r];
  } else {
    warp_sort_desc<ITEMS>(v, lane);
  }
}

// exact keys of one query (n_groups*32 slots in the workspace) -> best k_out, sorted; certificate.
// Sharded re-scoring (p.n_peers > 0): the sorted exact keys of query b go straight into the
// exchange buffer of the GPU that owns query b (peer memory over NVLink), non-empty prefix only
// (the owner zeroes its buffer before the exchange); no certificate here — the owner evaluates
// it on the merged lists.
template <int ITEMS>
__global__ void __launch_bounds__(128)
    rescore_select_kernel(RescoreParams p, const uint64_t* __restrict__ tmp, int n_groups) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = int64_t(blockIdx.x) * 4 + warp;
  if (b >= p.B) return;
  const uint64_t* mine = tmp + b * n_groups * 32;
  uint64_t v[ITEMS];
  int groups = 0;  // groups holding at least one key (candidate lists are filled from the front)
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    // a slot holds an exact key only where the candidate list holds a candidate (lists are filled
    // from the front; the dot kernel does not touch the slots of empty units)
    const int c = r * 32 + lane;
    const bool has = r < n_groups && c < p.k_in && p.cand[b * p.k_in + c] != 0ull;
    v[r] = has ? mine[c] : 0ull;
    if (__any_sync(kFull, has)) groups = r + 1;
  }
  sort_groups<ITEMS>(v, groups, lane);
  if (p.n_peers > 0) {
    const int64_t owner = b / p.rows_per_owner;
    uint64_t* o = p.peer_out[owner] +
                  (int64_t(p.my_rank) * p.rows_per_owner + (b - owner * p.rows_per_owner)) * p.k_out;
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
      const int i = r * 32 + lane;
      if (i < p.k_out && v[r] != 0ull) o[i] = v[r];
    }
    return;
  }
  const float* qrow = static_cast<const float*>(p.q) + b * p.q_ld;
  float qq = 0.0f;
  for (int d = lane; d < p.dim; d += 32) qq = fmaf(qrow[d], qrow[d], qq);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);
  uint64_t kth = 0;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = r * 32 + lane;
    if (i < p.k_out) p.out[b * p.k_out + i] = v[r];
    kth = (i == p.k_out - 1) ? v[r] : kth;
  }
  kth = __shfl_sync(kFull, kth, (p.k_out - 1) & 31);
  if (lane == 0) {
    const uint64_t last = p.cand[b * p.k_in + p.k_in - 1];  // worst candidate under the approximate order
    const int ok = certified(p, kth, last, sqrtf(qq) * 1.001f);
    p.uncertified[b] = ok ? 0 : 1;
    if (!ok) atomicAdd(p.n_uncertified, 1);
  }
}

// The certificate alone (sharded fp32 mode: the exact keys of a query come back from the shards
// that own the candidate rows and are merged by the query's owner): p.out holds the merged exact
// keys (B, k_out), p.cand the merged approximate candidates (B, k_in).  One warp per query.
__global__ void __launch_bounds__(128) certify_kernel(RescoreParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = int64_t(blockIdx.x) * 4 + warp;
  if (b >= p.B) return;
  float qq = 0.0f;
  for (int d = lane; d < p.dim; d += 32) {
    const float v = ldq(p.q, p.q_dtype, b * p.q_ld + d);
    qq = fmaf(v, v, qq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(kFull, qq, o);
  if (lane == 0) {
    const int ok = certified(p, p.out[b * p.k_out + p.k_out - 1], p.cand[b * p.k_in + p.k_in - 1],
                             sqrtf(qq) * 1.001f);
    p.uncertified[b] = ok ? 0 : 1;
    if (!ok) atomicAdd(p.n_uncertified, 1);
  }
}

// out[g][r][:] = the keys of row r whose bank row belongs to shard g (rows [g*rows_per_shard, ...)),
// in their original order, compacted to the front and zero-padded: the query owner's candidate list
// split by the shard that can re-score each candidate.  Compact lists let the re-scoring kernel
// skip the empty 32-slot units (with G shards only ~1/G of the slots are used).  One warp per row.
__global__ void __launch_bounds__(128)
    route_keys_kernel(const uint64_t* __restrict__ keys, int64_t n, int k, int64_t rows_per_shard, int G,
                      uint64_t* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * 4 + warp;
  if (r >= n) return;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint64_t* row = keys + r * k;
  for (int g = 0; g < G; ++g) {
    uint64_t* o = out + (int64_t(g) * n + r) * k;
    int cnt = 0;
    for (int j0 = 0; j0 < k; j0 += 32) {
      const uint64_t key = j0 + lane < k ? row[j0 + lane] : 0ull;
      const bool mine = key != 0 && key_idx(key) / rows_per_shard == g;
      const unsigned bm = __ballot_sync(kFull, mine);
      if (mine) o[cnt + __popc(bm & lt_mask)] = key;
      cnt += __popc(bm);
    }
    for (int j = cnt + lane; j < k; j += 32) o[j] = 0ull;
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


template <int ITEMS>
cudaError_t launch_select_t(const RescoreParams& p, const uint64_t* tmp, int n_groups, cudaStream_t stream) {
  rescore_select_kernel<ITEMS><<<unsigned((p.B + 3) / 4), 128, 0, stream>>>(p, tmp, n_groups);
  return cudaGetLastError();
}

// experiment switches B200KNN_RESCORE_CHUNK=64|128 (columns per stage; 64 = two chains per lane) and B200KNN_RESCORE_STAGES=2|3
int dot_chunk() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B200KNN_RESCORE_CHUNK");
    v = (e != nullptr && atoi(e) == 128) ? 128 : 64;
  }
  return v;
}
int dot_rows() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B200KNN_RESCORE_ROWS");
    v = (e != nullptr && atoi(e) == 2) ? 2 : 1;
  }
  return v;
}
int dot_stages() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B200KNN_RESCORE_STAGES");
    const int x = e != nullptr ? atoi(e) : 0;
    v = (x == 3) ? 3 : 2;
  }
  return v;
}

template <int CHUNK, int STAGES, int R>
cudaError_t launch_dot_t(const RescoreParams& p, uint64_t* tmp, cudaStream_t stream) {
  constexpr int kSmem = 232448 - 1024;
  const size_t per_warp = STAGES * (sizeof(DotStage<CHUNK, R>) + sizeof(uint64_t));
  int warps = int(kSmem / per_warp);
  if (warps > 12) warps = 12;
  if (warps < 1) return cudaErrorNotSupported;
  const int n_groups = ws_slots(p.k_in) / (32 * R);
  const int64_t n_units = p.B * n_groups;
  int64_t blocks = (n_units + warps - 1) / warps;
  if (blocks > 148) blocks = 148;
  const size_t smem = per_warp * warps;
  auto kern = rescore_dot_kernel<CHUNK, STAGES, R>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  const uintptr_t qa = reinterpret_cast<uintptr_t>(p.q);
  const int q_vec16 = (qa % 16 == 0 && p.q_ld % 4 == 0) ? 1 : 0;
  kern<<<unsigned(blocks), warps * 32, smem, stream>>>(p, tmp, n_groups, n_units, q_vec16);
  return cudaGetLastError();
}
}  // namespace

size_t rescore_workspace_bytes(int64_t B, int k_in) { return size_t(B) * size_t(ws_slots(k_in)) * sizeof(uint64_t); }

cudaError_t launch_rescore(const RescoreParams& p, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream) {
  if (p.B == 0) return cudaSuccess;
  const int items = (p.k_in + 31) / 32;
  const bool pipelined = workspace != nullptr && workspace_bytes >= rescore_workspace_bytes(p.B, p.k_in) &&
                         p.rows_b == nullptr && p.q_dtype == B200KNN_F32 &&
                         reinterpret_cast<uintptr_t>(p.rows_a) % 16 == 0;
  if (p.n_peers > 0 && !pipelined) return cudaErrorInvalidValue;  // the scatter lives in the pipelined variant
  if (pipelined) {
    uint64_t* tmp = static_cast<uint64_t*>(workspace);
    cudaError_t e;
    const int ch = dot_chunk(), stg = dot_stages();
    // default: 64 columns x 32 candidates per stage, 11 warps per SM.  Measured at the north-star step
    // (74.5 GB of candidate rows, power-capped clocks): 128 columns / 6 warps 14.95 ms, 64 columns / 11 warps
    // 13.96 ms, 3-4 stages with fewer warps 19.6-26.2 ms, two fma chains per lane over 64 candidates x 64
    // columns (B200KNN_RESCORE_ROWS=2) 23.1 ms: the kernel is bound by the rate of its per-row bulk copies and
    // by how many warps keep them in flight, not by the dependent-fma latency.
    // small calls (the reference-shaped B = 64) are latency-bound chains of steps: fewer, larger steps
    const bool small_call = p.B * int64_t(ws_slots(p.k_in) / 32) < 148 * 12 * 8;
    if (dot_rows() == 2) e = launch_dot_t<64, 2, 2>(p, tmp, stream);
    else if ((ch == 128 || small_call) && stg == 2) e = launch_dot_t<128, 2, 1>(p, tmp, stream);
    else if (ch == 128) e = launch_dot_t<128, 3, 1>(p, tmp, stream);
    else if (stg == 3) e = launch_dot_t<64, 3, 1>(p, tmp, stream);
    else e = launch_dot_t<64, 2, 1>(p, tmp, stream);
    if (e != cudaSuccess) return e;
    const int g32 = ws_slots(p.k_in) / 32;  // 32-slot groups per query in the workspace (always even)
    if (g32 <= 2) return launch_select_t<2>(p, tmp, g32, stream);
    if (g32 <= 4) return launch_select_t<4>(p, tmp, g32, stream);
    if (g32 <= 8) return launch_select_t<8>(p, tmp, g32, stream);
    if (g32 <= 16) return launch_select_t<16>(p, tmp, g32, stream);
    if (g32 <= 32) return launch_select_t<32>(p, tmp, g32, stream);
    return cudaErrorInvalidValue;
  }
  if (items <= 2) return launch_rescore_t<2>(p, stream);
  if (items <= 4) return launch_rescore_t<4>(p, stream);
  if (items <= 8) return launch_rescore_t<8>(p, stream);
  if (items <= 16) return launch_rescore_t<16>(p, stream);
  if (items <= 32) return launch_rescore_t<32>(p, stream);
  return cudaErrorInvalidValue;
}

cudaError_t launch_certify(const RescoreParams& p, cudaStream_t stream) {
  if (p.B == 0) return cudaSuccess;
  certify_kernel<<<unsigned((p.B + 3) / 4), 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_route_keys(const uint64_t* keys, int64_t n, int k, int64_t rows_per_shard, int G,
                              uint64_t* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  route_keys_kernel<<<unsigned((n + 3) / 4), 128, 0, stream>>>(keys, n, k, rows_per_shard, G, out);
  return cudaGetLastError();
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
