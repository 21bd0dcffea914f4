// tc_topk.cu — MODE_BF16 / MODE_TF32X3: the fused similarity contraction +
// streaming per-row top-k on the sm_100a tensor cores.
//
// Replaces torch.mm(feature, feature_bank) + sim_matrix.topk(k) of lightly's
// knn_predict (reference call site src/ssl_wafermap/models/knn.py:91-98).  The
// (B,N) similarity matrix lives only in TMEM: one 128 x BLOCK_N fp32 tile at a
// time, double buffered, drained by the epilogue warps straight into the
// per-row candidate lists of common.cuh.
//
// CTA = 10 warps, one CTA per SM, persistent over work items
// (query tile of 128 rows) x (bank split):
//   warp 0  : TMA producer  (one elected lane)   global -> smem, SWIZZLE_128B
//   warp 1  : tcgen05.mma issuer (one lane); owns the TMEM allocation
//   warp 2-9: epilogue; warp w may only touch TMEM lanes 32*(w%4)..+31, so two
//             warps share each lane quarter and each owns 16 of its 32 query
//             rows (two warps per scheduler hide each other's latencies; the
//             selection code is latency-bound, measured CPI 8 with one warp).
// Candidate extraction is warp-cooperative: a row whose 32-column chunk holds a
// similarity above its threshold is staged through 128 B of shared memory so
// that lane j tests column j (one ballot, coalesced appends), instead of one
// thread walking its 32 registers through 32 divergent branches.
// Pipelines (mbarrier): smem stage full/empty (TMA <-> MMA), TMEM accumulator
// full/empty (MMA <-> epilogue), query tile full/empty (BF16 mode keeps the
// 128 x D query tile resident in smem for the whole work item).
//
// MODE_BF16  : operands bf16 K-major, kind::f16, UMMA 128 x BLOCK_N x 16.
// MODE_TF32X3: operands fp32 hi/lo split (prepare.cu), kind::tf32, UMMA
//              128 x BLOCK_N x 8; per k-step  hi*lo + lo*hi + hi*hi  accumulate
//              into the same TMEM tile; A and B are both streamed per k-block.
// The full D is accumulated in one fixed order in one MMA chain (no split-K),
// so sim(q, n) does not depend on tiling, batch size, split or shard count.
#include <cuda.h>

#include "../../include/b200knn.h"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace b200knn {

int tc_tile_n(int mode, int dim) {
  if (mode == B200KNN_MODE_BF16) return ((dim + 63) / 64 * 64) <= 512 ? 256 : 128;
  return 128;
}

namespace {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kRowBytes = 128;                   // one swizzle row = one k-block of a vector
constexpr int kABlockBytes = kTileM * kRowBytes;  // 16 KB: 128 query rows x one k-block
#ifndef B200KNN_EPI_PER_QUARTER
#define B200KNN_EPI_PER_QUARTER 2
#endif
constexpr int kEpiPerQuarter = B200KNN_EPI_PER_QUARTER;  // epilogue warps per TMEM lane quarter
constexpr int kEpiWarps = 4 * kEpiPerQuarter;
constexpr int kRowsPerWarp = 32 / kEpiPerQuarter;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemLimit = 232448;  // 227 KB

struct alignas(16) Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t q_full;
  uint64_t q_empty;
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float stage[kEpiWarps][32];  // per epilogue warp: one row's 32-column chunk, lane j <-> column j
};

struct TcKernelArgs {
  int64_t B, N;
  int k;
  int n_kblocks;  // D_pad / elements-per-128B-row
  int n_stages;
  int64_t idx_offset;
  int64_t n_qtiles, n_items, split_rows;
  uint64_t* lists;
  uint64_t* out;
  const float* tau0;  // optional (B,) initial admission thresholds (nullptr: -inf)
  float* dump;  // optional (B, N) fp32 similarity dump for unit tests (nullptr in production)
  int32_t* diag;
  int flags;    // experiment switches of the debug entry point (0 in production):
                // 1 no selection, 2 no TMEM loads, 4 no MMA issue, 8 no bank TMA loads
};

template <int MODE, int BLOCK_N, int ITEMS, bool DEBUG>
__global__ void __launch_bounds__(kThreads, 1)
    tc_topk_kernel(const __grid_constant__ CUtensorMap map_q_hi,
                   const __grid_constant__ CUtensorMap map_q_lo,
                   const __grid_constant__ CUtensorMap map_b_hi,
                   const __grid_constant__ CUtensorMap map_b_lo, const TcKernelArgs a) {
  constexpr bool kBf16 = (MODE == B200KNN_MODE_BF16);
  constexpr int CAP = ITEMS * 32;
  constexpr int kBBlockBytes = BLOCK_N * kRowBytes;
  constexpr int kStageBytes = kBf16 ? kBBlockBytes : 2 * (kABlockBytes + kBBlockBytes);
  constexpr int kElemsPerRow = kBf16 ? 64 : 32;  // elements of one 128-byte k-block row
  constexpr int kUmmaKBytes = 32;                // one MMA consumes 32 bytes of k per row
  constexpr uint32_t kIdesc = ptx::make_idesc(kBf16 ? 1u : 2u, kTileM, BLOCK_N);
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int q_bytes = kBf16 ? a.n_kblocks * kABlockBytes : 0;
  uint8_t* q_smem = smem;
  uint8_t* stage_smem = smem + q_bytes;
  Barriers* bars = reinterpret_cast<Barriers*>(stage_smem + size_t(a.n_stages) * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_q_hi);
    ptx::prefetch_tensormap(&map_b_hi);
    if (!kBf16) {
      ptx::prefetch_tensormap(&map_q_lo);
      ptx::prefetch_tensormap(&map_b_lo);
    }
    for (int s = 0; s < a.n_stages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_full[b]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->tmem_empty[b]), kEpiWarps);  // one arrive per epilogue warp
    }
    ptx::mbar_init(ptx::smem_u32(&bars->q_full), 1);
    ptx::mbar_init(ptx::smem_u32(&bars->q_empty), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, q_phase = 0;
      for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int64_t qt = item % a.n_qtiles, sp = item / a.n_qtiles;
        const int m0 = int(qt * kTileM);
        const int64_t n_begin = sp * a.split_rows;
        const int64_t n_end = (n_begin + a.split_rows < a.N) ? n_begin + a.split_rows : a.N;
        if (kBf16) {
          ptx::mbar_wait(ptx::smem_u32(&bars->q_empty), q_phase ^ 1, a.diag, 1);
          ptx::mbar_expect_tx(ptx::smem_u32(&bars->q_full), uint32_t(q_bytes));
          for (int kb = 0; kb < a.n_kblocks; ++kb)
            ptx::tma_load_2d(ptx::smem_u32(q_smem + kb * kABlockBytes), &map_q_hi,
                             kb * kElemsPerRow, m0, ptx::smem_u32(&bars->q_full));
          q_phase ^= 1;
        }
        for (int64_t n0 = n_begin; n0 < n_end; n0 += BLOCK_N) {
          for (int kb = 0; kb < a.n_kblocks; ++kb) {
            ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1, a.diag, 2);
            const uint32_t full = ptx::smem_u32(&bars->full[stage]);
            uint8_t* st = stage_smem + size_t(stage) * kStageBytes;
            if (DEBUG && (a.flags & 8)) {
              ptx::mbar_arrive(full);
            } else if (kBf16) {
              ptx::mbar_expect_tx(full, uint32_t(kStageBytes));
              ptx::tma_load_2d(ptx::smem_u32(st), &map_b_hi, kb * kElemsPerRow, int(n0), full);
            } else {
              ptx::mbar_expect_tx(full, uint32_t(kStageBytes));
              ptx::tma_load_2d(ptx::smem_u32(st), &map_q_hi, kb * kElemsPerRow, m0, full);
              ptx::tma_load_2d(ptx::smem_u32(st + kABlockBytes), &map_q_lo, kb * kElemsPerRow, m0,
                               full);
              ptx::tma_load_2d(ptx::smem_u32(st + 2 * kABlockBytes), &map_b_hi, kb * kElemsPerRow,
                               int(n0), full);
              ptx::tma_load_2d(ptx::smem_u32(st + 2 * kABlockBytes + kBBlockBytes), &map_b_lo,
                               kb * kElemsPerRow, int(n0), full);
            }
            if (++stage == a.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, q_phase = 0;
      uint32_t tcount = 0;
      for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int64_t sp = item / a.n_qtiles;
        const int64_t n_begin = sp * a.split_rows;
        const int64_t n_end = (n_begin + a.split_rows < a.N) ? n_begin + a.split_rows : a.N;
        if (kBf16) {
          ptx::mbar_wait(ptx::smem_u32(&bars->q_full), q_phase, a.diag, 3);
          q_phase ^= 1;
        }
        for (int64_t n0 = n_begin; n0 < n_end; n0 += BLOCK_N) {
          const uint32_t buf = tcount & 1u, aphase = (tcount >> 1) & 1u;
          ptx::mbar_wait(ptx::smem_u32(&bars->tmem_empty[buf]), aphase ^ 1, a.diag, 4);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * BLOCK_N;
          for (int kb = 0; kb < a.n_kblocks; ++kb) {
            ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase, a.diag, 5);
            ptx::tc_fence_after();
            const uint32_t st = ptx::smem_u32(stage_smem + size_t(stage) * kStageBytes);
            if (DEBUG && (a.flags & 4)) {
            } else if (kBf16) {
              const uint32_t qa = ptx::smem_u32(q_smem + kb * kABlockBytes);
#pragma unroll
              for (int ks = 0; ks < kRowBytes / kUmmaKBytes; ++ks) {
                ptx::umma_f16(tmem_d, ptx::smem_desc_sw128(qa + ks * kUmmaKBytes),
                              ptx::smem_desc_sw128(st + ks * kUmmaKBytes), kIdesc,
                              uint32_t((kb | ks) != 0));
              }
            } else {
              const uint32_t a_hi = st, a_lo = st + kABlockBytes;
              const uint32_t b_hi = st + 2 * kABlockBytes, b_lo = b_hi + kBBlockBytes;
#pragma unroll
              for (int ks = 0; ks < kRowBytes / kUmmaKBytes; ++ks) {
                const uint32_t o = ks * kUmmaKBytes;
                ptx::umma_tf32(tmem_d, ptx::smem_desc_sw128(a_hi + o),
                               ptx::smem_desc_sw128(b_lo + o), kIdesc, uint32_t((kb | ks) != 0));
                ptx::umma_tf32(tmem_d, ptx::smem_desc_sw128(a_lo + o),
                               ptx::smem_desc_sw128(b_hi + o), kIdesc, 1u);
                ptx::umma_tf32(tmem_d, ptx::smem_desc_sw128(a_hi + o),
                               ptx::smem_desc_sw128(b_hi + o), kIdesc, 1u);
              }
            }
            ptx::umma_commit(ptx::smem_u32(&bars->empty[stage]));  // frees the smem stage
            if (++stage == a.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          ptx::umma_commit(ptx::smem_u32(&bars->tmem_full[buf]));  // accumulator ready
          ++tcount;
        }
        if (kBf16) ptx::umma_commit(ptx::smem_u32(&bars->q_empty));  // query tile reusable
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int sub = (warp - 2) / 4;           // which of the quarter's warps
    const bool owner = (lane / kRowsPerWarp) == sub;  // this lane's row is selected by this warp
    const int row_in_tile = quarter * 32 + lane;
    uint64_t* warp_lists = a.lists + (size_t(blockIdx.x) * kTileM + size_t(quarter) * 32) * CAP;
    const uint32_t stage = ptx::smem_u32(bars->stage[warp - 2]);
    const float neg_inf = __int_as_float(0xff800000);
    const float pos_inf = __int_as_float(0x7f800000);
    const unsigned lt_mask = (1u << lane) - 1u;
    // free slots below which a row is pruned between tiles (off the critical path)
    const int soft_slack = min(96, (CAP - a.k) / 2);
    uint32_t tcount = 0;
    for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int64_t qt = item % a.n_qtiles, sp = item / a.n_qtiles;
      const int64_t m0 = qt * kTileM;
      const int64_t n_begin = sp * a.split_rows;
      const int64_t n_end = (n_begin + a.split_rows < a.N) ? n_begin + a.split_rows : a.N;
      const int64_t grow = m0 + row_in_tile;
      RowState st;
      st.cnt = 0;
      st.tau = (owner && grow < a.B) ? (a.tau0 != nullptr ? a.tau0[grow] : neg_inf) : pos_inf;
      for (int64_t n0 = n_begin; n0 < n_end; n0 += BLOCK_N) {
        const uint32_t buf = tcount & 1u, aphase = (tcount >> 1) & 1u;
        ptx::mbar_wait(ptx::smem_u32(&bars->tmem_full[buf]), aphase, a.diag, 6);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + buf * BLOCK_N;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          float s[32];
          if (DEBUG && (a.flags & 2)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) s[j] = neg_inf;
          } else {
            ptx::tmem_ld32(taddr + c0, s);
          }
          if (c0 + 32 == BLOCK_N) {
            // all of this warp's TMEM reads of the buffer are done: hand it back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->tmem_empty[buf]));
          }
          if (DEBUG && a.dump != nullptr && owner && grow < a.B) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t gn = n0 + c0 + j;
              if (gn < n_end) a.dump[grow * a.N + gn] = s[j];
            }
          }
          float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
          for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], s[j]);
          const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          unsigned hits = __ballot_sync(kFull, mx > st.tau);
          if (DEBUG && (a.flags & 1)) hits = 0;
          if (hits == 0) continue;
          // ---- some row of this warp has a candidate in this chunk (about half the chunks)
          const int64_t gn = n0 + c0 + lane;
          const bool col_ok = gn < n_end;
          const uint32_t gidx = uint32_t(gn + a.idx_offset);
          while (hits) {
            const int r = __ffs(hits) - 1;
            hits &= hits - 1;
            if (lane == r) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                ptx::st_shared_v4(stage + j * 4, s[j], s[j + 1], s[j + 2], s[j + 3]);
            }
            __syncwarp();
            const float v = ptx::ld_shared_f32(stage + lane * 4);
            const float tr = __shfl_sync(kFull, st.tau, r);
            const uint32_t cr = __shfl_sync(kFull, st.cnt, r);
            const bool p = (v > tr) && col_ok;
            const unsigned pm = __ballot_sync(kFull, p);
            if (p) warp_lists[size_t(r) * CAP + cr + __popc(pm & lt_mask)] = make_key(v, gidx);
            if (lane == r) st.cnt += __popc(pm);
            __syncwarp();
          }
          warp_maintain<ITEMS>(warp_lists, st, a.k, lane, 32);  // emergency only (list would overflow)
        }
        // The TMEM buffer went back to the MMA warp before the last chunk was processed:
        // prune here, off the accumulator's critical path, a little before it becomes mandatory.
        warp_maintain<ITEMS>(warp_lists, st, a.k, lane, soft_slack);
        ++tcount;
      }
      const int64_t row0 = m0 + quarter * 32;
      unsigned valid = 0;
      if (row0 < a.B) {
        const int64_t nv = a.B - row0;
        valid = nv >= 32 ? kFull : ((1u << nv) - 1u);
      }
      valid &= (kRowsPerWarp == 32 ? kFull : ((1u << kRowsPerWarp) - 1u)) << (sub * kRowsPerWarp);
      uint64_t* out = a.out + (size_t(sp) * a.B + row0) * a.k;
      warp_flush<ITEMS>(warp_lists, st, a.k, lane, out, size_t(a.k), valid);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// rows x cols matrix of `esize`-byte elements, row pitch = cols*esize; box = box_rows x 128 bytes
bool make_map(CUtensorMap* m, const void* base, bool bf16, uint64_t rows, uint64_t cols,
              uint32_t box_rows, uint64_t row_stride = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const uint64_t esize = bf16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * esize * row_stride};  // row_stride > 1: every row_stride-th row
  cuuint32_t box[2] = {cuuint32_t(kRowBytes / esize), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int MODE, int BLOCK_N, int ITEMS, bool DEBUG>
cudaError_t launch_t(const TcParams& p, int grid, cudaStream_t stream, float* dump, int32_t* diag,
                     int flags, const char** why) {
  constexpr bool kBf16 = (MODE == B200KNN_MODE_BF16);
  const int d_pad = (p.D + 63) / 64 * 64;
  const int elems_per_row = kBf16 ? 64 : 32;
  TcKernelArgs a;
  a.B = p.B;
  a.N = p.N;
  a.k = p.k;
  a.n_kblocks = d_pad / elems_per_row;
  a.idx_offset = p.idx_offset;
  a.n_qtiles = p.n_qtiles;
  a.n_items = p.n_items;
  a.split_rows = p.split_rows;
  a.lists = p.lists;
  a.out = p.out;
  a.tau0 = p.tau0;
  a.dump = dump;
  a.diag = diag;
  a.flags = flags;
  const int b_block = BLOCK_N * kRowBytes;
  const int stage_bytes = kBf16 ? b_block : 2 * (kABlockBytes + b_block);
  const int q_bytes = kBf16 ? a.n_kblocks * kABlockBytes : 0;
  const int fixed = q_bytes + int(sizeof(Barriers)) + 1024;  // 1024: manual alignment slack
  int stages = (kSmemLimit - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    *why = "vector dimension too large for the resident query tile";
    return cudaErrorNotSupported;
  }
  a.n_stages = stages;
  const size_t smem = size_t(fixed) + size_t(stages) * stage_bytes;

  CUtensorMap mq_hi, mq_lo, mb_hi, mb_lo;
  bool ok = make_map(&mq_hi, p.q_hi, kBf16, uint64_t(p.B), uint64_t(d_pad), kTileM) &&
            make_map(&mb_hi, p.bank_hi, kBf16, uint64_t(p.N), uint64_t(d_pad), BLOCK_N, uint64_t(p.bank_row_stride));
  if (ok && !kBf16)
    ok = make_map(&mq_lo, p.q_lo, false, uint64_t(p.B), uint64_t(d_pad), kTileM) &&
         make_map(&mb_lo, p.bank_lo, false, uint64_t(p.N), uint64_t(d_pad), BLOCK_N, uint64_t(p.bank_row_stride));
  if (!ok) {
    *why = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  if (kBf16) {
    mq_lo = mq_hi;
    mb_lo = mb_hi;
  }
  auto kern = tc_topk_kernel<MODE, BLOCK_N, ITEMS, DEBUG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, stream>>>(mq_hi, mq_lo, mb_hi, mb_lo, a);
  return cudaGetLastError();
}

template <int MODE, int BLOCK_N, bool DEBUG>
cudaError_t launch_cap(const TcParams& p, int grid, int cap, cudaStream_t stream, float* dump,
                       int32_t* diag, int flags, const char** why) {
  switch (cap) {
    case 64: return launch_t<MODE, BLOCK_N, 2, DEBUG>(p, grid, stream, dump, diag, flags, why);
    case 128: return launch_t<MODE, BLOCK_N, 4, DEBUG>(p, grid, stream, dump, diag, flags, why);
    case 256: return launch_t<MODE, BLOCK_N, 8, DEBUG>(p, grid, stream, dump, diag, flags, why);
    case 512: return launch_t<MODE, BLOCK_N, 16, DEBUG>(p, grid, stream, dump, diag, flags, why);
    case 1024: return launch_t<MODE, BLOCK_N, 32, DEBUG>(p, grid, stream, dump, diag, flags, why);
    default: *why = "unsupported k"; return cudaErrorNotSupported;
  }
}

template <bool DEBUG>
cudaError_t launch_mode(const TcParams& p, int grid, int cap, cudaStream_t stream, float* dump,
                        int32_t* diag, int flags, const char** why) {
  if (p.mode == B200KNN_MODE_BF16) {
    if (tc_tile_n(p.mode, p.D) == 256)
      return launch_cap<B200KNN_MODE_BF16, 256, DEBUG>(p, grid, cap, stream, dump, diag, flags, why);
    return launch_cap<B200KNN_MODE_BF16, 128, DEBUG>(p, grid, cap, stream, dump, diag, flags, why);
  }
  if (p.mode == B200KNN_MODE_TF32X3)
    return launch_cap<B200KNN_MODE_TF32X3, 128, DEBUG>(p, grid, cap, stream, dump, diag, flags, why);
  *why = "unknown mode";
  return cudaErrorNotSupported;
}

}  // namespace

cudaError_t launch_tc(const TcParams& p, int grid, int cap, cudaStream_t stream, float* dump,
                      int32_t* diag, int flags, const char** why) {
  *why = "";
  // the debug instantiation (dump / experiment flags) is only reachable from the test hook
  if (dump != nullptr || flags != 0) return launch_mode<true>(p, grid, cap, stream, dump, diag, flags, why);
  return launch_mode<false>(p, grid, cap, stream, dump, diag, flags, why);
}

}  // namespace b200knn
