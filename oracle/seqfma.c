/*
 * seqfma.c — ORACLE (test infrastructure only; never imported by the product).
 *
 * CPU restatement of the similarity + top-k head of lightly's knn_predict
 *     sim_matrix = torch.mm(feature, feature_bank)
 *     sim_weight, sim_indices = sim_matrix.topk(k=knn_k, dim=-1)
 * (reference call site src/ssl_wafermap/models/knn.py:91-98; lightly itself is
 * an unpinned, un-vendored dependency — requirements.txt:1 — so the arithmetic
 * is restated from its published 7-line algorithm)
 *
 * The restatement fixes what torch leaves unspecified, so that it can be
 * compared bit for bit with the CUDA "exact" mode:
 *   sim(q, n) = fmaf(q[D-1], b[D-1], ... fmaf(q[1], b[1], fmaf(q[0], b[0], +0.0f)))
 *   order     = (sim desc, bank index asc), -0.0 == +0.0
 * PARITY UNPINNED: the reference ships no tests / golden vectors for this path
 * (tests/conftest.py:1-10 is a dummy) and lightly cannot be imported here.
 *
 * Build: see oracle/Makefile (gcc -O3 -mfma; the Python wrapper threads over query rows; fmaf is correctly rounded
 * with or without hardware FMA, -mfma only makes it fast).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* sims[b*N + n] for bank in (D,N) layout (vectors are columns, ld = bank_ld). */
void seqfma_sims_dn(const float* q, int64_t q_ld, const float* bank, int64_t bank_ld, int64_t B,
                    int64_t N, int64_t D, float* sims) {
  for (int64_t b = 0; b < B; ++b) {
    float* out = sims + b * N;
    for (int64_t n = 0; n < N; ++n) out[n] = 0.0f;
    for (int64_t d = 0; d < D; ++d) {
      const float qd = q[b * q_ld + d];
      const float* row = bank + d * bank_ld;
      for (int64_t n = 0; n < N; ++n) out[n] = fmaf(qd, row[n], out[n]);
    }
  }
}

static inline uint32_t orderable(float s) {
  s = s + 0.0f;
  uint32_t u;
  memcpy(&u, &s, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

static int cmp_desc_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? 1 : (x > y ? -1 : 0);
}

/* Canonical top-k of each row of sims (B,N): out_sims (B,k), out_idx (B,k). */
void canonical_topk(const float* sims, int64_t B, int64_t N, int64_t k, int64_t idx_offset,
                    float* out_sims, int64_t* out_idx) {
  {
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)N);
    for (int64_t b = 0; b < B; ++b) {
      const float* row = sims + b * N;
      for (int64_t n = 0; n < N; ++n)
        keys[n] = ((uint64_t)orderable(row[n]) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(n + idx_offset));
      qsort(keys, (size_t)N, sizeof(uint64_t), cmp_desc_u64);
      for (int64_t j = 0; j < k; ++j) {
        const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(keys[j] & 0xFFFFFFFFu);
        out_idx[b * k + j] = (int64_t)idx;
        out_sims[b * k + j] = row[idx - idx_offset] + 0.0f;
      }
    }
    free(keys);
  }
}
