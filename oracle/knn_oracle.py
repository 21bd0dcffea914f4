"""ORACLE — test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the checker or
the timed CPU baseline — never as part of the product path (the product has no
CPU fallback and raises when the CUDA extension is missing).

CPU restatement of the hot path of faris-k/self-supervised-wafermaps:
``lightly.utils.benchmarking.knn_predict`` as bound at
``src/ssl_wafermap/models/knn.py:16`` and called at ``:91-98`` / ``:205-212``.

The arithmetic lives in a third-party dependency that is absent from
``/root/reference``: **lightly, unpinned** (``requirements.txt:1``; the repo
dates from Apr–May 2023, i.e. lightly ≈1.4.x) and not installable here (no
network).  ``knn_predict_r32`` restates its published 7-line algorithm
(InstDisc / MoCo demo code) verbatim on torch CPU fp32, anchored
on the reference's own call sites and caller conventions
(``knn.py:77,80-81,90,99``).

**PARITY UNPINNED**: the reference holds no tests, golden vectors or fixtures
for this path (``tests/conftest.py:1-10`` is a PyScaffold dummy) and lightly
cannot be imported to generate outputs.  The committed goldens under
``tests/golden/`` are outputs of THIS restatement (``tests/golden/make_golden.py``)
on the reference's two shipped real embedding banks
(``data/interim/model_preds/{FastSiam,SimSiam}_preds_subset.pkl.xz``) and on
seeded synthetic inputs; they pin the restatement against regressions, not
against lightly.

Three oracles:
  R32   ``knn_predict_r32``  — the reference algorithm, torch CPU fp32, op for op.
  O64   ``topk_o64`` / ``vote_o64`` — fp64 similarities, canonical total orders
        (sim desc, index asc) / (score desc, class asc): what "the right answer" is.
  SEQ   ``topk_seqfma`` — fp32 sequential-fma similarities + canonical order in C
        (oracle/seqfma.c): bit-exact model of the CUDA "exact" mode.
Comparators implement the ambiguity-aware contract of SURVEY.md §7.3.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- R32
def knn_predict_r32_full(feature, feature_bank, feature_labels, num_classes, knn_k=200, knn_t=0.1):
    """lightly's knn_predict, restated op for op (torch, any device, input dtype).

    Returns (pred_labels, sim_weight_before_exp, sim_indices, pred_scores)."""
    import torch

    # lightly/utils/benchmarking.py knn_predict [recalled, SURVEY.md §3.2] line by line:
    sim_matrix = torch.mm(feature, feature_bank)  # (B,D)@(D,N) -> (B,N)
    sim_weight, sim_indices = sim_matrix.topk(k=knn_k, dim=-1)  # (B,K)
    sim_labels = torch.gather(feature_labels.expand(feature.size(0), -1), dim=-1, index=sim_indices)
    sims = sim_weight
    sim_weight = (sim_weight / knn_t).exp()
    one_hot_label = torch.zeros(feature.size(0) * knn_k, num_classes, device=sim_labels.device)
    one_hot_label = one_hot_label.scatter(dim=-1, index=sim_labels.view(-1, 1), value=1.0)
    pred_scores = torch.sum(
        one_hot_label.view(feature.size(0), -1, num_classes) * sim_weight.unsqueeze(dim=-1), dim=1)
    pred_labels = pred_scores.argsort(dim=-1, descending=True)
    return pred_labels, sims, sim_indices, pred_scores


def knn_predict_r32(feature, feature_bank, feature_labels, num_classes, knn_k=200, knn_t=0.1):
    """Same signature and return value as the reference symbol."""
    return knn_predict_r32_full(feature, feature_bank, feature_labels, num_classes, knn_k, knn_t)[0]


# --------------------------------------------------------------------------- O64
def sims_o64(feature: np.ndarray, feature_bank: np.ndarray) -> np.ndarray:
    return np.asarray(feature, dtype=np.float64) @ np.asarray(feature_bank, dtype=np.float64)


def canonical_topk_np(sims: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k of each row under (sim desc, index asc).  Returns (sims (B,k), idx (B,k) int64)."""
    B, N = sims.shape
    s = sims + 0.0  # -0.0 -> +0.0
    # stable argsort on the negated values keeps ascending index order among equals
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order.astype(np.int64)


def topk_o64(feature: np.ndarray, feature_bank: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    return canonical_topk_np(sims_o64(feature, feature_bank), k)


def vote_o64(sims: np.ndarray, idx: np.ndarray, labels: np.ndarray, num_classes: int,
             knn_t: float) -> Tuple[np.ndarray, np.ndarray]:
    """fp64 class scores accumulated in rank order and the canonical class ranking
    (score desc, class asc).  sims are used at the precision given (pass fp32 sims to model
    the CUDA vote kernel).  Returns (pred (B,C) int64, scores (B,C) f64)."""
    B, K = sims.shape
    w = np.exp(sims.astype(np.float64) / np.float64(knn_t))
    lab = np.asarray(labels)[idx]
    if lab.min(initial=0) < 0 or lab.max(initial=0) >= num_classes:
        raise RuntimeError("index out of bounds: label outside [0, num_classes)")
    scores = np.zeros((B, num_classes), dtype=np.float64)
    rows = np.arange(B)
    for j in range(K):  # rank order, one neighbour at a time: fixed accumulation order
        np.add.at(scores, (rows, lab[:, j]), w[:, j])
    pred = np.argsort(-scores, axis=1, kind="stable").astype(np.int64)
    return pred, scores


def knn_predict_o64(feature, feature_bank, feature_labels, num_classes, knn_k=200, knn_t=0.1):
    s, i = topk_o64(feature, feature_bank, knn_k)
    return vote_o64(s, i, feature_labels, num_classes, knn_t)[0]


# --------------------------------------------------------------------------- SEQ (C)
_seq = None


def build_c(force: bool = False) -> str:
    """Compile oracle/seqfma.c -> oracle/_build/libseqfma.so (gcc)."""
    out = os.path.join(_HERE, "_build", "libseqfma.so")
    src = os.path.join(_HERE, "seqfma.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return out


def _seqlib():
    global _seq
    if _seq is None:
        lib = ctypes.CDLL(build_c())
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        lib.seqfma_sims_dn.argtypes = [vp, i64, vp, i64, i64, i64, i64, vp]
        lib.seqfma_sims_dn.restype = None
        lib.canonical_topk.argtypes = [vp, i64, i64, i64, i64, vp, vp]
        lib.canonical_topk.restype = None
        _seq = lib
    return _seq


def sims_seqfma(feature: np.ndarray, feature_bank: np.ndarray, threads: Optional[int] = None) -> np.ndarray:
    """fp32 similarities, each an fmaf chain over d = 0..D-1 from +0.0f (bitwise model of the
    CUDA exact mode).  feature (B,D) f32, feature_bank (D,N) f32."""
    lib = _seqlib()
    q = np.ascontiguousarray(feature, dtype=np.float32)
    bank = np.ascontiguousarray(feature_bank, dtype=np.float32)
    B, D = q.shape
    N = bank.shape[1]
    out = np.empty((B, N), dtype=np.float32)
    threads = threads or min(os.cpu_count() or 1, max(1, B))
    bounds = np.linspace(0, B, threads + 1).astype(int)

    def run(t):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        if hi > lo:
            lib.seqfma_sims_dn(q[lo:].ctypes.data, D, bank.ctypes.data, N, hi - lo, N, D,
                               out[lo:].ctypes.data)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(run, range(threads)))
    return out


def canonical_topk_c(sims: np.ndarray, k: int, idx_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    lib = _seqlib()
    s = np.ascontiguousarray(sims, dtype=np.float32)
    B, N = s.shape
    out_s = np.empty((B, k), dtype=np.float32)
    out_i = np.empty((B, k), dtype=np.int64)
    threads = min(os.cpu_count() or 1, max(1, B))
    bounds = np.linspace(0, B, threads + 1).astype(int)

    def run(t):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        if hi > lo:
            lib.canonical_topk(s[lo:].ctypes.data, hi - lo, N, k, idx_offset, out_s[lo:].ctypes.data,
                               out_i[lo:].ctypes.data)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(run, range(threads)))
    return out_s, out_i


def topk_seqfma(feature: np.ndarray, feature_bank: np.ndarray, k: int,
                idx_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    return canonical_topk_c(sims_seqfma(feature, feature_bank), k, idx_offset)


# --------------------------------------------------------------------------- keys
def orderable_u32(s: np.ndarray) -> np.ndarray:
    b = (np.asarray(s, dtype=np.float32) + np.float32(0.0)).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def make_keys(sims: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """The 64-bit selection key of include/b200knn.h, as uint64."""
    return (orderable_u32(sims).astype(np.uint64) << np.uint64(32)) | (
        np.uint64(0xFFFFFFFF) - np.asarray(idx).astype(np.uint64))


def decode_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.asarray(keys).view(np.uint64)
    u = (keys >> np.uint64(32)).astype(np.uint32)
    b = np.where(u & np.uint32(0x80000000), u & np.uint32(0x7FFFFFFF), ~u).astype(np.uint32)
    sims = b.view(np.float32).copy()
    idx = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = keys == 0
    sims[empty] = -np.inf
    idx[empty] = -1
    return sims, idx


def merge_keys_np(keys_in: np.ndarray, k_out: int) -> np.ndarray:
    """(G,B,k_in) uint64 -> (B,k_out): top-k_out of the union per row, descending."""
    G, B, k_in = keys_in.shape
    allk = np.transpose(np.asarray(keys_in).view(np.uint64), (1, 0, 2)).reshape(B, G * k_in)
    return np.sort(allk, axis=1)[:, ::-1][:, :k_out].copy()


# --------------------------------------------------------------------------- comparators
def compare_topk(got_sims: np.ndarray, got_idx: np.ndarray, feature: np.ndarray,
                 feature_bank: np.ndarray, k: int, rel_tol: float = 1e-5,
                 eps_scale: float = 4e-7) -> Dict[str, float]:
    """Ambiguity-aware comparison of a (B,k) top-k result with the fp64 canonical oracle
    (SURVEY.md §7.3 contract).

    A position j of a row is *unambiguous* when the fp64 similarity at rank j differs from
    both rank-neighbours (j-1, j+1; rank k counts) by more than eps = eps_scale*max|s|.
    Returns counts; the caller asserts:
      idx_mismatch_unambiguous == 0, set_mismatch_rows_unambiguous == 0, max_rel_err <= rel_tol.
    """
    s64 = sims_o64(feature, feature_bank)
    B, N = s64.shape
    kk = min(k + 1, N)
    o_s, o_i = canonical_topk_np(s64, kk)
    smax = float(np.abs(s64).max()) if s64.size else 1.0
    eps = eps_scale * max(smax, 1e-30)
    gaps = np.abs(np.diff(o_s, axis=1))  # gaps[j] = s[j]-s[j+1], j in [0,kk-1)
    gap_next = np.full((B, k), np.inf)
    gap_prev = np.full((B, k), np.inf)
    gap_next[:, : gaps.shape[1]] = gaps[:, :k]
    gap_prev[:, 1:] = gaps[:, : k - 1]
    unamb = (gap_next > eps) & (gap_prev > eps)
    idx_eq = got_idx == o_i[:, :k]
    # similarity accuracy: compare got sim with the fp64 sim OF THE RETURNED INDEX
    safe_idx = np.clip(got_idx, 0, N - 1)
    s_at = np.take_along_axis(s64, safe_idx, axis=1)
    rel = np.abs(got_sims.astype(np.float64) - s_at) / max(smax, 1e-30)
    # set equality unless the k/k+1 boundary is ambiguous
    boundary_amb = gap_next[:, k - 1] <= eps
    set_eq = np.array([set(got_idx[b]) == set(o_i[b, :k]) for b in range(B)])
    return {
        "rows": float(B),
        "positions": float(B * k),
        "ambiguous_positions": float((~unamb).sum()),
        "idx_mismatch_total": float((~idx_eq).sum()),
        "idx_mismatch_unambiguous": float((~idx_eq & unamb).sum()),
        "set_mismatch_rows": float((~set_eq).sum()),
        "set_mismatch_rows_unambiguous": float((~set_eq & ~boundary_amb).sum()),
        "max_rel_err": float(rel.max()) if rel.size else 0.0,
        "recall_at_k": float(np.mean([len(set(got_idx[b]) & set(o_i[b, :k])) / k for b in range(B)])),
    }


def compare_pred(got_pred: np.ndarray, scores64: np.ndarray, rel_margin: float = 1e-5) -> Dict[str, float]:
    """Class-ranking comparison against fp64 scores: position r of a row must hold the oracle's
    class whenever that class's score differs from its rank-neighbours' by > rel_margin*max."""
    B, C = scores64.shape
    order = np.argsort(-scores64, axis=1, kind="stable")
    ss = np.take_along_axis(scores64, order, axis=1)
    scale = np.maximum(np.abs(ss).max(axis=1, keepdims=True), 1e-300)
    g = np.abs(np.diff(ss, axis=1)) / scale
    gn = np.full((B, C), np.inf)
    gp = np.full((B, C), np.inf)
    gn[:, : C - 1] = g
    gp[:, 1:] = g
    unamb = (gn > rel_margin) & (gp > rel_margin)
    eq = got_pred == order
    return {
        "rows": float(B),
        "top1_mismatch_unambiguous": float((~eq[:, 0] & unamb[:, 0]).sum()),
        "top1_mismatch_total": float((~eq[:, 0]).sum()),
        "rank_mismatch_unambiguous": float((~eq & unamb).sum()),
        "rank_mismatch_total": float((~eq).sum()),
    }


# --------------------------------------------------------------------------- §8(f) rows
def _sqnorm_f64_lane_order(x32: np.ndarray) -> np.ndarray:
    """sum_d x^2 per row in the accumulation order of csrc/prepare.cu (normalize_rows_kernel /
    row_sqnorm_kernel): lane l adds columns l, l+32, ... sequentially in fp64, then an xor
    butterfly (16, 8, 4, 2, 1) combines the 32 partials."""
    x = np.asarray(x32, dtype=np.float32).astype(np.float64)
    n, d = x.shape
    part = np.zeros((n, 32), dtype=np.float64)
    for d0 in range(0, d, 32):
        blk = x[:, d0:d0 + 32]
        part[:, : blk.shape[1]] += blk * blk  # one rounding per add; the square is exact
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        part = part + part[:, lanes ^ o]
    return part[:, 0]


def normalize_rows_ref(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """F.normalize(x, dim=1) (reference src/ssl_wafermap/models/knn.py:77, :90) with the fixed
    arithmetic of b200knn_normalize_rows: y = x / max(float32(sqrt(sum64 x^2)), eps) in fp32."""
    x32 = np.asarray(x).astype(np.float32)
    norm = np.sqrt(_sqnorm_f64_lane_order(x32)).astype(np.float32)
    denom = np.maximum(norm, np.float32(eps)).astype(np.float32)
    return (x32 / denom[:, None]).astype(np.float32)


def row_sqnorms_ref(x: np.ndarray) -> np.ndarray:
    return _sqnorm_f64_lane_order(np.asarray(x).astype(np.float32)).astype(np.float32)


def l2_augment(data_rows: np.ndarray, queries: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """x' = [x, ||x||^2], q' = [2q, -1] (fp32): argmax q'.x' = argmin ||x - q||, the search of
    notebooks/2.0-Figures-nearest-neighbors.ipynb:54.  Returns (queries' (B,D+1), bank' (D+1,N))."""
    d32 = np.asarray(data_rows).astype(np.float32)
    q32 = np.asarray(queries).astype(np.float32)
    bank = np.concatenate([d32, row_sqnorms_ref(d32)[:, None]], axis=1)
    qa = np.concatenate([q32 * np.float32(2.0), np.full((q32.shape[0], 1), -1.0, np.float32)], axis=1)
    return qa, np.ascontiguousarray(bank.T)


def l2_topk_o64(data_rows: np.ndarray, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """fp64 L2 distances, ties by lowest row index: what the notebooks' argsort means."""
    d = np.asarray(data_rows, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    dist = np.sqrt(np.maximum(((q * q).sum(1)[:, None] - 2.0 * q @ d.T + (d * d).sum(1)[None, :]), 0.0))
    idx = np.argsort(dist, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(dist, idx, axis=1), idx


def metrics_ref(pred: np.ndarray, target: np.ndarray, num_classes: int) -> Dict[str, np.ndarray]:
    """Confusion counts, macro accuracy / F1 and the row-normalised matrix of
    src/ssl_wafermap/models/knn.py:104-129 (torchmetrics semantics restated, see b200knn/metrics.py)."""
    C = int(num_classes)
    counts = np.zeros((C, C), dtype=np.int64)
    np.add.at(counts, (np.asarray(target), np.asarray(pred)), 1)
    c = counts.astype(np.float64)
    tp, support, predicted = np.diag(c), c.sum(1), c.sum(0)
    present = (support + predicted) > 0
    recall = np.divide(tp, support, out=np.zeros(C), where=support > 0)
    f1 = np.divide(2 * tp, support + predicted, out=np.zeros(C), where=(support + predicted) > 0)
    n = max(1, int(present.sum()))
    conf = np.divide(c, support[:, None], out=np.zeros_like(c), where=support[:, None] > 0)
    return {"counts": counts, "accuracy": float((recall * present).sum() / n),
            "f1": float((f1 * present).sum() / n), "confusion": conf}
