"""CPU: the oracle restatement against the committed golden vectors, and the
internal consistency of its three levels (R32 / O64 / SEQ) under the
ambiguity-aware contract of SURVEY.md §7.3."""
import os

import numpy as np
import pytest
import torch

import datagen
from oracle import knn_oracle as O


@pytest.fixture(scope="module")
def synth(golden_dir):
    both = dict(np.load(os.path.join(golden_dir, "synth.npz")))
    both.update(np.load(os.path.join(golden_dir, "synth_nonneg.npz")))  # non-negative cases (round 2)
    return both


@pytest.mark.parametrize("name", datagen.CASE_NAMES + datagen.RELU_CASE_NAMES)
def test_seq_and_o64_match_golden(name, synth):
    c = datagen.make_case(name)
    ss, si = O.topk_seqfma(c["feature"], c["bank"], c["k"])
    assert np.array_equal(si, synth[name + "_seq_idx"].astype(np.int64))
    assert np.array_equal(ss.view(np.uint32), synth[name + "_seq_sims"].view(np.uint32))  # bit-exact
    s64, i64 = O.topk_o64(c["feature"], c["bank"], min(c["k"] + 1, c["bank"].shape[1]))
    assert np.array_equal(i64, synth[name + "_o64_idx"].astype(np.int64))
    p64, _ = O.vote_o64(s64[:, : c["k"]], i64[:, : c["k"]], c["labels"], c["C"], c["t"])
    assert np.array_equal(p64, synth[name + "_o64_pred"].astype(np.int64))


@pytest.mark.parametrize("name", datagen.CASE_NAMES + datagen.RELU_CASE_NAMES)
def test_r32_matches_golden_under_contract(name, synth):
    """torch CPU fp32 (the reference algorithm verbatim) is not bitwise stable across thread
    counts / MKL versions, so it is pinned through the contract, not bitwise."""
    c = datagen.make_case(name)
    pred, sims, idx, _ = O.knn_predict_r32_full(torch.from_numpy(c["feature"]), torch.from_numpy(c["bank"]),
                                                torch.from_numpy(c["labels"]), c["C"], c["k"], c["t"])
    r = O.compare_topk(sims.numpy(), idx.numpy(), c["feature"], c["bank"], c["k"])
    assert r["idx_mismatch_unambiguous"] == 0 and r["set_mismatch_rows_unambiguous"] == 0
    assert r["max_rel_err"] <= 1e-5
    rp = O.compare_pred(pred.numpy(), synth[name + "_o64_scores"])
    assert rp["top1_mismatch_unambiguous"] == 0
    # and the stored R32 top-1 agrees wherever unambiguous
    rp2 = O.compare_pred(synth[name + "_r32_pred"].astype(np.int64), synth[name + "_o64_scores"])
    assert rp2["top1_mismatch_unambiguous"] == 0


@pytest.mark.parametrize("model", ["FastSiam", "SimSiam"])
@pytest.mark.parametrize("tag,k", [("norm_", 5), ("norm_", 200), ("raw_", 5), ("raw_", 200)])
def test_real_banks_golden(model, tag, k, golden_dir):
    """Inputs from the reference's shipped embedding banks (duplicate rows, 73-87 % exact zeros,
    row norms up to 277): the tie / un-normalised stress set."""
    g = np.load(os.path.join(golden_dir, f"real_{model}.npz"))
    b = g["bank_rows_f16"].astype(np.float32)
    q = g["query_rows_f16"].astype(np.float32)
    if tag == "norm_":
        b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-12)
        q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    bank = np.ascontiguousarray(b.T)
    ss, si = O.topk_seqfma(q, bank, k)
    assert np.array_equal(si, g[f"{tag}k{k}_seq_idx"].astype(np.int64))
    assert np.array_equal(ss.view(np.uint32), g[f"{tag}k{k}_seq_sims"].view(np.uint32))
    r = O.compare_topk(ss, si, q, bank, k)
    assert r["idx_mismatch_unambiguous"] == 0 and r["set_mismatch_rows_unambiguous"] == 0
    assert r["max_rel_err"] <= 1e-5


def test_keys_roundtrip_and_order():
    rng = np.random.default_rng(3)
    s = rng.standard_normal((5, 300)).astype(np.float32)
    s[0, :10] = 0.0
    s[0, 3] = -0.0
    s[1, :] = 1.5  # all ties -> index order
    s[2, 7] = np.inf
    s[2, 8] = -np.inf
    idx = np.tile(np.arange(300), (5, 1))
    keys = O.make_keys(s, idx)
    s2, i2 = O.decode_keys(keys)
    assert np.array_equal(i2, idx)
    assert np.array_equal((s + np.float32(0)).view(np.uint32), s2.view(np.uint32))
    order = np.argsort(-(keys.astype(np.float128)), axis=1, kind="stable")
    cs, ci = O.canonical_topk_np(s.astype(np.float64), 300)
    assert np.array_equal(order, ci)
    # empty key decodes to (-inf, -1)
    es, ei = O.decode_keys(np.zeros((1, 2), dtype=np.uint64))
    assert np.all(np.isneginf(es)) and np.all(ei == -1)


def test_canonical_topk_c_matches_numpy():
    rng = np.random.default_rng(5)
    s = rng.integers(-3, 4, size=(9, 257)).astype(np.float32)  # heavy ties
    a_s, a_i = O.canonical_topk_c(s, 50)
    b_s, b_i = O.canonical_topk_np(s.astype(np.float64), 50)
    assert np.array_equal(a_i, b_i) and np.array_equal(a_s, b_s.astype(np.float32))
    # index offset (bank row-sharding)
    c_s, c_i = O.canonical_topk_c(s, 50, idx_offset=1000)
    assert np.array_equal(c_i, b_i + 1000)


def test_merge_is_partition_invariant():
    """Merging per-shard top-k lists == top-k of the whole bank, for any partition (§8e)."""
    c = datagen.make_case("clustered_small")
    sims = O.sims_seqfma(c["feature"], c["bank"])
    k = 64
    full_s, full_i = O.canonical_topk_c(sims, k)
    full_keys = O.make_keys(full_s, full_i)
    N = sims.shape[1]
    for G in (2, 3, 8):
        per = -(-N // G)
        parts = []
        for g in range(G):
            lo, hi = min(N, g * per), min(N, (g + 1) * per)
            s_, i_ = O.canonical_topk_c(np.ascontiguousarray(sims[:, lo:hi]), min(k, hi - lo), idx_offset=lo)
            kk = np.zeros((sims.shape[0], k), dtype=np.uint64)
            kk[:, : s_.shape[1]] = O.make_keys(s_, i_)
            parts.append(kk)
        merged = O.merge_keys_np(np.stack(parts), k)
        assert np.array_equal(merged, full_keys)


def test_vote_ties_and_errors():
    sims = np.array([[0.5, 0.5, 0.25]], dtype=np.float32)
    idx = np.array([[0, 1, 2]])
    labels = np.array([3, 1, 1])
    pred, sc = O.vote_o64(sims, idx, labels, 5, 0.1)
    # class 1 has e^5 + e^2.5 > class 3 e^5 ; zero-score classes follow in ascending id
    assert pred.tolist() == [[1, 3, 0, 2, 4]]
    with pytest.raises(RuntimeError):
        O.vote_o64(sims, idx, np.array([0, 9, 1]), 5, 0.1)
