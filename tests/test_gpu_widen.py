"""GPU: the SURVEY.md §8(f) rows through the C ABI — fused normalise + bank build (f1), L2
retrieval over embedding tables (f2 / a5), on-device metrics (f3), kNN-graph export (f4)."""
import os

import numpy as np
import pytest
import torch

import b200knn
import datagen
from b200knn import knn as K
from oracle import knn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
@pytest.mark.parametrize("d", [72, 384, 512])
def test_normalize_rows_bit_exact_with_oracle(d, dtype):
    rng = np.random.default_rng(d)
    x = (rng.standard_normal((300, d)) * rng.uniform(0.05, 60, (300, 1))).astype(dtype)
    x[11] = 0
    got = b200knn.normalize_rows(_t(x))
    assert got.shape == (300, d) and got._base.shape[1] == K.padded_dim(d)
    assert bool((got._base[:, d:] == 0).all())  # pad columns are zero
    want = O.normalize_rows_ref(x)
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    ref = torch.nn.functional.normalize(_t(x).float(), dim=1)
    assert float((got - ref).abs().max()) <= 2.5e-7
    assert np.array_equal(b200knn.row_sqnorms(_t(x)).cpu().numpy().view(np.uint32), O.row_sqnorms_ref(x).view(np.uint32))


@pytest.mark.parametrize("mode", ["exact", "fp32", "bf16"])
def test_feature_bank_equals_reference_flow(mode):
    """FeatureBank (fused normalise + layout) must give bit for bit what the reference flow gives
    on the same normalised values: F.normalize -> cat -> .t().contiguous() -> knn_predict."""
    rng = np.random.default_rng(21)
    raw_bank = (rng.standard_normal((5000, 512)) * 3).astype(np.float32)
    raw_q = (rng.standard_normal((70, 512)) * 3).astype(np.float32)
    lab = datagen.labels(5000, 9, 5)
    fb = b200knn.FeatureBank.from_rows(_t(raw_bank), _t(lab), normalize=True)
    b200knn.set_default_mode(mode)
    try:
        got = fb.knn_predict(_t(raw_q), 9, 200, 0.1, normalize=True)
        # reference-shaped call on tensors holding the SAME normalised values
        bank_dn = _t(O.normalize_rows_ref(raw_bank)).t().contiguous()      # knn.py:77,80
        q = _t(O.normalize_rows_ref(raw_q))                                 # knn.py:90
        want = b200knn.knn_predict(q, bank_dn, _t(lab), 9, 200, 0.1)        # knn.py:91-98
        assert torch.equal(got, want)
        # the zero-copy (D,N) view also serves the reference's own call signature
        assert torch.equal(b200knn.knn_predict(q, fb.bank, fb.labels, 9, 200, 0.1), want)
        if mode != "bf16":
            ss, si = O.topk_seqfma(O.normalize_rows_ref(raw_q), np.ascontiguousarray(O.normalize_rows_ref(raw_bank).T), 200)
            assert np.array_equal(got.cpu().numpy(), O.vote_o64(ss, si, lab, 9, 0.1)[0])
    finally:
        b200knn.set_default_mode("exact")
    # batch-wise construction (the loop of on_validation_epoch_start) is the same bank
    chunks = [(_t(raw_bank[i:i + 640]), _t(lab[i:i + 640])) for i in range(0, 5000, 640)]
    fb2 = b200knn.FeatureBank.from_batches(chunks, normalize=True)
    assert torch.equal(fb2.rows, fb.rows) and torch.equal(fb2.labels, fb.labels)


def test_feature_bank_shadow_rows_are_not_copied():
    x = torch.nn.functional.normalize(torch.randn(3000, 512, device=DEV), dim=1)
    fb = b200knn.FeatureBank.from_rows(x, normalize=False)
    pb = K.bank_cache.get(fb.bank, "bf16")
    rows_a, rows_b = pb.rescore_rows()
    assert rows_a.data_ptr() == fb.rows.data_ptr() and rows_b is None


@pytest.mark.parametrize("model", ["FastSiam", "SimSiam"])
def test_l2_search_on_reference_tables(model, golden_dir):
    """The notebooks' search (2.0-Figures-nearest-neighbors.ipynb:54) on the reference's shipped
    embedding rows: exact mode bit-equal to the sequential-fma oracle on the augmented vectors,
    and equal to the fp64 L2 ranking wherever that ranking is unambiguous in fp32."""
    g = np.load(os.path.join(golden_dir, f"real_{model}.npz"))
    data = g["bank_rows_f16"]              # raw fp16 rows as stored in *_preds_subset.pkl.xz
    queries = data[:40]                    # the notebooks query with rows of the table itself
    k = 6
    index = b200knn.L2Index(_t(data))
    dist, idx = index.search(_t(queries), k, mode="exact")
    qa, bank = O.l2_augment(data, queries)
    ss, si = O.topk_seqfma(qa, bank, k)
    assert np.array_equal(idx.cpu().numpy(), si)
    d64, i64 = O.l2_topk_o64(data, queries, k)
    scale = float((data.astype(np.float64) ** 2).sum(1).max())
    gaps_ok = np.ones_like(i64, dtype=bool)
    full = np.sort(np.sqrt(np.maximum(((queries.astype(np.float64) ** 2).sum(1)[:, None]
                                       - 2.0 * queries.astype(np.float64) @ data.astype(np.float64).T
                                       + (data.astype(np.float64) ** 2).sum(1)[None, :]), 0)) ** 2, axis=1)[:, :k + 1]
    eps = 4e-7 * scale
    for j in range(k):
        gaps_ok[:, j] = (np.abs(full[:, j + 1] - full[:, j]) > eps) & ((j == 0) | (np.abs(full[:, j] - full[:, j - 1]) > eps))
    assert np.array_equal(idx.cpu().numpy()[gaps_ok], i64[gaps_ok])
    # rank 0 is the query row itself (or an exact duplicate with a lower index), distance ~0
    assert float(dist[:, 0].max()) <= 1e-2 * np.sqrt(scale)
    # the default (bit-exact tensor-core) mode returns the same neighbours
    d2, i2 = index.search(_t(queries), k, mode="fp32")
    assert torch.equal(i2, idx) and torch.equal(d2, dist)


def test_knn_graph_matches_topk_and_drops_self():
    x = torch.nn.functional.normalize(torch.randn(3000, 384, device=DEV), dim=1)
    sims, idx = b200knn.knn_graph(x, 15, normalize=False, mode="exact", batch=1024)
    assert sims.shape == (3000, 15) and idx.shape == (3000, 15)
    own = torch.arange(3000, device=DEV).view(-1, 1)
    assert not bool((idx == own).any())
    s_all, i_all = b200knn.knn_topk(x, x.t().contiguous(), 16, mode="exact")
    assert torch.equal(i_all[:, 0], own.view(-1))          # distinct rows: a row is its own best match
    assert torch.equal(idx, i_all[:, 1:]) and torch.equal(sims, s_all[:, 1:])
    s2, i2 = b200knn.knn_graph(x, 15, normalize=False, mode="exact", include_self=True)
    assert torch.equal(i2, i_all[:, :15])


@pytest.mark.parametrize("C", [9, 38])
def test_metrics_against_oracle_and_sklearn(C):
    from sklearn.metrics import f1_score, recall_score

    rng = np.random.default_rng(C)
    target = rng.integers(0, C - 1, 50000)        # the last class never occurs
    pred = np.where(rng.random(50000) < 0.6, target, rng.integers(0, C - 1, 50000))
    m = b200knn.knn_metrics(_t(pred), _t(target), C)
    ref = O.metrics_ref(pred, target, C)
    assert np.array_equal(m["counts"].cpu().numpy(), ref["counts"])
    assert abs(float(m["accuracy"]) - recall_score(target, pred, average="macro")) < 1e-12
    assert abs(float(m["f1"]) - f1_score(target, pred, average="macro")) < 1e-12
    assert np.allclose(m["confusion"].cpu().numpy(), ref["confusion"])
    with pytest.raises(RuntimeError, match="outside"):
        b200knn.confusion_counts(_t(np.array([C])), _t(np.array([0])), C)
