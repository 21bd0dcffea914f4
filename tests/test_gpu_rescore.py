"""GPU: b200knn_rescore through the C ABI, both implementations (TMA-pipelined with a workspace,
block-per-query without) against the sequential-fma oracle (oracle/seqfma.c) on hand-made
candidate lists: ragged dims, unaligned query rows, empty slots, k_in not a multiple of 32."""
import numpy as np
import pytest
import torch

from b200knn import _lib
from b200knn import knn as K
from oracle import knn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rescore(q, rows_pad, cand, N, D, k_out, idx_offset, err_coef, max_norm, use_ws):
    lib = _lib.load()
    B, k_in = cand.shape
    out = torch.zeros((B, k_out), dtype=torch.int64, device=DEV)
    flags = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    n_bad = torch.zeros((1,), dtype=torch.int32, device=DEV)
    ws_bytes = int(lib.b200knn_rescore_workspace_bytes(B, k_in)) if use_ws else 0
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=DEV)
    _lib.check(lib.b200knn_rescore(q.data_ptr(), K._DTYPES[q.dtype], q.stride(0), rows_pad.data_ptr(), None, N, D, cand.data_ptr(), B, k_in, k_out,
                                   idx_offset, float(err_coef), 0.0, 0.0, max_norm.data_ptr(), out.data_ptr(),
                                   flags.data_ptr(), n_bad.data_ptr(), ws.data_ptr() if use_ws else None, ws_bytes,
                                   torch.cuda.current_stream().cuda_stream), "rescore")
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64), flags.cpu().numpy(), int(n_bad.item())


@pytest.mark.parametrize("B,N,D,k_in,k_out,q_pad", [
    (37, 5000, 512, 216, 200, 0),     # the north-star shape of one call
    (5, 900, 72, 40, 20, 0),          # D padded to 128, k_in not a multiple of 32
    (130, 3000, 200, 64, 64, 3),      # query rows not 16-byte aligned (ld = D + 3)
    (9, 2000, 768, 300, 200, 0),      # config c5 dimension: 6 chunks of 128 columns
    (3, 400, 1, 33, 10, 0),           # D = 1
    (1, 50, 384, 50, 50, 0),          # every bank row is a candidate (k_in == N)
    (700, 20000, 64, 17, 5, 0),       # one chunk per unit, many units per warp
    (60, 4000, 512, 240, 240, -1),    # routed lists of the sharded mode: 0..40 candidates, whole empty units
])
def test_rescore_both_kernels_match_seqfma(B, N, D, k_in, k_out, q_pad):
    sparse, q_pad = q_pad < 0, max(q_pad, 0)
    rng = np.random.default_rng(B * 7 + D)
    bank = rng.standard_normal((N, D)).astype(np.float32)
    qfull = rng.standard_normal((B, D + q_pad)).astype(np.float32)
    q = qfull[:, :D]
    d_pad = K.padded_dim(D)
    rows = np.zeros((N, d_pad), dtype=np.float32)
    rows[:, :D] = bank
    idx_offset = 1000
    # candidates: distinct random rows per query, random approximate sims (sorted desc), and
    # for every third query a ragged tail of empty slots
    cand = np.zeros((B, k_in), dtype=np.uint64)
    idx = np.zeros((B, k_in), dtype=np.int64)
    for b in range(B):
        idx[b] = rng.choice(N, size=k_in, replace=False)
    approx = -np.sort(-rng.standard_normal((B, k_in)).astype(np.float32), axis=1)
    cand[:] = O.make_keys(approx, idx + idx_offset)
    n_valid = np.full(B, k_in)
    for b in range(0, B, 3):
        n_valid[b] = max(k_out, k_in - 1 - (b % 7)) if k_in > k_out else k_in
        cand[b, n_valid[b]:] = 0
    if sparse:
        n_valid = rng.integers(0, 41, size=B)
        n_valid[:3] = 0
        for b in range(B):
            cand[b, n_valid[b]:] = 0

    sims = O.sims_seqfma(np.ascontiguousarray(q), np.ascontiguousarray(bank.T))  # (B, N) sequential fma
    want = np.zeros((B, k_out), dtype=np.uint64)
    for b in range(B):
        keys = O.make_keys(sims[b, idx[b, :n_valid[b]]][None], (idx[b, :n_valid[b]] + idx_offset)[None])[0]
        keys = np.sort(keys)[::-1]
        want[b, :min(k_out, keys.size)] = keys[:k_out]

    tq = torch.from_numpy(qfull).to(DEV)[:, :D]
    trows = torch.from_numpy(rows).to(DEV)
    tcand = torch.from_numpy(cand.view(np.int64)).to(DEV)
    max_norm = torch.tensor([float(np.linalg.norm(bank, axis=1).max()) * 1.001], dtype=torch.float32, device=DEV)
    res = {}
    for use_ws in (True, False):
        got, flags, n_bad = _rescore(tq, trows, tcand, N, D, k_out, idx_offset, 0.0, max_norm, use_ws)
        assert np.array_equal(got, want), f"workspace={use_ws}: keys differ from the sequential-fma oracle"
        assert n_bad == int(flags.sum()) and set(np.unique(flags)) <= {0, 1}
        res[use_ws] = flags
    assert np.array_equal(res[True], res[False])
    # rows with an empty last slot are uncertified unless every bank row was a candidate
    starved = (cand[:, -1] == 0)
    if k_in < N:
        assert bool((res[True][starved] == 1).all())
    else:
        assert bool((res[True][starved] == 0).all())


def test_rescore_certificate_terms():
    """err_coef / err_abs / max_abs: E = err_coef*||q||*M + err_abs*(||q|| + M); max_abs refuses."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    N, D, B, k_in, k_out = 300, 64, 4, 40, 32
    bank = rng.standard_normal((N, D)).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    q = rng.standard_normal((B, D)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    sims = O.sims_seqfma(q, np.ascontiguousarray(bank.T))
    order = np.argsort(-sims, axis=1, kind="stable")[:, :k_in]
    cand = O.make_keys(np.take_along_axis(sims, order, 1), order)
    gap = np.take_along_axis(sims, order, 1)[:, k_out - 1] - np.take_along_axis(sims, order, 1)[:, k_in - 1]
    tq, trows = torch.from_numpy(q).to(DEV), torch.from_numpy(bank).to(DEV)
    tcand = torch.from_numpy(cand.view(np.int64)).to(DEV)
    one = torch.tensor([1.0], dtype=torch.float32, device=DEV)

    def run(err_coef, err_abs, max_abs):
        out = torch.zeros((B, k_out), dtype=torch.int64, device=DEV)
        flags = torch.zeros((B,), dtype=torch.int32, device=DEV)
        n_bad = torch.zeros((1,), dtype=torch.int32, device=DEV)
        ws_bytes = int(lib.b200knn_rescore_workspace_bytes(B, k_in))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=DEV)
        _lib.check(lib.b200knn_rescore(tq.data_ptr(), 0, D, trows.data_ptr(), None, N, D, tcand.data_ptr(), B, k_in,
                                       k_out, 0, err_coef, err_abs, max_abs, one.data_ptr(), out.data_ptr(),
                                       flags.data_ptr(), n_bad.data_ptr(), ws.data_ptr(), ws_bytes,
                                       torch.cuda.current_stream().cuda_stream), "rescore")
        return flags.cpu().numpy()

    assert not run(0.0, 0.0, 0.0).any()
    e = float(gap.min()) * 0.5
    assert not run(e / 1.002, 0.0, 0.0).any()            # E just below the smallest gap
    assert run(float(gap.max()) * 1.1, 0.0, 0.0).all()   # E above every gap
    assert run(0.0, float(gap.max()) * 0.6, 0.0).all()   # err_abs * (1 + 1) above every gap
    assert run(0.0, 0.0, 0.5).all()                      # ||q|| = 1 >= max_abs
    assert not run(0.0, 0.0, 2.0).any()


def test_route_keys_and_certify_match_numpy():
    """The two small kernels of the sharded fp32 mode against their numpy statements."""
    rng = np.random.default_rng(11)
    n, k_in, k, G, rows_per_shard, D = 37, 48, 20, 3, 1000, 72
    idx = rng.integers(0, G * rows_per_shard, size=(n, k_in))
    sims = -np.sort(-rng.standard_normal((n, k_in)).astype(np.float32), axis=1)
    keys = O.make_keys(sims, idx)
    keys[::4, -5:] = 0  # ragged tails
    out = K.route_keys(torch.from_numpy(keys.view(np.int64)).to(DEV), rows_per_shard, G).cpu().numpy().view(np.uint64)
    for g in range(G):
        want = np.zeros_like(keys)
        for r in range(n):
            mine = keys[r][(keys[r] != 0) & (idx[r] // rows_per_shard == g)]
            want[r, :mine.size] = mine  # original order, compacted to the front
        assert np.array_equal(out[g], want)

    q = rng.standard_normal((n, D)).astype(np.float32)
    exact = O.make_keys(sims[:, :k] + np.float32(0.01), idx[:, :k])
    level = dict(err_coef=1e-3, err_abs=1e-6, max_abs=0.0)
    m = 2.5
    flags = K.certify(torch.from_numpy(exact.view(np.int64)).to(DEV), torch.from_numpy(keys.view(np.int64)).to(DEV),
                      torch.from_numpy(q).to(DEV), level, torch.tensor([m], dtype=torch.float32, device=DEV),
                      all_rows=False).cpu().numpy()
    qn = np.linalg.norm(q.astype(np.float64), axis=1) * 1.001
    e = level["err_coef"] * qn * m + level["err_abs"] * np.sqrt(K.padded_dim(D)) * (qn + m)
    es, _ = O.decode_keys(exact)
    as_, _ = O.decode_keys(keys)
    margin = es[:, -1].astype(np.float64) - (as_[:, -1].astype(np.float64) + e)
    want = np.where(keys[:, -1] == 0, 1, (margin <= 0).astype(np.int32))
    decided = (keys[:, -1] == 0) | (np.abs(margin) > 1e-5)  # away from fp32 rounding of the bound
    assert np.array_equal(flags[decided], want[decided])
    assert flags[::4].all()  # empty k_in-th slot with k_in < N: never certified
