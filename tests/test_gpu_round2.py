"""GPU (through the C ABI): round-2 parity cases.
 * non-negative (post-ReLU-like) embeddings: the bit-exact mode on the distribution the reference
   really produces (timm ResNet-18 pooled features, knn.py:322, normalised at :77/:90);
 * the tensor core's accumulation error against the D-scaled allowance of the certificate, on
   all-positive and adversarial operands;
 * k > 992 (Tensor.topk takes any k <= N), k = 992 under tie floods (list-overflow regression);
 * wide vectors (D = 1024) through every entry point that used to bypass the mode fallback;
 * the CUDA-graph path of the reference-shaped call (B = 64);
 * the replacement validation hooks (install(hooks=True)) against the un-hooked flow."""
import numpy as np
import pytest
import torch

import b200knn
import datagen
from b200knn import knn as K
from oracle import knn_oracle as O
from test_gpu_tc import run_dump

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


# ------------------------------------------------------------------ non-negative embeddings
@pytest.mark.parametrize("mode", ["fp32", "fp32_f16", "fp32_f16x2", "fp32_bf16x3", "fp32_tf32", "exact"])
@pytest.mark.parametrize("name", datagen.RELU_CASE_NAMES)
def test_nonnegative_embeddings_bitwise_equal_oracle(name, mode):
    c = datagen.make_case(name)
    assert (c["feature"] >= 0).all() and (c["bank"] >= 0).all()
    f, bank, lab = _t(c["feature"]), _t(c["bank"]), _t(c["labels"])
    ss, si = O.topk_seqfma(c["feature"], c["bank"], c["k"])
    keys = b200knn.topk_keys(f, bank, c["k"], mode=mode)
    sims, idx = b200knn.decode_keys(keys)
    assert np.array_equal(idx.cpu().numpy(), si)
    assert np.array_equal(sims.cpu().numpy().view(np.uint32), ss.view(np.uint32))
    b200knn.set_default_mode(mode)
    try:
        pred = b200knn.knn_predict(f, bank, lab, c["C"], c["k"], c["t"]).cpu().numpy()
    finally:
        b200knn.set_default_mode("exact")
    assert np.array_equal(pred, O.vote_o64(ss, si, c["labels"], c["C"], c["t"])[0])
    if mode in K.RESCORED_MODES:
        print(f"{mode} {name}: cascade {K.last_rescore_stats['levels']}")


def _device_rows(kind, n, d, seed, C=9):
    g = torch.Generator(device=DEV).manual_seed(seed)
    cent = torch.nn.functional.normalize(torch.randn(C, d, generator=g, device=DEV), dim=1)
    out = torch.empty(n, d, device=DEV)
    for lo in range(0, n, 131072):
        hi = min(n, lo + 131072)
        z = torch.randn(hi - lo, d, generator=g, device=DEV)
        if kind == "absgauss":
            x = z.abs()
        else:
            lab = torch.randint(0, C, (hi - lo,), generator=g, device=DEV)
            x = (cent[lab] + 1.4 * z / d ** 0.5).clamp_min(0.0)
        out[lo:hi] = torch.nn.functional.normalize(x, dim=1)
    return out


@pytest.mark.parametrize("kind", ["relu", "absgauss"])
def test_large_bank_nonnegative_against_exact_mode(kind):
    """811,457 x 512 non-negative rows, k = 200: similarities are bunched (every pair of rows is
    similar) and all products have one sign.  The bit-exact mode must equal the on-device exact
    mode bitwise; the rows each cascade level leaves uncertified are reported."""
    N, D, k, B = 811457, 512, 200, 512
    bank = _device_rows(kind, N, D, 1).t().contiguous()
    q = _device_rows(kind, B, D, 2)
    ek = b200knn.topk_keys(q, bank, k, mode="exact")
    for mode in ("fp32", "fp32_f16x2"):
        tk = b200knn.topk_keys(q, bank, k, mode=mode)
        print(f"{kind} N={N} {mode}: cascade (level, rows in, left uncertified) {K.last_rescore_stats['levels']}")
        assert torch.equal(ek, tk)
    es, _ = b200knn.decode_keys(ek)
    print(f"   exact top-{k} similarities: best {float(es[:, 0].mean()):.4f}, k-th {float(es[:, -1].mean()):.4f}, "
          f"mean gap between ranks {float((es[:, 0] - es[:, -1]).mean()) / (k - 1):.2e}")


# ------------------------------------------------------------------ accumulation error of the tensor core
def _adversarial(D, pattern, rng):
    """fp16-exact operand rows that maximise what a truncating accumulator can lose."""
    q = np.zeros((8, D), dtype=np.float32)
    x = np.zeros((256, D), dtype=np.float32)
    if pattern == "big_first_half_ulp":      # 1 + (D-1) products of 2^-24 = half a unit in the last place of 1
        q[:], x[:] = 2.0 ** -12, 2.0 ** -12
        q[:, 0], x[:, 0] = 1.0, 1.0
    elif pattern == "big_first_just_under_2ulp":   # products of 1.998 x 2^-23: 0.998 ulp lost if truncated to the ulp grid
        q[:], x[:] = 2.0 ** -11, np.float32(np.float16(2.0 ** -12 * (2 - 2.0 ** -9)))
        q[:, 0], x[:, 0] = 1.0, 1.0
    elif pattern == "big_last":              # the large product arrives after D-1 small ones
        q[:], x[:] = 2.0 ** -12, 2.0 ** -12
        q[:, -1], x[:, -1] = 1.0, 1.0
    elif pattern == "ramp":                  # magnitudes spread over 10 binades, all positive
        e = rng.integers(-10, 1, size=D)
        q[:] = (2.0 ** e * rng.uniform(1, 2, D)).astype(np.float16).astype(np.float32)
        x[:] = rng.uniform(0.5, 1, (256, D)).astype(np.float16).astype(np.float32)
    return q, np.ascontiguousarray(x.T)


@pytest.mark.parametrize("D", [512, 768, 1024])
@pytest.mark.parametrize("mode", ["f16", "f16x2", "bf16", "bf16x3", "tf32x3"])
def test_tmem_accumulation_error_bound(D, mode):
    """The certificate allows acc_c * D_pad * 2^-23 * ||q|| ||x|| for the accumulation in the tensor
    core (b200knn/knn.py LEVELS).  Measured here on the raw TMEM tiles against an exact (fp64)
    evaluation of the SAME prepared operands — so operand rounding is excluded and only the
    accumulation is seen — for all-positive random rows and adversarial magnitude patterns."""
    if mode in ("f16", "f16x2", "bf16") and K.padded_dim(D) > K.MAX_BF16_DIM:
        pytest.skip("resident-query modes stop at D_pad = 768 (wider vectors run as bf16x3)")
    level = {"f16": "fp32_f16", "f16x2": "fp32_f16x2", "bf16": "fp32_bf16", "bf16x3": "fp32_bf16x3",
             "tf32x3": "fp32_tf32"}[mode]
    allow = K.LEVELS[level]["acc_c"] * K.padded_dim(D) * 2.0 ** -23
    rng = np.random.default_rng(D)
    worst = {}
    cases = [("absgauss", datagen.absgauss(16, D, 1), np.ascontiguousarray(datagen.absgauss(512, D, 2).T))]
    cases += [(p,) + _adversarial(D, p, rng) for p in ("big_first_half_ulp", "big_first_just_under_2ulp", "big_last", "ramp")]
    for name, q, bank in cases:
        _, dump, pq, pb = run_dump(q, bank, 5, mode)
        qh = pq.hi.double().cpu().numpy()[:, :D]
        bh = pb.hi.double().cpu().numpy()[:, :D]
        if mode in ("f16", "bf16"):
            ref = qh @ bh.T
        elif mode == "f16x2":
            ref = qh @ (bh + pb.lo.double().cpu().numpy()[:, :D]).T
        else:  # the three products the split kernels issue: hi*lo + lo*hi + hi*hi
            ql, bl = pq.lo.double().cpu().numpy()[:, :D], pb.lo.double().cpu().numpy()[:, :D]
            ref = qh @ bl.T + ql @ bh.T + qh @ bh.T
        scale = np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(bank.astype(np.float64), axis=0)[None, :]
        worst[name] = float((np.abs(dump.astype(np.float64) - ref) / scale).max())
    print(f"{mode} D={D}: accumulation error / (|q||x|) {({k_: f'{v:.2e}' for k_, v in worst.items()})} "
          f"allowance {allow:.2e} (worst uses {max(worst.values()) / allow:.1%} of it)")
    assert max(worst.values()) <= allow


# ------------------------------------------------------------------ k beyond the streaming lists
@pytest.mark.parametrize("mode", ["exact", "fp32", "bf16"])
@pytest.mark.parametrize("k", [993, 1000, 4096])
def test_k_above_992_matches_oracle(k, mode):
    """Tensor.topk accepts any k <= N (lightly's knn_predict passes knn_k straight through): larger
    k is served by peeling passes of the exact kernel, whatever the mode."""
    rng = np.random.default_rng(k)
    N, D, B = 6000, 64, 37
    bank = np.ascontiguousarray(rng.standard_normal((N, D)).astype(np.float32).T)
    bank[:, 100:140] = bank[:, 60:100]   # duplicates: ties across a pass boundary break by lowest index
    q = rng.standard_normal((B, D)).astype(np.float32)
    lab = datagen.labels(N, 9, 3)
    ss, si = O.topk_seqfma(q, bank, k)
    sims, idx = b200knn.knn_topk(_t(q), _t(bank), k, mode=mode)
    assert np.array_equal(idx.cpu().numpy(), si)
    assert np.array_equal(sims.cpu().numpy().view(np.uint32), ss.view(np.uint32))
    b200knn.set_default_mode(mode)
    try:
        pred = b200knn.knn_predict(_t(q), _t(bank), _t(lab), 9, k, 0.5).cpu().numpy()
    finally:
        b200knn.set_default_mode("exact")
    assert np.array_equal(pred, O.vote_o64(ss, si, lab, 9, 0.5)[0])


@pytest.mark.parametrize("mode", ["exact", "fp32", "bf16", "f16"])
@pytest.mark.parametrize("B", [3, 300])
def test_k_992_with_tie_floods(B, mode):
    """Largest list k with heavily duplicated bank rows, with (B = 3: the bank is split) and without
    bank splits: radix-select prunes meet ties at the cut in nearly every prune, which must never
    leave a list without room for the next 32 appends (round-1 advisor finding)."""
    rng = np.random.default_rng(B)
    D, k = 128, 992
    base = rng.standard_normal((60, D)).astype(np.float32)
    rows = np.repeat(base, 100, axis=0)                     # 6,000 rows, every row 100 times
    rows = rows[rng.permutation(rows.shape[0])]
    bank = np.ascontiguousarray(rows.T)
    q = rng.standard_normal((B, D)).astype(np.float32)
    ss, si = O.topk_seqfma(q, bank, k)
    sims, idx = b200knn.knn_topk(_t(q), _t(bank), k, mode=mode)
    if mode in ("exact", "fp32"):
        assert np.array_equal(idx.cpu().numpy(), si)
        assert np.array_equal(sims.cpu().numpy().view(np.uint32), ss.view(np.uint32))
    else:  # approximate similarities: every group of 100 duplicates still arrives whole and index-ordered
        i = idx.cpu().numpy()
        assert (i >= 0).all() and all(len(set(r.tolist())) == k for r in i)
        s = sims.cpu().numpy()
        assert (np.diff(s, axis=1) <= 0).all()
        same = np.diff(s, axis=1) == 0
        assert (np.diff(i, axis=1)[same] > 0).all()


# ------------------------------------------------------------------ wide vectors everywhere
@pytest.mark.parametrize("mode", ["bf16", "f16", "f16x2", "fp32"])
def test_wide_vectors_through_repair_and_sampling(mode, monkeypatch):
    """D = 1024: bf16 / f16 / f16x2 run as bf16x3.  The starved-row repair (recompute_rows), the
    sampling pre-pass and knn_predict must resolve the same effective mode (round-1 advisor
    finding: they used to pass the raw mode to the kernel and raise)."""
    D, N, B, k = 1024, 80000, 70, 200
    g = torch.Generator(device=DEV).manual_seed(4)
    bank = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=DEV), dim=1).t().contiguous()
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=DEV), dim=1)
    lab = torch.randint(0, 9, (N,), generator=g, device=DEV)
    assert K.prepass_stride(N, k) > 0
    ek = b200knn.topk_keys(q, bank, k, mode="exact")
    b200knn.set_default_mode("exact")
    want = b200knn.knn_predict(q, bank, lab, 9, k, 0.1)
    real_kth = K.kth_sim

    def starving_kth(keys):  # the sampled threshold of the first rows is too high: those rows starve
        t = real_kth(keys)
        if keys.shape[1] == K.PREPASS["r"]:
            t[:5] = 5.0
        return t

    monkeypatch.setattr(K, "kth_sim", starving_kth)
    monkeypatch.setitem(K.GRAPHS, "enabled", False)
    b200knn.set_default_mode(mode)
    try:
        pred = b200knn.knn_predict(q, bank, lab, 9, k, 0.1)
        tk = b200knn.topk_keys(q, bank, k, mode=mode)
    finally:
        b200knn.set_default_mode("exact")
    if mode == "fp32":
        assert torch.equal(tk, ek) and torch.equal(pred, want)
    else:
        assert bool((tk[:, -1] != 0).all())  # repaired
        ei, ti = b200knn.decode_keys(ek)[1].cpu().numpy(), b200knn.decode_keys(tk)[1].cpu().numpy()
        assert np.mean([len(set(a) & set(b)) / k for a, b in zip(ei, ti)]) >= 0.995
        assert float((pred[:, 0] == want[:, 0]).float().mean()) >= 0.97


# ------------------------------------------------------------------ the reference-shaped call through a CUDA graph
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("N,k", [(37348, 5), (120000, 200)])
def test_small_batch_graph_path_equals_eager(mode, N, k):
    D, B, C = 512, 64, 9
    g = torch.Generator(device=DEV).manual_seed(N + k)
    bank = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=DEV), dim=1).t().contiguous()
    lab = torch.randint(0, C, (N,), generator=g, device=DEV)
    qs = [torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=DEV), dim=1) for _ in range(4)]
    K.clear_call_graphs()
    b200knn.set_default_mode(mode)
    try:
        K.GRAPHS["enabled"] = False
        eager = [b200knn.knn_predict(x, bank, lab, C, k, 0.1) for x in qs]
        K.GRAPHS["enabled"] = True
        before = dict(K.graph_stats)
        graphed = [b200knn.knn_predict(x, bank, lab, C, k, 0.1) for x in qs]
        assert K.graph_stats["captures"] == before["captures"] + 1
        assert K.graph_stats["replays"] == before["replays"] + len(qs)
        for a, b in zip(eager, graphed):
            assert torch.equal(a, b)
        # results survive the next replay (the graph's output buffer is not handed out)
        again = b200knn.knn_predict(qs[0], bank, lab, C, k, 0.1)
        assert torch.equal(graphed[1], eager[1]) and torch.equal(again, eager[0])
        # an in-place update of the bank is a different bank: new capture, new result
        bank[:, :1000] = bank[:, 1000:2000].clone()
        fresh = b200knn.knn_predict(qs[0], bank, lab, C, k, 0.1)
        assert K.graph_stats["captures"] == before["captures"] + 2
        K.GRAPHS["enabled"] = False
        assert torch.equal(fresh, b200knn.knn_predict(qs[0], bank, lab, C, k, 0.1))
        K.GRAPHS["enabled"] = True
        # errors surface exactly like on the ordinary path
        bad = lab.clone()
        bad[:] = C
        with pytest.raises(RuntimeError, match="out of bounds"):
            b200knn.knn_predict(qs[0], bank, bad, C, k, 0.1)
    finally:
        K.GRAPHS["enabled"] = True
        b200knn.set_default_mode("exact")
        K.clear_call_graphs()


def test_small_batch_graph_path_with_uncertifiable_rows():
    """A bank of identical rows: no candidate level can certify anything, every call takes the
    recompute path from inside the graph route — and still returns the exact mode's result."""
    D, N, B, C, k = 256, 30000, 64, 5, 50
    g = torch.Generator(device=DEV).manual_seed(9)
    row = torch.nn.functional.normalize(torch.randn(1, D, generator=g, device=DEV), dim=1)
    bank = row.repeat(N, 1).t().contiguous()
    lab = torch.randint(0, C, (N,), generator=g, device=DEV)
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=DEV), dim=1)
    K.clear_call_graphs()
    b200knn.set_default_mode("exact")
    want = b200knn.knn_predict(q, bank, lab, C, k, 0.1)
    b200knn.set_default_mode("fp32")
    try:
        before = K.graph_stats["fallbacks"]
        for _ in range(3):
            assert torch.equal(b200knn.knn_predict(q, bank, lab, C, k, 0.1), want)
        assert K.graph_stats["fallbacks"] == before + 3
    finally:
        b200knn.set_default_mode("exact")
        K.clear_call_graphs()


# ------------------------------------------------------------------ validation hooks
def _loaders(n_bank, n_val, D, C, seed, bs=64):
    rng = np.random.default_rng(seed)
    lab = datagen.labels(n_bank, C, seed)
    emb = (datagen.relu(n_bank, D, C, seed + 1, lab) * rng.uniform(0.5, 30, (n_bank, 1))).astype(np.float32)
    vlab = datagen.labels(n_val, C, seed + 2)
    vemb = (datagen.relu(n_val, D, C, seed + 3, vlab) * rng.uniform(0.5, 30, (n_val, 1))).astype(np.float32)
    bank_loader = [(torch.from_numpy(emb[i:i + bs]), torch.from_numpy(lab[i:i + bs])) for i in range(0, n_bank, bs)]
    val_batches = [(_t(vemb[i:i + bs]), _t(vlab[i:i + bs])) for i in range(0, n_val, bs)]
    return bank_loader, val_batches


@pytest.mark.parametrize("mode", ["fp32", "exact"])
def test_hooks_equal_unhooked_flow(mode):
    """install(hooks=True) on a module with the reference's hook interface: one validation epoch
    through the fused bank build / query normalise / on-device metrics gives the predictions,
    max_accuracy / max_f1, logged values and confusion matrix of the un-hooked flow (the
    reference's F.normalize -> cat -> .t().contiguous() -> knn_predict -> metrics)."""
    import standin_module as S

    C, D = 9, 512
    bank_loader, val_batches = _loaders(3000, 500, D, C, 17)
    b200knn.set_default_mode(mode)
    try:
        # un-hooked: the drop-in symbol only, normalisation by the oracle's bit-defined F.normalize
        S.knn_predict = b200knn.knn_predict
        S.normalize = lambda x, dim=1: _t(O.normalize_rows_ref(x.float().cpu().numpy()))
        plain = S.StandInKNNModule(bank_loader, C, knn_k=5, knn_t=0.1)
        plain._device = torch.device(DEV)
        p0 = S.run_validation_epoch(plain, val_batches)
        assert plain.feature_bank.shape == (D, 3000)
        done = b200knn.install(hook_classes=(S.StandInKNNModule,))
        assert done["standin_module.StandInKNNModule"]
        try:
            hooked = S.StandInKNNModule(bank_loader, C, knn_k=5, knn_t=0.1)
            hooked._device = torch.device(DEV)
            p1 = S.run_validation_epoch(hooked, val_batches)
        finally:
            b200knn.uninstall()
        assert S.StandInKNNModule.validation_step is not b200knn.hooks.validation_step
        assert torch.equal(p0, p1)
        assert hooked.feature_bank.shape == (D, 3000) and torch.equal(hooked.feature_bank, plain.feature_bank)
        assert torch.equal(hooked.targets_bank, plain.targets_bank)
        assert abs(hooked.max_accuracy - plain.max_accuracy) < 1e-12 and abs(hooked.max_f1 - plain.max_f1) < 1e-12
        assert abs(hooked.logged["knn_accuracy"] - plain.logged["knn_accuracy"]) < 1e-12
        assert np.allclose(hooked.confusion_matrix[0], plain.confusion_matrix[0], atol=1e-12)
        assert hooked.all_preds == [] and hooked.all_targets == []
        # against torch's own F.normalize the normalised values differ by <= 2 ulp: predictions agree
        S.normalize = torch.nn.functional.normalize
        plain2 = S.StandInKNNModule(bank_loader, C, knn_k=5, knn_t=0.1)
        plain2._device = torch.device(DEV)
        p2 = S.run_validation_epoch(plain2, val_batches)
        assert float((p2 == p1).float().mean()) >= 0.995
    finally:
        b200knn.set_default_mode("exact")
        S.normalize = torch.nn.functional.normalize


def test_from_batches_preallocated_buffer():
    """from_batches writes every batch into its slice of one buffer; total_rows too small / absent
    grows it; the result equals from_rows."""
    rng = np.random.default_rng(3)
    rows = (rng.standard_normal((1000, 200)) * 4).astype(np.float32)
    lab = datagen.labels(1000, 9, 1)
    ref = b200knn.FeatureBank.from_rows(_t(rows), _t(lab))
    for total in (1000, 1200, 300, None):
        chunks = [(_t(rows[i:i + 96]), torch.from_numpy(lab[i:i + 96])) for i in range(0, 1000, 96)]
        fb = b200knn.FeatureBank.from_batches(chunks, total_rows=total)
        assert fb.n_rows == 1000 and fb.rows.is_contiguous()
        assert torch.equal(fb.rows, ref.rows) and torch.equal(fb.labels, ref.labels)


# ------------------------------------------------------------------ chunk-major order of the tensor-core kernel
@pytest.mark.parametrize("mode,k", [("bf16", 50), ("f16", 50), ("f16x2", 50), ("bf16x3", 50), ("tf32x3", 50), ("fp32", 50),
                                    ("f16", 200), ("fp32", 200)])
def test_chunk_major_order_does_not_change_results(mode, k):
    """With several query tiles per worker the kernel scans the bank chunk by chunk, parking every
    row's threshold and list fill between the chunks of its tile.  Keys must be bitwise those of
    the single-pass order (chunking off), for ragged chunk / tile / batch ends and k near the list
    capacity; fp32 must equal the exact mode."""
    from b200knn import _lib

    lib = _lib.load()
    D, N, B = 128, 20011, 37883
    g = torch.Generator(device=DEV).manual_seed(77)
    bank = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=DEV), dim=1).t().contiguous()
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=DEV), dim=1)
    raw = K.RESCORED_MODES[mode]["cand"] if mode in K.RESCORED_MODES else mode
    try:
        assert lib.b200knn_set_l2_chunk_bytes(0) == 0
        assert K.plan_info(B, N, D, k, raw)["chunks"] == 1
        one_pass = b200knn.topk_keys(q, bank, k, mode=mode)
        for chunk_bytes in (1 << 20, 300 << 10):
            assert lib.b200knn_set_l2_chunk_bytes(chunk_bytes) == 0
            plan = K.plan_info(B, N, D, k + (K.RESCORED_MODES[mode]["margin"] if mode in K.RESCORED_MODES else 0), raw)
            assert plan["chunks"] > 1 and plan["slots"] > 1 and plan["splits"] == 1, plan
            chunked = b200knn.topk_keys(q, bank, k, mode=mode)
            assert torch.equal(chunked, one_pass), (mode, plan)
        if mode == "fp32":
            sample = torch.arange(0, B, 97, device=DEV)
            assert torch.equal(chunked[sample], b200knn.topk_keys(q[sample].contiguous(), bank, k, mode="exact"))
    finally:
        lib.b200knn_set_l2_chunk_bytes(40 << 20)
