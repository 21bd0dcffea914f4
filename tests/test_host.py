"""CPU: host-side logic of the drop-in boundary — errors, install(), layout detection,
shard bounds.  No compute calls (those are the -m gpu tests)."""
import sys
import types

import pytest
import torch

import b200knn
from b200knn import knn as K


def test_signature_matches_reference_symbol():
    import inspect

    sig = inspect.signature(b200knn.knn_predict)
    assert list(sig.parameters) == ["feature", "feature_bank", "feature_labels", "num_classes", "knn_k", "knn_t"]
    assert sig.parameters["knn_k"].default == 200 and sig.parameters["knn_t"].default == 0.1


def test_cpu_tensors_raise_no_fallback():
    f, bank, lab = torch.randn(4, 8), torch.randn(8, 16), torch.zeros(16, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200knn.knn_predict(f, bank, lab, 3, 5, 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200knn.knn_topk(f, bank, 5)


def test_layout_detection():
    bank = torch.randn(8, 16)  # (D,N) contiguous
    t, layout, ld = K._layout_of(bank, True)
    assert (layout, ld) == (K._lib.LAYOUT_DN, 16) and t.data_ptr() == bank.data_ptr()
    nd = torch.randn(16, 8)
    t, layout, ld = K._layout_of(nd.t(), True)  # a transposed view of an (N,D) matrix
    assert (layout, ld) == (K._lib.LAYOUT_ND, 8) and t.data_ptr() == nd.data_ptr()
    q = torch.randn(4, 8)
    assert K._layout_of(q, False)[1:] == (K._lib.LAYOUT_ND, 8)
    sl = torch.randn(8, 32)[:, ::2]  # non-unit stride -> copied
    t, layout, ld = K._layout_of(sl, True)
    assert t.is_contiguous() and (layout, ld) == (K._lib.LAYOUT_DN, 16)


def test_modes():
    assert b200knn.get_default_mode() in b200knn.ALL_MODES
    with pytest.raises(ValueError):
        b200knn.set_default_mode("fp8")
    old = b200knn.get_default_mode()
    b200knn.set_default_mode("bf16")
    assert b200knn.get_default_mode() == "bf16"
    b200knn.set_default_mode(old)


def test_install_rebinds_lightly_and_consumers():
    """lightly is absent in this image: a stub package stands in for it, as the reference
    imports it (src/ssl_wafermap/models/knn.py:16)."""
    def orig(*a, **k):
        return "lightly"

    orig.__module__ = "lightly.utils.benchmarking"
    pkgs = {}
    for name in ("lightly", "lightly.utils", "lightly.utils.benchmarking"):
        pkgs[name] = types.ModuleType(name)
        pkgs[name].__path__ = []
    pkgs["lightly.utils.benchmarking"].knn_predict = orig
    consumer = types.ModuleType("ssl_wafermap.models.knn")
    consumer.knn_predict = orig  # bound at import time by `from ... import knn_predict`
    saved = {n: sys.modules.get(n) for n in list(pkgs) + ["ssl_wafermap.models.knn"]}
    try:
        sys.modules.update(pkgs)
        sys.modules["ssl_wafermap.models.knn"] = consumer
        done = b200knn.install()
        assert done["lightly.utils.benchmarking"] and done["ssl_wafermap.models.knn"]
        assert pkgs["lightly.utils.benchmarking"].knn_predict is b200knn.knn_predict
        assert consumer.knn_predict is b200knn.knn_predict
        b200knn.install()  # idempotent
        b200knn.uninstall()
        assert pkgs["lightly.utils.benchmarking"].knn_predict is orig and consumer.knn_predict is orig
    finally:
        b200knn.uninstall()
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def test_shard_bounds_cover_and_are_disjoint():
    for n in (0, 1, 7, 811457, 16777216):
        for ws in (1, 2, 3, 4, 8):
            spans = [b200knn.shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d


def test_padded_dim():
    assert [K.padded_dim(d) for d in (1, 64, 72, 384, 512, 768)] == [64, 64, 128, 384, 512, 768]


def test_bench_inputs_are_shard_reproducible():
    """bench.py generates the synthetic bank in 65,536-row chunks with one generator per chunk, so a
    rank that only materialises its own rows (sharded runs, config c5) gets exactly the rows of the
    full bank, and the replicated labels / queries do not depend on the shard."""
    import sys

    import torch

    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    n, q, d = 150000, 37, 16
    full_bank, full_lab, full_q = bench.make_inputs("cpu", n, q, d, seed=3)
    assert full_bank.shape == (n, d) and full_lab.shape == (n,) and full_q.shape == (q, d)
    assert torch.allclose(full_bank.norm(dim=1), torch.ones(n), atol=1e-5)
    for lo, hi in ((0, 50000), (50000, 100000), (100000, 150000), (65530, 65540)):
        bank, lab, qq = bench.make_inputs("cpu", n, q, d, seed=3, row_range=(lo, hi))
        assert torch.equal(bank, full_bank[lo:hi]) and torch.equal(lab, full_lab) and torch.equal(qq, full_q)


def test_fp16_candidate_levels_operand_error_bounds():
    """The rigorous (operand-rounding) part of the certificate coefficients of the fp16 candidate
    levels (b200knn/knn.py LEVELS): with fp16-rounded operands and EXACT accumulation,
        f16  : |q16.x16 - q.x|          <= 2^-10 (1+2^-12) |q||x| + 2^-25 sqrt(D) (|q|+|x|)
        f16x2: |q16.(x_hi+x_lo) - q.x|  <= (2^-11 + 2^-22)(1+2^-11) |q||x| + 2^-25 sqrt(D) (|q|+|x|)
    including fp16's subnormal range (the 2^-25 terms).  The levels' err_coef / err_abs must dominate
    these (the remainder of err_coef is the D-scaled worst-case allowance for the accumulation in the
    tensor core, acc_c * D_pad * 2^-23, see b200knn/knn.py)."""
    import numpy as np

    from b200knn.knn import level_config

    rng = np.random.default_rng(0)
    D = 512
    worst = {"f16": 0.0, "f16x2": 0.0}
    for scale_q, scale_x in [(1.0, 1.0), (1e-3, 1.0), (1.0, 1e-4), (3e-4, 3e-4), (50.0, 200.0), (1e-5, 1e-5)]:
        q = (rng.standard_normal((64, D)) * scale_q / np.sqrt(D)).astype(np.float32)
        x = (rng.standard_normal((256, D)) * scale_x / np.sqrt(D)).astype(np.float32)
        q[0] = np.abs(q[0])  # aligned signs: rounding errors add up instead of cancelling
        x[0] = np.abs(x[0])
        q16 = q.astype(np.float16).astype(np.float64)
        x_hi = x.astype(np.float16)
        x_lo = (x - x_hi.astype(np.float32)).astype(np.float16)
        exact = q.astype(np.float64) @ x.astype(np.float64).T
        qn = np.linalg.norm(q.astype(np.float64), axis=1)[:, None]
        xn = np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
        tail = 2.0 ** -25 * np.sqrt(D) * (qn + xn)
        for name, approx, coef in (
                ("f16", q16 @ x_hi.astype(np.float64).T, 2.0 ** -10 * (1 + 2.0 ** -12)),
                ("f16x2", q16 @ (x_hi.astype(np.float64) + x_lo.astype(np.float64)).T,
                 (2.0 ** -11 + 2.0 ** -22) * (1 + 2.0 ** -11))):
            err = np.abs(approx - exact)
            bound = coef * qn * xn + tail
            assert (err <= bound).all(), (name, scale_q, scale_x, float((err / bound).max()))
            worst[name] = max(worst[name], float((err / bound).max()))
            lv = level_config("fp32_" + name, D)
            assert lv["op_coef"] >= coef and lv["err_coef"] >= lv["op_coef"] + lv["acc_c"] * D * 2.0 ** -23 and lv["err_coef"] >= coef and lv["err_abs"] >= 2.0 ** -25 and lv["max_abs"] < 65504.0
            # the certificate's E (host side multiplies err_abs by sqrt(padded D)) dominates the bound
            e_cert = lv["err_coef"] * qn * xn + lv["err_abs"] * np.sqrt(D) * (qn + xn)
            assert (e_cert >= bound).all()
    print("largest observed error / rigorous bound:", worst)
