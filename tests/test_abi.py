"""CPU: the C-ABI shared library loads and exports every symbol include/b200knn.h
declares; argument validation that needs no GPU returns the documented codes."""
import ctypes
import os
import re

import pytest

from b200knn import _lib

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200knn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200knn_\w+)\s*\(", text)))


def test_header_symbols_all_bound_in_python():
    assert set(declared_symbols()) == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.b200knn_version() == 142
    assert isinstance(lib.b200knn_last_error(), bytes)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # k > N: Tensor.topk's "selected index k out of range"
    rc = lib.b200knn_topk(_lib.MODE_EXACT, 1, None, 0, 4, 1, None, 0, 0, 8, 2, 8, 4, 9, 0, 1, 1, 1 << 20, None)
    assert rc == -1 and b"k out of range" in lib.b200knn_last_error()
    rc = lib.b200knn_merge(1, 2, 4, 3, 7, 1, None)  # k_out > G*k_in
    assert rc == -1
    rc = lib.b200knn_vote(1, 1, 4, 3, 10, 0, 5, 0.0, 1, None, 1, None)  # t == 0
    assert rc == -1
    rc = lib.b200knn_prepare_rows(1, 7, 0, 4, 8, 4, 1, 1, None, None)  # bad dtype
    assert rc == -1
    assert lib.b200knn_topk_workspace_bytes(64, 1000, 512, 2000, 0) == 0  # k too large


def test_plan_is_deterministic_and_covers_bank():
    from b200knn import plan_info

    for (B, N, D, k) in [(64, 811457, 512, 200), (75776, 811457, 512, 200), (5703, 30000, 512, 20), (1, 5, 8, 5)]:
        for mode in ("exact", "bf16", "tf32x3"):
            p = plan_info(B, N, D, k, mode)
            assert p == plan_info(B, N, D, k, mode)
            assert p["splits"] * p["split_rows"] >= N > (p["splits"] - 1) * p["split_rows"]
            assert p["n_items"] == p["n_qtiles"] * p["splits"] and 1 <= p["grid"] <= p["n_items"]
            assert p["list_capacity"] >= k + 32


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setenv("B200KNN_LIB", "/nonexistent/libb200knn.so")
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_graft_entry_build_runs():
    """The driver's build check: compiles whatever is stale, builds the oracle's C restatement and
    verifies the loaded library against include/b200knn.h."""
    import __graft_entry__

    __graft_entry__.build()
