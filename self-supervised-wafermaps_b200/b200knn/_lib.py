"""ctypes binding of libb200knn.so (the C ABI declared in include/b200knn.h).

There is deliberately no fallback: if the shared library is missing or the
device is not an sm_100 GPU the product path raises.  The library is looked up
in-tree (``self-supervised-wafermaps_b200/lib/libb200knn.so``, built by
``csrc/build.sh`` / ``__graft_entry__.build()``) or at ``$B200KNN_LIB``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT = os.path.normpath(os.path.join(_HERE, "..", "lib", "libb200knn.so"))

# mirrors of the #defines in include/b200knn.h
F32, F16, BF16 = 0, 1, 2
LAYOUT_DN, LAYOUT_ND = 0, 1
MODE_EXACT, MODE_BF16, MODE_TF32X3, MODE_F32ROWS = 0, 1, 2, 3
SAMPLE_R = 16  # B200KNN_SAMPLE_R
MODES = {"exact": MODE_EXACT, "bf16": MODE_BF16, "tf32x3": MODE_TF32X3}

# every symbol include/b200knn.h declares: (restype, argtypes)
SIGNATURES = {
    "b200knn_version": (c_int, []),
    "b200knn_last_error": (c_char_p, []),
    "b200knn_device_ok": (c_int, []),
    "b200knn_prepare_rows": (
        c_int,
        [c_void_p, c_int, c_int, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "b200knn_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "b200knn_topk": (
        c_int,
        [c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int64,
         c_int64, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "b200knn_topk_ex": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_int64,
         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "b200knn_topk_sample": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64,
         c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "b200knn_merge": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "b200knn_decode_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200knn_vote": (
        c_int,
        [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_int, c_double, c_void_p, c_void_p,
         c_void_p, c_void_p],
    ),
    "b200knn_row_norm_max": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "b200knn_rescore": (
        c_int,
        [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
         c_int64, ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "b200knn_plan_info": (c_int, [c_int, c_int64, c_int64, c_int, c_int, ctypes.POINTER(c_int64)]),
    "b200knn_debug_topk_dump": (
        c_int,
        [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
         c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p],
    ),
}

_lib = None


def lib_path() -> str:
    return os.environ.get("B200KNN_LIB", _DEFAULT)


def load() -> ctypes.CDLL:
    """Load the shared library once; raise (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"b200knn: CUDA extension not found at {path}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` "
            "(self-supervised-wafermaps_b200/csrc/build.sh). There is no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200knn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"b200knn {what} failed ({rc}): {msg}")
