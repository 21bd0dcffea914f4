"""Host-side mirror of the reference interface for the kNN hot path.

``knn_predict`` keeps the exact signature, defaults, tensor layouts and return
type of ``lightly.utils.benchmarking.knn_predict`` as the reference binds and
calls it (``src/ssl_wafermap/models/knn.py:16`` import, ``:91-98`` and
``:205-212`` calls): ``feature`` (B,D), ``feature_bank`` (D,N) — the
``.t().contiguous()`` layout of ``knn.py:80`` — ``feature_labels`` (N,) int64,
returning (B, num_classes) int64 class rankings whose column 0 the caller
consumes (``knn.py:99``).  No normalisation happens inside (the caller does it,
``knn.py:77`` / ``:90``).

``knn_topk`` is the additive retrieval entry point (``torch.mm(...).topk(k)``
with the canonical (sim desc, index asc) order) that generalises the
notebooks' brute-force search
(``notebooks/2.0-Figures-nearest-neighbors.ipynb:54``).

Everything here is plumbing: argument checks that mirror torch's errors,
operand-preparation caching, workspace allocation through the PyTorch caching
allocator, and calls through the C ABI on the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}

TC_MODES = ("bf16", "tf32x3", "bf16x3", "f16x2", "f16")  # raw tensor-core similarity modes
MAX_K = 992  # largest k of the streaming candidate lists (capacity 1024, include/b200knn.h)
MAX_BF16_DIM = 768  # widest (padded) vector whose query tile the BF16 kernel can keep resident

_default_mode = os.environ.get("B200KNN_MODE", "fp32")

# Measurement hook (bench.py): when set to a dict, the similarity + top-k C calls ("topk") and the
# exact re-scoring calls ("rescore") are bracketed by CUDA events recorded on the launching stream
# and the (start, end) pairs are appended to the list under that name.
profile_events = None


class _Timed:
    """with _Timed("topk"): <C call>  — no-op unless bench.py switched profile_events on."""

    __slots__ = ("name", "ev")

    def __init__(self, name: str, on: bool = True):
        self.name = name
        self.ev = None
        if on and profile_events is not None and not torch.cuda.is_current_stream_capturing():
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def __enter__(self):
        if self.ev is not None:
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            profile_events.setdefault(self.name, []).append(self.ev)
        return False


# "fp32" modes: tensor-core candidate generation + exact re-scoring (csrc/rescore.cu) + a per-row
# certificate; uncertified rows fall through to the next level and finally to the exact kernel, so
# every one of these modes returns bit for bit the "exact" keys.
#
# Certificate bound of a level: |approx - exact| <= err_coef * ||q|| * max||bank row||
#                                                   + err_abs * sqrt(D_pad) * (||q|| + max||bank row||)
# with err_coef = op_coef + acc_c * D_pad * 2^-23.
#
# op_coef — operand rounding, RIGOROUS (Cauchy-Schwarz on the element-wise rounding errors; checked
# element by element in tests/test_host.py::test_fp16_operand_rounding_bounds):
#   f16   : both operands rounded to fp16: 2^-10 (1 + 2^-12), + err_abs 2^-25 per element for
#           fp16's subnormal range; operands at or beyond max_abs are refused (saturation);
#   f16x2 : fp16 queries x (fp16 hi + fp16 lo) bank: 2^-11 + 2^-22 (only the query keeps 11 bits);
#   bf16  : 2^-7 (two roundings of unit roundoff 2^-8) x 1.02 for their product term;
#   bf16x3: hi + lo operands, |x - hi - lo| <= 2^-16 |x| each + the dropped lo*lo term: 3 * 2^-16;
#   tf32x3: 2^-20 (dropped lo*lo, truncated lo).
# acc_c * D_pad * 2^-23 — accumulation in the tensor core, WORST CASE under this model of
# tcgen05.mma (validated by tests/test_gpu_tc.py::test_tmem_accumulation_error_bound on all-positive
# and adversarial operands at D = 512 / 768 / 1024): the K products of one instruction (K = 16 for
# kind::f16, 8 for kind::tf32) are exact; they and the fp32 accumulator are aligned to the largest
# exponent among them, each addend loses less than one unit in the last place of that binade
# (truncation, no guard bits assumed), and the normalised result loses less than one more.  One
# instruction therefore errs by at most (K + 2) * 2^-23 * A, A = sum_i |q_i x_i| <= ||q|| ||x||,
# and a similarity is n_mma = (MMAs per k-step) * D_pad / K chained instructions:
#       acc error <= n_mma * (K + 2) * 2^-23 * ||q|| ||x||  =  acc_c * D_pad * 2^-23 * ||q|| ||x||,
#   acc_c = (MMAs per k-step) * (K + 2) / K:  f16, bf16 1.125;  f16x2 2.25;  bf16x3 3.375;  tf32x3 3.75.
# Unlike round 1's flat 2e-5 this grows with D and holds for same-sign products (the reference's
# live embeddings are non-negative: timm ResNet-18 pooled post-ReLU, knn.py:322), where truncation
# errors do not cancel: 6.9e-5 (f16) ... 2.3e-4 (tf32x3) at D = 512.
ACC_ULP = 2.0 ** -23
LEVELS = {
    "fp32_f16": dict(cand="f16", margin=40, op_coef=1.001 * 2.0 ** -10, acc_c=1.125,
                     err_abs=1.01 * 2.0 ** -25, max_abs=6.0e4),
    "fp32_f16x2": dict(cand="f16x2", margin=40, op_coef=1.001 * (2.0 ** -11 + 2.0 ** -22), acc_c=2.25,
                       err_abs=1.01 * 2.0 ** -25, max_abs=6.0e4),
    "fp32_bf16": dict(cand="bf16", margin=128, op_coef=1.02 * 2.0 ** -7, acc_c=1.125),
    "fp32_bf16x3": dict(cand="bf16x3", margin=16, op_coef=1.001 * 3 * 2.0 ** -16, acc_c=3.375),
    "fp32_tf32": dict(cand="tf32x3", margin=8, op_coef=1.001 * 2.0 ** -20, acc_c=3.75),
}


def level_config(spec: str, dim: int) -> dict:
    """The level `spec` = "<LEVELS name>" or "<LEVELS name>@<margin>" (a wider candidate margin than
    the level's default) with the certificate's err_coef resolved for vectors of dimension `dim`."""
    name, _, margin = spec.partition("@")
    cfg = dict(LEVELS[name], name=spec)
    if margin:
        cfg["margin"] = int(margin)
    cfg["err_coef"] = cfg["op_coef"] + 1.01 * cfg["acc_c"] * padded_dim(dim) * ACC_ULP
    return cfg


# mode -> levels tried in order (then "exact").  What decides whether a row certifies is the
# candidate margin against the level's error bound E: the exact k-th similarity must beat the
# approximate (k + margin)-th by E, i.e. the row's similarities must spread by E over `margin`
# ranks.  Embeddings whose similarities are bunched (dense non-negative rows: every pair of rows is
# similar) need a wider margin, not more MMAs — re-scoring 80 more candidates costs a few percent,
# a second MMA per k-step doubles the pass.  So "fp32" runs
#   fp16 (1 MMA, margin 40 x boost)  ->  fp16 x split-fp16 (2 MMAs) with margin 160
#   ->  the same with the widest margin the lists allow (k_in = 992)  ->  exact kernel,
# where boost in {1, 2, 4} follows the fraction of rows the first level left uncertified on this
# bank (CASCADE_WIDEN / CASCADE_NARROW), and the first level is skipped for CASCADE_RETRY_CALLS
# calls if even the widest boost leaves more than CASCADE_GIVE_UP.  Vectors wider than the resident
# query tile (D_pad > 768) use the split-BF16 kernel at every level.  These statistics are
# per-process state of the single-GPU path; the sharded driver keeps its own, derived from
# all-gathered counts only, so that every rank takes the same decisions.
CASCADES = {"fp32": ("fp32_f16", "fp32_f16x2@160", "fp32_f16x2@992"), "fp32_f16": ("fp32_f16",),
            "fp32_f16x2": ("fp32_f16x2",),
            "fp32_bf16x3": ("fp32_bf16x3",),
            "fp32_bf16": ("fp32_bf16",), "fp32_tf32": ("fp32_tf32",)}
CASCADE_WIDE = ("fp32_bf16x3@64", "fp32_bf16x3@992")  # "fp32" when D_pad > MAX_BF16_DIM
CASCADE_GIVE_UP = 0.25
CASCADE_RETRY_CALLS = 64
CASCADE_WIDEN = 0.03     # more than this fraction uncertified: double the first level's margin (up to x4)
CASCADE_NARROW = 0.003   # less than this for CASCADE_CALM_CALLS calls in a row: halve it again
CASCADE_CALM_CALLS = 16
CASCADE_MAX_BOOST = 4


def next_boost(boost: int, calm: int, rows: int, uncertified: int):
    """(boost, calm) after a call that left `uncertified` of `rows` rows uncertified at the first level."""
    if rows < 128:
        return boost, calm
    frac = uncertified / rows
    if frac > CASCADE_WIDEN and boost < CASCADE_MAX_BOOST:
        return boost * 2, 0
    if frac < CASCADE_NARROW and boost > 1:
        return (boost // 2, 0) if calm + 1 >= CASCADE_CALM_CALLS else (boost, calm + 1)
    return boost, 0


# first level of each mode (bench / docs)
RESCORED_MODES = {m: LEVELS[c[0].partition("@")[0]] for m, c in CASCADES.items()}
ALL_MODES = tuple(_lib.MODES) + tuple(RESCORED_MODES)

# statistics of the last rescored call (bench / tests): rows that failed the certificate
last_rescore_stats = {"rows": 0, "uncertified": 0, "level": None, "levels": []}


def set_default_mode(mode: str) -> None:
    """Select the similarity mode ``knn_predict``/``knn_topk`` use when none is passed.

    ``"exact"``     fp32 CUDA-core contraction, sequential-fma similarities (bitwise reproducible);
    ``"fp32"``      tensor-core candidates + exact re-scoring + certificate, cascading fp16 (1 MMA per
                    k-step, k + 40..160 candidates, the margin following how bunched the bank's
                    similarities are) -> fp16 x split-fp16 (2 MMAs, k + 160) -> the same with the widest
                    margin the lists allow (k_in = 992) -> exact, for the rows each level cannot certify:
                    bitwise the ``"exact"`` result at tensor-core speed (the fp32-matching mode);
    ``"fp32_f16"`` / ``"fp32_f16x2"`` / ``"fp32_bf16x3"`` / ``"fp32_tf32"`` / ``"fp32_bf16"``  a single level
                    (then exact);
    ``"f16"``       raw tcgen05 fp16 similarities (2^-10 relative to ||q|| ||x||; BF16 mode's speed);
    ``"f16x2"``     raw tcgen05 fp16 x (fp16 hi + lo) similarities (~2.5e-4 relative to ||q|| ||x||);
    ``"bf16x3"``    raw tcgen05 bf16 hi/lo-split similarities (~1e-5 relative);
    ``"tf32x3"``    raw tcgen05 hi/lo-split TF32 similarities (fp32-class accuracy, ~1e-6);
    ``"bf16"``      raw tcgen05 BF16 operands / fp32 accumulate (fastest; recall@k reported by bench).
    """
    global _default_mode
    if mode not in ALL_MODES:
        raise ValueError(f"unknown mode {mode!r}; expected one of {sorted(ALL_MODES)}")
    _default_mode = mode


def get_default_mode() -> str:
    return _default_mode


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(name: str, t: torch.Tensor) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"b200knn: {name} is on {t.device}; this path runs on sm_100a CUDA tensors only "
            "(no CPU fallback by design)"
        )


def padded_dim(dim: int) -> int:
    return (dim + 63) // 64 * 64


def effective_mode(mode: str, dim: int) -> str:
    """The kernel mode a raw tensor-core mode runs as for vectors of dimension `dim`: bf16 / f16 /
    f16x2 keep a 128-row query tile resident in shared memory (D_pad * 256 B), so wider vectors
    go through the split-BF16 kernel, which streams both operands (and is more accurate)."""
    if mode in ("bf16", "f16x2", "f16") and padded_dim(dim) > MAX_BF16_DIM:
        return "bf16x3"
    return mode


# ----------------------------------------------------------------------------
# operand preparation (+ cache for the bank, which the reference rebuilds once
# per validation epoch and then reuses for every validation step, knn.py:67-98)
# ----------------------------------------------------------------------------
class PreparedRows:
    """K-major rows in the layout the tensor-core kernel streams with TMA."""

    __slots__ = ("mode", "n", "dim", "hi", "lo", "_f32", "_max_norm", "_src", "_delegate")

    def __init__(self, mode: str, n: int, dim: int, hi: torch.Tensor, lo: Optional[torch.Tensor]):
        self.mode, self.n, self.dim, self.hi, self.lo = mode, n, dim, hi, lo
        self._f32 = None       # (n, dim_pad) fp32 shadow rows for the exact re-scoring (bf16 candidates)
        self._max_norm = None  # device scalar: max row norm (certificate)
        self._src = None       # (tensor, vectors_are_columns) to build the shadow lazily
        self._delegate = None  # another preparation of the same bank that owns the shadow rows / norm

    def rescore_rows(self):
        """(rows_a, rows_b) fp32 row-major operands whose sum is the caller's exact value."""
        if self._delegate is not None:
            return self._delegate.rescore_rows()
        if self.mode == "tf32x3":  # hi + lo is exactly the caller's fp32 value
            return self.hi, self.lo
        if self._f32 is None:
            src, cols = self._src
            # a FeatureBank already holds its rows as zero-padded (N, D_pad) fp32: no second copy
            own = _padded_rows.get((src.data_ptr(), tuple(src.shape), tuple(src.stride())))
            own = own() if own is not None else None
            self._f32 = own if own is not None else prepare_rows(src, "f32rows", cols).hi
        return self._f32, None

    def max_norm(self) -> torch.Tensor:
        if self._delegate is not None:
            return self._delegate.max_norm()
        if self._max_norm is None:
            a, b = self.rescore_rows()
            out = torch.zeros((1,), dtype=torch.float32, device=a.device)
            with torch.cuda.device(a.device):
                _lib.check(_lib.load().b200knn_row_norm_max(a.data_ptr(), _ptr(b), self.n, a.shape[1],
                                                            out.data_ptr(), _stream()), "row_norm_max")
            self._max_norm = out
        return self._max_norm


# (D,N) bank views backed by zero-padded (N, D_pad) fp32 rows (b200knn.bank.FeatureBank)
_padded_rows = {}


def register_padded_rows(bank_view: torch.Tensor, rows_padded: torch.Tensor) -> None:
    if len(_padded_rows) > 64:
        for key in [k for k, ref in _padded_rows.items() if ref() is None]:
            _padded_rows.pop(key)
    _padded_rows[(bank_view.data_ptr(), tuple(bank_view.shape), tuple(bank_view.stride()))] = weakref.ref(rows_padded)


def normalize_rows(x: torch.Tensor, eps: float = 1e-12, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.normalize(x, dim=1)`` of (n, D) rows (reference ``knn.py:77`` / ``:90``) into zero-padded
    fp32 rows; returns the (n, D) view of the (n, D_pad) result.  fp64 norm in a fixed order.
    out: optional contiguous (n, D_pad) fp32 destination (a slice of a preallocated bank)."""
    _require_cuda("x", x)
    if x.dim() != 2:
        raise RuntimeError("normalize_rows expects (n, D) row vectors")
    if x.dtype not in _DTYPES:
        x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    n, dim = x.shape
    if out is None:
        out = torch.empty((n, padded_dim(dim)), dtype=torch.float32, device=x.device)
    elif out.shape != (n, padded_dim(dim)) or out.dtype != torch.float32 or not out.is_contiguous() \
            or out.device != x.device:
        raise ValueError("out must be a contiguous (n, padded_dim(D)) fp32 tensor on the rows' device")
    if n:
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().b200knn_normalize_rows(x.data_ptr(), _DTYPES[x.dtype], n, dim, x.stride(0),
                                                          float(eps), out.data_ptr(), _stream()), "normalize_rows")
    return out[:, :dim]


def row_sqnorms(x: torch.Tensor) -> torch.Tensor:
    """(n,) fp32 squared L2 norms of (n, D) rows, fp64 accumulation in a fixed order."""
    _require_cuda("x", x)
    if x.dtype not in _DTYPES:
        x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    n, dim = x.shape
    out = torch.empty((n,), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().b200knn_row_sqnorms(x.data_ptr(), _DTYPES[x.dtype], n, dim, x.stride(0),
                                                       out.data_ptr(), _stream()), "row_sqnorms")
    return out


def _layout_of(x: torch.Tensor, vectors_are_columns: bool) -> Tuple[torch.Tensor, int, int]:
    """Return (tensor, layout, ld) describing n vectors of dimension dim without copying
    when the strides allow it.  vectors_are_columns: x is (D, N) (bank); else (N, D)."""
    if vectors_are_columns:
        if x.stride(1) == 1 and x.stride(0) >= max(1, x.shape[1]):
            return x, _lib.LAYOUT_DN, x.stride(0)
        if x.stride(0) == 1 and x.stride(1) >= max(1, x.shape[0]):  # a .t() view of an (N, D) tensor
            return x, _lib.LAYOUT_ND, x.stride(1)
        x = x.contiguous()
        return x, _lib.LAYOUT_DN, x.stride(0)
    if x.stride(1) == 1 and x.stride(0) >= max(1, x.shape[1]):
        return x, _lib.LAYOUT_ND, x.stride(0)
    if x.stride(0) == 1 and x.stride(1) >= max(1, x.shape[0]):
        return x, _lib.LAYOUT_DN, x.stride(1)
    x = x.contiguous()
    return x, _lib.LAYOUT_ND, x.stride(0)


def prepare_rows(x: torch.Tensor, mode: str, vectors_are_columns: bool) -> PreparedRows:
    lib = _lib.load()
    if x.dtype not in _DTYPES:
        x = x.float()
    n, dim = (x.shape[1], x.shape[0]) if vectors_are_columns else (x.shape[0], x.shape[1])
    x, layout, ld = _layout_of(x, vectors_are_columns)
    dpad = padded_dim(dim)
    if mode == "bf16":
        hi = torch.empty((n, dpad), dtype=torch.bfloat16, device=x.device)
        lo = None
    elif mode == "tf32x3":
        hi = torch.empty((n, dpad), dtype=torch.float32, device=x.device)
        lo = torch.empty((n, dpad), dtype=torch.float32, device=x.device)
    elif mode == "bf16x3":
        hi = torch.empty((n, dpad), dtype=torch.bfloat16, device=x.device)
        lo = torch.empty((n, dpad), dtype=torch.bfloat16, device=x.device)
    elif mode == "f16":
        hi = torch.empty((n, dpad), dtype=torch.float16, device=x.device)
        lo = None
    elif mode == "f16x2":  # queries: one fp16 array; bank: hi + lo
        hi = torch.empty((n, dpad), dtype=torch.float16, device=x.device)
        lo = torch.empty((n, dpad), dtype=torch.float16, device=x.device) if vectors_are_columns else None
    elif mode == "f32rows":
        hi = torch.empty((n, dpad), dtype=torch.float32, device=x.device)
        lo = None
    else:
        raise ValueError(f"mode {mode!r} takes caller tensors directly")
    code = _lib.MODE_F32ROWS if mode == "f32rows" else _lib.MODES[mode]
    if n > 0:
        with torch.cuda.device(x.device):
            _lib.check(
                lib.b200knn_prepare_rows(x.data_ptr(), _DTYPES[x.dtype], layout, n, dim, ld,
                                         code, hi.data_ptr(), _ptr(lo), _stream()),
                "prepare_rows",
            )
    prep = PreparedRows(mode, n, dim, hi, lo)
    prep._src = (x, vectors_are_columns)
    return prep


class _BankCache:
    """Prepared banks keyed on the identity *and version* of the caller's tensor, so an
    in-place update or a rebuilt bank (new validation epoch) is re-prepared."""

    def __init__(self, capacity: int = 6):
        self.capacity = capacity
        self._entries = {}
        self._state = {}

    def get(self, bank: torch.Tensor, mode: str) -> PreparedRows:
        key = (bank.data_ptr(), bank._version, tuple(bank.shape), tuple(bank.stride()), bank.dtype,
               bank.device.index, mode)
        hit = self._entries.get(key)
        if hit is not None and hit[0]() is not None:
            return hit[1]
        # the reference rebuilds its bank every validation epoch (knn.py:80): drop the prepared
        # copies (several GB at 811k x 512) of banks that no longer exist before preparing a new one
        for dead in [k_ for k_, v in self._entries.items() if v[0]() is None]:
            self._entries.pop(dead)
        if mode == "f16":
            # fp16(x) is the hi array of the f16x2 split (the next cascade level): share it
            both = self.get(bank, "f16x2")
            prep = PreparedRows("f16", both.n, both.dim, both.hi, None)
            prep._src, prep._delegate = both._src, both
        else:
            prep = prepare_rows(bank, mode, vectors_are_columns=True)
        if len(self._entries) >= self.capacity:
            self._entries.pop(next(iter(self._entries)))
        try:
            ref = weakref.ref(bank)
        except TypeError:  # pragma: no cover
            ref = lambda: bank  # noqa: E731
        self._entries[key] = (ref, prep)
        return prep

    def state(self, bank: torch.Tensor) -> dict:
        """Mutable per-bank notes (cascade statistics), dropped with the bank's cache entries."""
        key = (bank.data_ptr(), bank._version, tuple(bank.shape))
        st = self._state.get(key)
        if st is None:
            if len(self._state) >= 4 * self.capacity:
                self._state.pop(next(iter(self._state)))
            st = self._state[key] = {}
        return st

    def clear(self) -> None:
        self._entries.clear()
        self._state.clear()


bank_cache = _BankCache()


class _QueryCache:
    """The prepared copy of the last query batch (used twice per call: pre-pass and main pass)."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, feature: torch.Tensor, mode: str) -> PreparedRows:
        key = (feature.data_ptr(), feature._version, tuple(feature.shape), tuple(feature.stride()),
               feature.dtype, feature.device.index, mode)
        if key != self._key or self._val is None or self._val[0]() is None:
            prep = prepare_rows(feature, mode, vectors_are_columns=False)
            self._key, self._val = key, (weakref.ref(feature), prep)
        return self._val[1]

    def clear(self) -> None:
        self._key = self._val = None


query_cache = _QueryCache()


# ----------------------------------------------------------------------------
# selection keys
# ----------------------------------------------------------------------------
def _check_feature_bank(feature: torch.Tensor, feature_bank: torch.Tensor) -> None:
    _require_cuda("feature", feature)
    _require_cuda("feature_bank", feature_bank)
    if feature.device != feature_bank.device:
        raise RuntimeError("Expected all tensors to be on the same device, but found "
                           f"{feature.device} and {feature_bank.device}")
    if feature.dim() != 2:
        raise RuntimeError("self must be a matrix")  # torch.mm's message (knn.py:89 .squeeze() hazard)
    if feature_bank.dim() != 2:
        raise RuntimeError("mat2 must be a matrix")
    if feature.shape[1] != feature_bank.shape[0]:
        raise RuntimeError(
            f"mat1 and mat2 shapes cannot be multiplied ({feature.shape[0]}x{feature.shape[1]} and "
            f"{feature_bank.shape[0]}x{feature_bank.shape[1]})"
        )


# Sampling pre-pass (tensor-core modes): the r-th best similarity of every s-th bank row is,
# except with probability P[Binomial(k, 1/s) >= r] (2e-7 for the defaults), below the true k-th
# best, so it is a valid starting threshold for the streaming selection.  It removes the list
# warm-up (the first ~k*ln(N/k) insertions), which is what a (query tile, bank shard) work item
# otherwise pays as a fixed cost; rows where the bound fails end with an empty k-th slot and are
# recomputed without it.
PREPASS = {"enabled": os.environ.get("B200KNN_PREPASS", "1") == "1", "r": 16, "rank_factor": 5.0,
           "min_k": 64, "min_sample_rows": 1024, "small_batch": 512,
           # r == 16: use the values-only kernel variant (row top-16 in registers) for the sample
           "register_sample": os.environ.get("B200KNN_REGISTER_SAMPLE", "1") == "1"}


def prepass_stride(n_rows: int, k: int, batch: Optional[int] = None) -> int:
    """Row stride of the sampling pre-pass, or 0 when it does not pay.  Large batches: only for
    k >= min_k (short lists warm up quickly).  Small batches (the reference-shaped B = 64 call) are
    planned with one bank split per SM, every split pays the list warm-up and its merge handles
    `splits` lists per row: there the threshold pays for any k (measured at N = 37,348, k = 5+40:
    125 us of warm-up prunes + 60 us of merge without it)."""
    if not PREPASS["enabled"]:
        return 0
    if k < PREPASS["min_k"] and (batch is None or batch > PREPASS["small_batch"]):
        return 0
    s = max(8, min(256, int(round(PREPASS["rank_factor"] * k / PREPASS["r"]))))
    return s if n_rows // s >= PREPASS["min_sample_rows"] else 0


def _tc_call(mode, pq, pb, B, n_visit, D, k, idx_offset, stride, tau0, dev, timed=False, sample=False,
             out=None):
    lib = _lib.load()
    if out is not None:
        assert out.shape == (B, k) and out.dtype == torch.int64 and out.is_contiguous()
    keys = out if out is not None else torch.empty((B, k), dtype=torch.int64, device=dev)
    ws_bytes = lib.b200knn_topk_workspace_bytes(B, n_visit, D, k, _lib.MODES[mode])
    if ws_bytes == 0:
        raise RuntimeError(f"b200knn: unsupported problem (B={B}, N={n_visit}, D={D}, k={k})")
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    if sample:  # values-only sampling variant (k == 16, indices are not produced)
        _lib.check(lib.b200knn_topk_sample(_lib.MODES[mode], pq.hi.data_ptr(), _ptr(pq.lo), pb.hi.data_ptr(),
                                           _ptr(pb.lo), B, n_visit, D, stride, keys.data_ptr(), ws.data_ptr(),
                                           ws_bytes, _stream()), "topk_sample")
        return keys
    with _Timed("topk", timed):
        rc = lib.b200knn_topk_ex(_lib.MODES[mode], pq.hi.data_ptr(), _ptr(pq.lo), pb.hi.data_ptr(), _ptr(pb.lo),
                                 B, n_visit, D, k, idx_offset, stride, _ptr(tau0), keys.data_ptr(),
                                 ws.data_ptr(), ws_bytes, _stream())
    _lib.check(rc, "topk")
    return keys


def kth_sim(keys: torch.Tensor) -> torch.Tensor:
    """Similarity of the last key of every row (-inf for an empty slot): a (B,) fp32 tensor."""
    B, k = keys.shape
    out = torch.empty((B,), dtype=torch.float32, device=keys.device)
    if B:
        keys = keys.contiguous()
        with torch.cuda.device(keys.device):
            _lib.check(_lib.load().b200knn_key_sim_column(keys.data_ptr(), B, k, k - 1, out.data_ptr(),
                                                          _stream()), "key_sim_column")
    return out


def vote_packed(keys: torch.Tensor, feature_labels: torch.Tensor, num_classes: int, knn_t: float,
                n_rows_out: int) -> torch.Tensor:
    """(n_rows_out, C+1) int64: class rankings of keys' rows in columns [0,C) and the per-row
    status word of b200knn_vote_ex in column C; rows past keys.shape[0] are zero."""
    lib = _lib.load()
    B, k = keys.shape
    C = int(num_classes)
    labels = feature_labels if feature_labels.dtype == torch.int64 else feature_labels.long()
    labels = labels.contiguous().view(-1)
    dev = keys.device
    out = torch.zeros((n_rows_out, C + 1), dtype=torch.int64, device=dev) if n_rows_out > B \
        else torch.empty((n_rows_out, C + 1), dtype=torch.int64, device=dev)
    if B:
        with torch.cuda.device(dev):
            flag = torch.zeros((1,), dtype=torch.int32, device=dev)
            _lib.check(lib.b200knn_vote_ex(keys.data_ptr(), labels.data_ptr(), B, k, labels.numel(), 0, C,
                                           float(knn_t), out.data_ptr(), C + 1, C, None, flag.data_ptr(),
                                           _stream()), "vote_ex")
    return out


def sample_keys(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str,
                n_rows_for_decision: Optional[int] = None) -> Optional[torch.Tensor]:
    """Pre-pass: (B, r) best keys among every s-th bank row, or None when the pre-pass is off."""
    N = feature_bank.shape[1]
    B, D = feature.shape
    s = prepass_stride(n_rows_for_decision if n_rows_for_decision is not None else N, k, B)
    if s == 0 or mode not in TC_MODES:
        return None
    mode = effective_mode(mode, D)
    r = PREPASS["r"]
    n_visit = (N + s - 1) // s
    if n_visit < r:
        return None
    with torch.cuda.device(feature.device):
        pb = bank_cache.get(feature_bank, mode)
        pq = query_cache.get(feature, mode)
        return _tc_call(mode, pq, pb, B, n_visit, D, r, 0, s, None, feature.device,
                        sample=(r == _lib.SAMPLE_R and PREPASS["register_sample"]))


def sample_scatter_supported(B: int, n_rows: int, D: int, k: int, mode: str, n_rows_for_decision: int) -> bool:
    """Whether sample_scatter applies to a shard of n_rows rows (pre-pass on, sample planned without splits)."""
    s = prepass_stride(n_rows_for_decision, k, B)
    if s == 0 or mode not in TC_MODES or PREPASS["r"] != _lib.SAMPLE_R or not PREPASS["register_sample"]:
        return False
    n_visit = (n_rows + s - 1) // s
    if n_visit < PREPASS["r"]:
        return False
    try:
        return plan_info(B, n_visit, D, PREPASS["r"], effective_mode(mode, D))["splits"] == 1
    except RuntimeError:
        return False


def sample_scatter(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str, n_rows_for_decision: int,
                   peer_ptrs, rank: int, rows_per_owner: int) -> None:
    """sample_keys whose (B, 16) sample of query row b is stored straight into the exchange buffer of
    the GPU that owns b (the sharded threshold exchange over NVLink peer memory)."""
    lib = _lib.load()
    N = feature_bank.shape[1]
    B, D = feature.shape
    s = prepass_stride(n_rows_for_decision, k, B)
    mode = effective_mode(mode, D)
    n_visit = (N + s - 1) // s
    dev = feature.device
    with torch.cuda.device(dev):
        pb = bank_cache.get(feature_bank, mode)
        pq = query_cache.get(feature, mode)
        ws_bytes = lib.b200knn_topk_workspace_bytes(B, n_visit, D, PREPASS["r"], _lib.MODES[mode])
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.b200knn_topk_sample_scatter(_lib.MODES[mode], pq.hi.data_ptr(), _ptr(pq.lo), pb.hi.data_ptr(),
                                                   _ptr(pb.lo), B, n_visit, D, s, _ptr_array(peer_ptrs), len(peer_ptrs),
                                                   rank, rows_per_owner, ws.data_ptr(), ws_bytes, _stream()),
                   "topk_sample_scatter")


def broadcast_f32(src: torch.Tensor, peer_ptrs, dst_offset: int) -> None:
    """dst[g][dst_offset + i] = src[i] on every peer g (device pointers of every rank's fp32 buffer)."""
    src = src.contiguous()
    assert src.dtype == torch.float32
    if src.numel():
        with torch.cuda.device(src.device):
            _lib.check(_lib.load().b200knn_broadcast_f32(src.data_ptr(), src.numel(), _ptr_array(peer_ptrs),
                                                         len(peer_ptrs), dst_offset, _stream()), "broadcast_f32")


def topk_keys(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: Optional[str] = None,
              idx_offset: int = 0, tau0: Optional[torch.Tensor] = None, repair: bool = True,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,k) selection keys (uint64 bit patterns in an int64 tensor), sorted descending under
    the canonical (sim desc, bank index asc) order; bank indices are offset by idx_offset.

    tau0 (tensor-core modes): caller-supplied (B,) admission thresholds — then no pre-pass and
    no validation happen here and rows may come back with empty (0) slots (sharded driver).
    repair=False (tensor-core modes): skip the host-synchronising check for rows the sampled
    threshold starved (empty k-th slot); the caller detects and repairs them (``repair_rows``)."""
    lib = _lib.load()
    mode = mode or _default_mode
    if mode not in _lib.MODES and mode not in RESCORED_MODES:
        raise ValueError(f"unknown mode {mode!r}")
    if int(k) > MAX_K:
        # Tensor.topk takes any k <= N; the streaming lists hold MAX_K keys.  Larger k is served by
        # the exact kernel in passes of <= MAX_K (a completeness path, whatever the mode).
        _check_feature_bank(feature, feature_bank)
        return _topk_keys_peeled(feature, feature_bank, int(k), idx_offset)
    if mode in RESCORED_MODES:
        return _topk_keys_rescored(feature, feature_bank, k, mode, idx_offset)
    _check_feature_bank(feature, feature_bank)
    B, D = feature.shape
    mode = effective_mode(mode, D)
    N = feature_bank.shape[1]
    k = int(k)
    if k <= 0 or k > N:
        raise RuntimeError("selected index k out of range")  # Tensor.topk's message
    dev = feature.device
    with torch.cuda.device(dev):
        if B == 0:
            return torch.empty((B, k), dtype=torch.int64, device=dev)
        if mode == "exact":
            keys = torch.empty((B, k), dtype=torch.int64, device=dev)
            ws_bytes = lib.b200knn_topk_workspace_bytes(B, N, D, k, _lib.MODE_EXACT)
            if ws_bytes == 0:
                raise RuntimeError(f"b200knn: unsupported problem (B={B}, N={N}, D={D}, k={k})")
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            q = feature if feature.dtype in _DTYPES else feature.float()
            if q.stride(1) != 1:
                q = q.contiguous()
            bank = feature_bank if feature_bank.dtype in _DTYPES else feature_bank.float()
            bank, layout, ld = _layout_of(bank, vectors_are_columns=True)
            with _Timed("topk"):
                rc = lib.b200knn_topk(_lib.MODE_EXACT, q.data_ptr(), None, _DTYPES[q.dtype], q.stride(0),
                                      bank.data_ptr(), None, _DTYPES[bank.dtype], layout, ld,
                                      B, N, D, k, idx_offset, keys.data_ptr(), ws.data_ptr(), ws_bytes,
                                      _stream())
            _lib.check(rc, "topk")
            return keys
        pb = bank_cache.get(feature_bank, mode)
        pq = query_cache.get(feature, mode)
        if tau0 is not None:  # `out`: write the keys straight into a caller buffer (sharded exchange)
            return _tc_call(mode, pq, pb, B, N, D, k, idx_offset, 1, tau0.contiguous(), dev, timed=True, out=out)
        sk = sample_keys(feature, feature_bank, k, mode)
        if sk is None:
            return _tc_call(mode, pq, pb, B, N, D, k, idx_offset, 1, None, dev, timed=True)
        keys = _tc_call(mode, pq, pb, B, N, D, k, idx_offset, 1, kth_sim(sk), dev, timed=True)
        if not repair:
            return keys
        return _repair_rows(keys, lambda rows: _tc_call(mode, _rows_of(pq, rows), pb, rows.numel(), N, D, k,
                                                        idx_offset, 1, None, dev))


def _topk_keys_peeled(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, idx_offset: int) -> torch.Tensor:
    """Exact keys for k > MAX_K: pass i selects the best min(MAX_K, remaining) keys strictly below
    the last key of pass i-1 (b200knn_topk_exact_below); ceil(k / MAX_K) scans of the bank."""
    lib = _lib.load()
    B, D = feature.shape
    N = feature_bank.shape[1]
    if k > N:
        raise RuntimeError("selected index k out of range")
    dev = feature.device
    keys = torch.empty((B, k), dtype=torch.int64, device=dev)
    if B == 0:
        return keys
    q = feature if feature.dtype in _DTYPES else feature.float()
    if q.stride(1) != 1:
        q = q.contiguous()
    bank = feature_bank if feature_bank.dtype in _DTYPES else feature_bank.float()
    bank, layout, ld = _layout_of(bank, vectors_are_columns=True)
    upper = None
    done = 0
    with torch.cuda.device(dev):
        while done < k:
            kp = min(MAX_K, k - done)
            part = torch.empty((B, kp), dtype=torch.int64, device=dev)
            ws_bytes = lib.b200knn_topk_workspace_bytes(B, N, D, kp, _lib.MODE_EXACT)
            if ws_bytes == 0:
                raise RuntimeError(f"b200knn: unsupported problem (B={B}, N={N}, D={D}, k={kp})")
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            _lib.check(lib.b200knn_topk_exact_below(q.data_ptr(), _DTYPES[q.dtype], q.stride(0), bank.data_ptr(),
                                                    _DTYPES[bank.dtype], layout, ld, B, N, D, kp, idx_offset,
                                                    _ptr(upper), part.data_ptr(), ws.data_ptr(), ws_bytes,
                                                    _stream()), "topk_exact_below")
            keys[:, done:done + kp] = part
            upper = part[:, -1].contiguous()
            done += kp
    return keys


def topk_scatter(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str, idx_offset: int,
                 tau0: Optional[torch.Tensor], peer_ptrs, rank: int, rows_per_owner: int) -> bool:
    """Fused top-k + exchange (b200knn_topk_scatter): this shard's keys of query row b are stored
    into peer_ptrs[b // rows_per_owner] (device pointers of every rank's (G, rows_per_owner, k)
    exchange buffer, mapped in this process).  Returns False when the problem is planned with
    bank splits (caller falls back to the NCCL exchange)."""
    lib = _lib.load()
    if mode not in TC_MODES:
        return False
    _check_feature_bank(feature, feature_bank)
    B, D = feature.shape
    mode = effective_mode(mode, D)
    N = feature_bank.shape[1]
    G = len(peer_ptrs)
    if B == 0 or k > N or G > 8:
        return False
    if plan_info(B, N, D, k, mode)["splits"] != 1:
        return False
    dev = feature.device
    with torch.cuda.device(dev):
        pb = bank_cache.get(feature_bank, mode)
        pq = query_cache.get(feature, mode)
        ws_bytes = lib.b200knn_topk_workspace_bytes(B, N, D, k, _lib.MODES[mode])
        if ws_bytes == 0:
            return False
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        ptrs = (ctypes.c_void_p * G)(*[int(p) for p in peer_ptrs])
        with _Timed("topk"):
            rc = lib.b200knn_topk_scatter(_lib.MODES[mode], pq.hi.data_ptr(), _ptr(pq.lo), pb.hi.data_ptr(),
                                          _ptr(pb.lo), B, N, D, k, idx_offset,
                                          _ptr(None if tau0 is None else tau0.contiguous()), ptrs, G, rank,
                                          rows_per_owner, ws.data_ptr(), ws_bytes, _stream())
        _lib.check(rc, "topk_scatter")
    return True


def recompute_rows(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str, rows: torch.Tensor,
                   idx_offset: int = 0) -> torch.Tensor:
    """Keys of the given query rows computed without any admission threshold."""
    sub = feature[rows].contiguous()
    if mode == "exact" or mode in RESCORED_MODES:
        return topk_keys(sub, feature_bank, k, "exact", idx_offset)
    B, D = sub.shape
    mode = effective_mode(mode, D)
    with torch.cuda.device(feature.device):
        pb = bank_cache.get(feature_bank, mode)
        pq = prepare_rows(sub, mode, vectors_are_columns=False)
        return _tc_call(mode, pq, pb, B, feature_bank.shape[1], D, k, idx_offset, 1, None, feature.device)


def _rows_of(pq: "PreparedRows", rows: torch.Tensor) -> "PreparedRows":
    return PreparedRows(pq.mode, rows.numel(), pq.dim, pq.hi[rows].contiguous(),
                        None if pq.lo is None else pq.lo[rows].contiguous())


def _repair_rows(keys: torch.Tensor, recompute) -> torch.Tensor:
    """Rows whose k-th slot is empty were selected under a threshold that was too high (the
    ~1e-7 tail of the sampling pre-pass): recompute them without a threshold."""
    bad = keys[:, -1] == 0
    last_prepass_stats["rows"] = keys.shape[0]
    n_bad = int(bad.sum().item())
    last_prepass_stats["repaired"] = n_bad
    if n_bad:
        rows = bad.nonzero(as_tuple=False).view(-1)
        keys[rows] = recompute(rows)
    return keys


last_prepass_stats = {"rows": 0, "repaired": 0}


def _rescored_level(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, level: str, idx_offset: int,
                    n_bad: Optional[torch.Tensor] = None):
    """One cascade level, nothing synchronises: (keys (B,k), uncertified flags (B,) int32, count int32[1]).
    n_bad: optional zeroed int32[1] the count is accumulated into (the caller's status word)."""
    lib = _lib.load()
    B, D = feature.shape
    cfg = level_config(level, D)
    N = feature_bank.shape[1]
    # the streaming lists hold at most MAX_K keys: near that limit the margin shrinks (and with it
    # the chance to certify; uncertified rows still end up exact through the next level)
    k_in = min(N, k + cfg["margin"], max(k, MAX_K))
    dev = feature.device
    # no repair pass on the candidates: a row its sampled threshold starved has an empty k_in-th
    # slot, which the re-scoring kernel reports as uncertified
    cand = topk_keys(feature, feature_bank, k_in, cfg["cand"], idx_offset, repair=False)
    pb = bank_cache.get(feature_bank, cfg["cand"])
    rows_a, rows_b = pb.rescore_rows()
    max_norm = pb.max_norm()
    # fp32 queries select the TMA-pipelined re-scoring kernels (fp16/bf16 -> fp32 is exact)
    q = feature if feature.dtype == torch.float32 else feature.float()
    if q.stride(1) != 1:
        q = q.contiguous()
    out = torch.empty((B, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        if n_bad is None:
            n_bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws_bytes = int(lib.b200knn_rescore_workspace_bytes(B, k_in)) if rows_b is None else 0
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        with _Timed("rescore"):
            rc = lib.b200knn_rescore(q.data_ptr(), _DTYPES[q.dtype], q.stride(0), rows_a.data_ptr(),
                                     _ptr(rows_b), N, D, cand.data_ptr(), B, k_in, k, idx_offset,
                                     float(cfg["err_coef"]), float(cfg.get("err_abs", 0.0)) * math.sqrt(padded_dim(D)),
                                     float(cfg.get("max_abs", 0.0)), max_norm.data_ptr(), out.data_ptr(),
                                     flags.data_ptr(), n_bad.data_ptr(), _ptr(ws), ws_bytes, _stream())
        _lib.check(rc, "rescore")
    return out, flags, n_bad


# ---- pieces of the sharded fp32 mode (b200knn/sharded.py): candidates are merged by the owner of
# a query, re-scored by the shard that owns each candidate's bank row, merged again and certified
def cascade_levels(feature_bank: torch.Tensor, mode: str, boost: int = 1) -> list:
    """The candidate levels a rescored mode tries in order on this bank (LEVELS entries + names)."""
    return [level_config(name, feature_bank.shape[0])
            for name in _cascade_levels(feature_bank, mode, track=False, boost=boost)]


def route_keys(keys: torch.Tensor, rows_per_shard: int, n_shards: int) -> torch.Tensor:
    """(n, k) keys -> (n_shards, n, k): block g keeps the keys whose bank row shard g owns."""
    n, k = keys.shape
    out = torch.empty((n_shards, n, k), dtype=torch.int64, device=keys.device)
    if n:
        keys = keys.contiguous()
        with torch.cuda.device(keys.device):
            _lib.check(_lib.load().b200knn_route_keys(keys.data_ptr(), n, k, rows_per_shard, n_shards,
                                                      out.data_ptr(), _stream()), "route_keys")
    return out


def rescore_sparse(feature: torch.Tensor, feature_bank: torch.Tensor, cand: torch.Tensor, cand_mode: str,
                   idx_offset: int) -> torch.Tensor:
    """Exact (sequential-fma) keys of the candidates in `cand` (B, k_in; lists filled from the front, empty slots only behind the last candidate;
    indices offset by idx_offset into this bank), sorted descending, zero-padded: (B, k_in)."""
    lib = _lib.load()
    B, D = feature.shape
    N = feature_bank.shape[1]
    k_in = cand.shape[1]
    dev = feature.device
    pb = bank_cache.get(feature_bank, cand_mode)
    rows_a, rows_b = pb.rescore_rows()
    q = feature if feature.dtype == torch.float32 else feature.float()
    if q.stride(1) != 1:
        q = q.contiguous()
    out = torch.empty((B, k_in), dtype=torch.int64, device=dev)
    if B == 0:
        return out
    cand = cand.contiguous()
    with torch.cuda.device(dev):
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        n_bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws_bytes = int(lib.b200knn_rescore_workspace_bytes(B, k_in)) if rows_b is None else 0
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        _lib.check(lib.b200knn_rescore(q.data_ptr(), _DTYPES[q.dtype], q.stride(0), rows_a.data_ptr(),
                                       _ptr(rows_b), N, D, cand.data_ptr(), B, k_in, k_in, idx_offset,
                                       0.0, 0.0, 0.0, pb.max_norm().data_ptr(), out.data_ptr(),
                                       flags.data_ptr(), n_bad.data_ptr(), _ptr(ws), ws_bytes, _stream()),
                   "rescore")
    return out


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def route_scatter(keys: torch.Tensor, rows_per_shard: int, n_shards: int, inbox_ptrs, row_offset: int) -> None:
    """route_keys fused with its exchange: the keys of row r that shard g owns go, compacted, into
    row (row_offset + r) of shard g's inbox (device pointers of every rank's (Q_pad, k) inbox)."""
    n, k = keys.shape
    if n:
        keys = keys.contiguous()
        with torch.cuda.device(keys.device):
            _lib.check(_lib.load().b200knn_route_scatter(keys.data_ptr(), n, k, rows_per_shard, n_shards,
                                                         _ptr_array(inbox_ptrs), row_offset, _stream()),
                       "route_scatter")


def rescore_scatter(feature: torch.Tensor, feature_bank: torch.Tensor, cand: torch.Tensor, cand_mode: str,
                    idx_offset: int, peer_ptrs, rank: int, rows_per_owner: int) -> None:
    """rescore_sparse whose sorted exact keys of query row b are stored straight into the exchange
    buffer of the GPU that owns b (peer memory; the owners zero their buffers beforehand)."""
    lib = _lib.load()
    B, D = feature.shape
    if B == 0:
        return
    N = feature_bank.shape[1]
    k_in = cand.shape[1]
    dev = feature.device
    pb = bank_cache.get(feature_bank, effective_mode(cand_mode, D))
    rows_a, rows_b = pb.rescore_rows()
    if rows_b is not None:  # tf32x3 operands (hi + lo): reassemble once
        rows_a = rows_a + rows_b
    q = feature if feature.dtype == torch.float32 else feature.float()
    if q.stride(1) != 1:
        q = q.contiguous()
    cand = cand.contiguous()
    with torch.cuda.device(dev):
        ws_bytes = int(lib.b200knn_rescore_workspace_bytes(B, k_in))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        with _Timed("rescore"):
            rc = lib.b200knn_rescore_scatter(q.data_ptr(), q.stride(0), rows_a.data_ptr(), N, D, cand.data_ptr(),
                                             B, k_in, idx_offset, _ptr_array(peer_ptrs), len(peer_ptrs), rank,
                                             rows_per_owner, ws.data_ptr(), ws_bytes, _stream())
        _lib.check(rc, "rescore_scatter")


def compact_rows(packed: torch.Tensor, col: int, mask: int, cap: int):
    """(rows int64 (cap,), count int32 (1,)): ascending numbers of the rows of `packed` (n, ld) int64
    whose column `col` has a bit of `mask` set, zero-padded; count may exceed cap (overflow)."""
    n, ld = packed.shape
    assert packed.dtype == torch.int64 and packed.stride(1) == 1
    rows = torch.empty((cap,), dtype=torch.int64, device=packed.device)
    count = torch.empty((1,), dtype=torch.int32, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(_lib.load().b200knn_compact_rows(packed.data_ptr() + 8 * col, packed.stride(0), n, mask,
                                                    rows.data_ptr(), cap, count.data_ptr(), _stream()),
                   "compact_rows")
    return rows, count


def scatter_rows(dst: torch.Tensor, src: torch.Tensor, rows: torch.Tensor, count: torch.Tensor) -> None:
    """dst[rows[i]] = src[i] for i < min(len(rows), count) — (n, w) int64 row-major tensors."""
    assert dst.dtype == torch.int64 and src.dtype == torch.int64 and dst.stride(1) == 1 and src.stride(1) == 1
    assert dst.shape[1] == src.shape[1] and src.shape[0] == rows.numel()
    with torch.cuda.device(dst.device):
        _lib.check(_lib.load().b200knn_scatter_rows(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0),
                                                    rows.data_ptr(), rows.numel(), count.data_ptr(), dst.shape[1],
                                                    _stream()), "scatter_rows")


def local_exact_keys(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str,
                     idx_offset: int, n_shards: int = 1) -> torch.Tensor:
    """(B, k+1) int64, nothing synchronises: columns [0, k) the exact top-min(k, N) keys of this bank
    shard (zero-padded), column k = 1 where the certificate failed.  One level of `mode`'s cascade at
    its base margin: a shard holds 1/G of the bank, so its rank gaps are G times wider than the
    global ones the first level failed on.  With G >= 4 shards the FIRST level (fp16, 1 MMA per
    k-step) is therefore tried again on the shard; below that the second (fp16 x split-fp16)."""
    B = feature.shape[0]
    N = feature_bank.shape[1]
    levels = _cascade_levels(feature_bank, mode, track=False)
    level = (levels[0] if n_shards >= 4 or len(levels) == 1 else levels[1]).partition("@")[0]
    k_loc = min(int(k), N)
    out = torch.zeros((B, k + 1), dtype=torch.int64, device=feature.device)
    if B and k_loc:
        keys, flags, _ = _rescored_level(feature, feature_bank, k_loc, level, idx_offset)
        out[:, :k_loc] = keys
        out[:, k] = flags
    return out


def certify(exact: torch.Tensor, approx: torch.Tensor, feature: torch.Tensor, level: dict,
            max_norm: torch.Tensor, all_rows: bool) -> torch.Tensor:
    """(n,) int32 flags: 1 where the exact k-th key does not beat the approximate k_in-th by the
    level's error bound (see b200knn_rescore)."""
    n, k = exact.shape
    k_in = approx.shape[1]
    D = feature.shape[1]
    dev = exact.device
    flags = torch.zeros((n,), dtype=torch.int32, device=dev)
    if n == 0:
        return flags
    q = feature if feature.dtype in _DTYPES else feature.float()
    if q.stride(1) != 1:
        q = q.contiguous()
    exact, approx = exact.contiguous(), approx.contiguous()
    with torch.cuda.device(dev):
        n_bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        _lib.check(_lib.load().b200knn_certify(
            q.data_ptr(), _DTYPES[q.dtype], q.stride(0), D, exact.data_ptr(), k, approx.data_ptr(), k_in, n,
            1 if all_rows else 0, float(level["err_coef"]),
            float(level.get("err_abs", 0.0)) * math.sqrt(padded_dim(D)), float(level.get("max_abs", 0.0)),
            max_norm.data_ptr(), flags.data_ptr(), n_bad.data_ptr(), _stream()), "certify")
    return flags


def bank_max_norm(feature_bank: torch.Tensor, cand_mode: str) -> torch.Tensor:
    """Device scalar: max row norm of this bank (x1.001), as the certificate uses it."""
    return bank_cache.get(feature_bank, cand_mode).max_norm()


def _cascade_levels(feature_bank: torch.Tensor, mode: str, track: bool = True, boost: int = 1):
    """The level specs `mode` tries in order on this bank.  track=False (sharded driver, captured
    graphs): the per-process statistics (margin boost, "skip the first level") are neither read nor
    updated — the caller passes its own `boost` — so every rank of a process group derives the
    same list (ranks that disagreed would mismatch their collectives)."""
    levels = list(CASCADES[mode])
    if len(levels) > 1 and padded_dim(feature_bank.shape[0]) > MAX_BF16_DIM:
        levels = list(CASCADE_WIDE)  # no resident query tile at this width
    if len(levels) > 1:
        skip = False
        if track:
            st = bank_cache.state(feature_bank)
            boost = st.get("boost", 1)
            if st.get("skip_first", 0) > 0:
                st["skip_first"] -= 1
                skip = True
        if skip:
            levels = levels[1:]
        elif boost > 1:
            name = levels[0].partition("@")[0]
            levels[0] = f"{name}@{level_config(levels[0], 1)['margin'] * boost}"
    return levels


def _cascade_fix(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, levels, rows: torch.Tensor,
                 idx_offset: int) -> torch.Tensor:
    """Keys of `rows` (uncertified at the previous level) from the remaining levels, then exact."""
    sub = feature[rows].contiguous()
    for level in levels:
        out, flags, n_bad = _rescored_level(sub, feature_bank, k, level, idx_offset)
        last_rescore_stats["levels"].append((level, int(rows.numel()), int(n_bad.item())))
        if int(n_bad.item()) == 0:
            return out
        left = flags.nonzero(as_tuple=False).view(-1)
        out[left] = _cascade_fix(sub, feature_bank, k, levels[levels.index(level) + 1:], left, idx_offset)
        return out
    last_rescore_stats["levels"].append(("exact", int(rows.numel()), 0))
    return topk_keys(sub, feature_bank, k, "exact", idx_offset)


def _note_first_level(feature_bank: torch.Tensor, mode: str, levels, rows: int, uncertified: int,
                      track: bool = True) -> None:
    last_rescore_stats["rows"], last_rescore_stats["uncertified"] = rows, uncertified
    last_rescore_stats["level"] = levels[0]
    last_rescore_stats["levels"] = [(levels[0], rows, uncertified)]  # (level, rows in, rows left uncertified)
    first = CASCADES[mode][0].partition("@")[0]
    if track and len(CASCADES[mode]) > 1 and levels[0].partition("@")[0] in (first, CASCADE_WIDE[0].partition("@")[0]) \
            and rows > 0:
        st = bank_cache.state(feature_bank)
        boost = st.get("boost", 1)
        if boost >= CASCADE_MAX_BOOST and rows >= 128 and uncertified > CASCADE_GIVE_UP * rows:
            st["skip_first"] = CASCADE_RETRY_CALLS
        st["boost"], st["calm"] = next_boost(boost, st.get("calm", 0), rows, uncertified)


def _topk_keys_rescored(feature: torch.Tensor, feature_bank: torch.Tensor, k: int, mode: str,
                        idx_offset: int = 0, defer: bool = False, track: bool = True, boost: int = 1,
                        n_bad: Optional[torch.Tensor] = None):
    """Tensor-core candidates -> exact sequential-fma re-scoring -> best k, with a per-row
    certificate; rows a level cannot certify go to the next level and finally to the "exact"
    kernel, so the keys are bitwise those of mode "exact".
    defer=True: nothing synchronises; returns (keys, flags, n_bad, fix) and the caller, after
    reading n_bad in its own synchronisation, replaces keys[rows] by fix(rows)."""
    _check_feature_bank(feature, feature_bank)
    B, D = feature.shape
    N = feature_bank.shape[1]
    k = int(k)
    if k <= 0 or k > N:
        raise RuntimeError("selected index k out of range")
    dev = feature.device
    if B == 0:
        out = torch.empty((B, k), dtype=torch.int64, device=dev)
        return (out, None, None, None) if defer else out
    levels = _cascade_levels(feature_bank, mode, track, boost)
    out, flags, n_bad = _rescored_level(feature, feature_bank, k, levels[0], idx_offset, n_bad)

    def fix(rows, n_rows_bad):
        _note_first_level(feature_bank, mode, levels, B, n_rows_bad, track)
        if rows is None:
            return None
        return _cascade_fix(feature, feature_bank, k, levels[1:], rows, idx_offset)

    if defer:
        return out, flags, n_bad, fix
    bad = int(n_bad.item())
    if bad:
        rows = flags.nonzero(as_tuple=False).view(-1)
        out[rows] = fix(rows, bad)
    else:
        _note_first_level(feature_bank, mode, levels, B, 0, track)
    return out


def decode_keys(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    sims = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    idx = torch.empty(keys.shape, dtype=torch.int64, device=keys.device)
    if keys.numel():
        with torch.cuda.device(keys.device):
            _lib.check(lib.b200knn_decode_keys(keys.data_ptr(), keys.numel(), sims.data_ptr(),
                                               idx.data_ptr(), _stream()), "decode_keys")
    return sims, idx


def merge_keys(keys_in: torch.Tensor, k_out: int) -> torch.Tensor:
    """(G,B,k_in) sorted candidate lists -> (B,k_out): the shard-merge step."""
    lib = _lib.load()
    _require_cuda("keys_in", keys_in)
    G, B, k_in = keys_in.shape
    keys_in = keys_in.contiguous()
    out = torch.empty((B, k_out), dtype=torch.int64, device=keys_in.device)
    if B:
        with torch.cuda.device(keys_in.device):
            _lib.check(lib.b200knn_merge(keys_in.data_ptr(), G, B, k_in, k_out, out.data_ptr(),
                                         _stream()), "merge")
    return out


def vote(keys: torch.Tensor, feature_labels: torch.Tensor, num_classes: int, knn_t: float,
         label_offset: int = 0, return_scores: bool = False, check_labels: bool = True,
         return_flag: bool = False, flag: Optional[torch.Tensor] = None):
    """keys (B,k) + labels (N,) -> (B,C) int64 class ranking (score desc, class asc)."""
    lib = _lib.load()
    _require_cuda("feature_labels", feature_labels)
    B, k = keys.shape
    labels = feature_labels
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.contiguous().view(-1)
    C = int(num_classes)
    dev = keys.device
    pred = torch.empty((B, C), dtype=torch.int64, device=dev)
    scores = torch.empty((B, C), dtype=torch.float64, device=dev) if return_scores else None
    if B:
        with torch.cuda.device(dev):
            if flag is None:
                flag = torch.zeros((1,), dtype=torch.int32, device=dev)
            _lib.check(lib.b200knn_vote(keys.data_ptr(), labels.data_ptr(), B, k, labels.numel(),
                                        label_offset, C, float(knn_t), pred.data_ptr(), _ptr(scores),
                                        flag.data_ptr(), _stream()), "vote")
            if check_labels:
                f = int(flag.item())
                if f == 1:
                    # the reference raises here from zeros(...).scatter(...) (lightly knn_predict)
                    raise RuntimeError("index out of bounds: a feature_labels entry is outside "
                                       f"[0, num_classes={C})")
                if f == 2:
                    raise RuntimeError("index out of bounds: neighbour index outside feature_labels")
    if return_flag:  # device int32[1]: 0 ok, 1 label outside [0,C), 2 neighbour index outside labels
        return pred, (flag if B else torch.zeros((1,), dtype=torch.int32, device=dev))
    return (pred, scores) if return_scores else pred


# ----------------------------------------------------------------------------
# the reference's public symbols for this path
# ----------------------------------------------------------------------------
def _predict_enqueue(feature, feature_bank, feature_labels, num_classes, knn_k, knn_t, mode, track=True, boost=1):
    """Everything of one tensor-core-mode call, enqueued without a host synchronisation (so it can
    be captured into a CUDA graph): (pred (B,C), status int32[2] = [vote flag, rows to recompute],
    bad_rows (B,) mask/flags, fix).  status[0]: 1 label / 2 neighbour index out of range;
    status[1]: rows a sampled threshold starved, or rows whose re-scoring could not be certified."""
    # one zeroed status word per call: [vote flag, rows to recompute] — the kernels write into it
    status = torch.zeros((2,), dtype=torch.int32, device=feature.device)
    if mode in RESCORED_MODES:
        keys, bad_rows, _, fix = _topk_keys_rescored(feature, feature_bank, knn_k, mode, defer=True,
                                                     track=track, boost=boost, n_bad=status[1:2])
    else:
        keys = topk_keys(feature, feature_bank, knn_k, mode, repair=False)
        bad_rows = keys[:, -1] == 0
        status[1:2] = bad_rows.sum().to(torch.int32)
        fix = None
    pred, _ = vote(keys, feature_labels, num_classes, knn_t, check_labels=False, return_flag=True,
                   flag=status[0:1])
    return pred, status, bad_rows, fix


def _predict_finish(pred, status, bad_rows, fix, feature, feature_bank, feature_labels, num_classes, knn_k,
                    knn_t, mode):
    """The host side of a call after its one synchronisation (`status` is a host list here)."""
    n_fix = int(status[1])
    if mode not in RESCORED_MODES:
        last_prepass_stats["rows"], last_prepass_stats["repaired"] = pred.shape[0], n_fix
    elif n_fix == 0:
        fix(None, 0)  # records the statistics of a fully certified call
    if n_fix:
        rows = bad_rows.nonzero(as_tuple=False).view(-1)
        fixed = fix(rows, n_fix) if mode in RESCORED_MODES else \
            recompute_rows(feature, feature_bank, knn_k, mode, rows)
        pred[rows], flag2 = vote(fixed, feature_labels, num_classes, knn_t, check_labels=False, return_flag=True)
        status[0] = max(int(status[0]), int(flag2.item()))
    if status[0] == 1:
        # the reference raises here from zeros(...).scatter(...) (lightly knn_predict)
        raise RuntimeError("index out of bounds: a feature_labels entry is outside "
                           f"[0, num_classes={int(num_classes)})")
    if status[0] == 2:
        raise RuntimeError("index out of bounds: neighbour index outside feature_labels")
    return pred


# ---- the reference-shaped call (B = 64 per validation_step, knn.py:87-99; scripts/WM811k_benchmark.py:71)
# is launch-latency bound: a dozen small kernels and as many Python -> C transitions per call.  For
# batches of at most GRAPH_MAX_BATCH rows the enqueued part of a call is captured ONCE per
# (bank, batch shape, k, C, t, mode) into a CUDA graph and replayed: one copy of the queries into
# the graph's static input, one graph launch, one read of the status word.  A non-zero status (rows
# to recompute, label errors — rare) re-runs the call on the ordinary path, so results and errors
# are exactly those of the un-captured call.
GRAPH_MAX_BATCH = 512
GRAPHS = {"enabled": os.environ.get("B200KNN_GRAPHS", "1") == "1", "capacity": 4}
graph_stats = {"captures": 0, "replays": 0, "fallbacks": 0}


class _CallGraph:
    __slots__ = ("graph", "q_static", "pred", "status_host", "bad_rows", "fix", "bank_ref", "keep")


_call_graphs = {}


def _graph_boost(feature_bank, mode) -> int:
    """The first level's margin boost a captured call of this bank uses (part of the graph key)."""
    if mode in RESCORED_MODES and len(CASCADES[mode]) > 1:
        return bank_cache.state(feature_bank).get("boost", 1)
    return 1


def _graph_key(feature, feature_bank, labels, num_classes, knn_k, knn_t, mode):
    return (_graph_boost(feature_bank, mode), feature_bank.data_ptr(), feature_bank._version, tuple(feature_bank.shape), tuple(feature_bank.stride()),
            feature_bank.dtype, feature_bank.device.index, labels.data_ptr(), labels._version,
            tuple(feature.shape), feature.dtype, int(num_classes), int(knn_k), float(knn_t), mode)


def _graph_for(feature, feature_bank, labels, num_classes, knn_k, knn_t, mode):
    for dead in [k_ for k_, g in _call_graphs.items() if g.bank_ref() is None]:
        _call_graphs.pop(dead)
    key = _graph_key(feature, feature_bank, labels, num_classes, knn_k, knn_t, mode)
    hit = _call_graphs.get(key)
    if hit is not None:
        return hit
    dev = feature.device
    entry = _CallGraph()
    entry.q_static = torch.empty_like(feature, memory_format=torch.contiguous_format)
    entry.q_static.copy_(feature)
    entry.status_host = torch.zeros((2,), dtype=torch.int32).pin_memory()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on a side stream, as graph capture requires
        query_cache.clear()
        _predict_enqueue(entry.q_static, feature_bank, labels, num_classes, knn_k, knn_t, mode, track=False,
                         boost=key[0])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(dev)
    query_cache.clear()
    entry.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(entry.graph):
        pred, status, bad_rows, fix = _predict_enqueue(entry.q_static, feature_bank, labels, num_classes, knn_k,
                                                       knn_t, mode, track=False, boost=key[0])
        entry.status_host.copy_(status, non_blocking=True)
    query_cache.clear()  # it now refers to tensors of the graph's private pool
    entry.pred, entry.bad_rows, entry.fix = pred, bad_rows, fix
    entry.bank_ref = weakref.ref(feature_bank)
    # the graph holds raw pointers into the prepared copies of this bank (operands, re-scoring rows,
    # max norm) and into the labels: keep them alive for as long as the graph
    entry.keep = [labels] + [prep for ref, prep in bank_cache._entries.values() if ref() is feature_bank]
    if len(_call_graphs) >= GRAPHS["capacity"]:
        _call_graphs.pop(next(iter(_call_graphs)))
    _call_graphs[key] = entry
    graph_stats["captures"] += 1
    return entry


def clear_call_graphs() -> None:
    _call_graphs.clear()


def knn_predict(feature: torch.Tensor, feature_bank: torch.Tensor, feature_labels: torch.Tensor,
                num_classes: int, knn_k: int = 200, knn_t: float = 0.1) -> torch.Tensor:
    """Drop-in for ``lightly.utils.benchmarking.knn_predict`` (reference call sites
    ``src/ssl_wafermap/models/knn.py:91-98``, ``:205-212``).

    Args and return value are the reference's: feature (B,D), feature_bank (D,N),
    feature_labels (N,), returns (B, num_classes) int64 with the predicted class in
    column 0.  Neighbour order is (similarity desc, bank index asc); class order is
    (score desc, class id asc).
    """
    _require_cuda("feature_labels", feature_labels)
    mode = _default_mode
    if feature_labels.numel() != feature_bank.shape[1]:
        _check_feature_bank(feature, feature_bank)
        raise RuntimeError(
            f"feature_labels has {feature_labels.numel()} entries for a bank of {feature_bank.shape[1]}")
    if mode == "exact" or int(knn_k) > MAX_K:
        return vote(topk_keys(feature, feature_bank, knn_k, mode), feature_labels, num_classes, knn_t)
    _check_feature_bank(feature, feature_bank)
    knn_k = int(knn_k)
    if knn_k <= 0 or knn_k > feature_bank.shape[1]:
        raise RuntimeError("selected index k out of range")
    if feature.shape[0] == 0:
        return vote(torch.empty((0, knn_k), dtype=torch.int64, device=feature.device), feature_labels,
                    num_classes, knn_t)
    # Tensor-core modes: everything is enqueued first and ONE host synchronisation reads the
    # status word (vote flag; rows to recompute: rows a sampled threshold starved, or rows whose
    # exact re-scoring could not be certified).
    if GRAPHS["enabled"] and feature.shape[0] <= GRAPH_MAX_BATCH and feature_labels.dtype == torch.int64 \
            and feature_labels.is_contiguous() and not torch.cuda.is_current_stream_capturing():
        with torch.cuda.device(feature.device):
            g = _graph_for(feature, feature_bank, feature_labels, num_classes, knn_k, knn_t, mode)
            g.q_static.copy_(feature)
            g.graph.replay()
            pred = g.pred.clone()
            torch.cuda.current_stream().synchronize()
            graph_stats["replays"] += 1
            if mode in RESCORED_MODES and len(CASCADES[mode]) > 1:
                st = bank_cache.state(feature_bank)  # the margin boost follows the captured calls too
                st["boost"], st["calm"] = next_boost(st.get("boost", 1), st.get("calm", 0), feature.shape[0],
                                                     int(g.status_host[1]))
            if not g.status_host.any():
                return pred
            # rows to recompute / label errors: the ordinary host side, on the graph's own buffers
            # (q_static still holds this call's queries)
            graph_stats["fallbacks"] += 1
            return _predict_finish(pred, g.status_host.tolist(), g.bad_rows, g.fix, g.q_static, feature_bank,
                                   feature_labels, num_classes, knn_k, knn_t, mode)
    pred, status, bad_rows, fix = _predict_enqueue(feature, feature_bank, feature_labels, num_classes, knn_k,
                                                   knn_t, mode)
    return _predict_finish(pred, status.tolist(), bad_rows, fix, feature, feature_bank, feature_labels,
                           num_classes, knn_k, knn_t, mode)


def knn_topk(feature: torch.Tensor, feature_bank: torch.Tensor, k: int,
             mode: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``torch.mm(feature, feature_bank).topk(k, dim=-1)`` without materialising the
    similarity matrix: (sims fp32 (B,k) sorted desc, idx int64 (B,k)), ties by lowest index."""
    return decode_keys(topk_keys(feature, feature_bank, k, mode))


def plan_info(B: int, N: int, D: int, k: int, mode: Optional[str] = None) -> dict:
    lib = _lib.load()
    out = (ctypes.c_int64 * 9)()
    _lib.check(lib.b200knn_plan_info_ex(_lib.MODES[mode or _default_mode], B, N, D, k, out), "plan_info")
    names = ("n_qtiles", "splits", "split_rows", "n_items", "grid", "list_capacity", "chunks", "chunk_rows", "slots")
    return dict(zip(names, [int(v) for v in out]))
