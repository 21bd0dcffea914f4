"""Seeded synthetic inputs for the kNN path (SURVEY.md §8d generators), numpy, CPU.

gauss     : x ~ N(0, I_D), rows L2-normalised (the caller's F.normalize, knn.py:77/:90)
clustered : C unit centroids; label ~ class prior; x = normalise(centroid[label] + 1.4*N(0,I)/sqrt(D))
relu      : clustered, passed through max(0, .) before the normalisation — NON-NEGATIVE rows like the
            reference's live embeddings (timm ResNet-18 pooled post-ReLU features, knn.py:322, then
            F.normalize, :77): all similarities positive and bunched, accumulation errors cannot cancel
absgauss  : |N(0, I_D)| normalised (dense non-negative rows, every pair of rows has similarity ~0.64)
Label priors follow the reference's shipped data: WM-811K 9-class counts
(data/interim/model_preds/*_preds_subset.pkl.xz failureCode histogram) and a flat
38-class MixedWM38 prior.
"""
from __future__ import annotations

import numpy as np

WM811K_PRIOR = np.array([859, 111, 1037, 1936, 719, 30, 173, 239, 7345], dtype=np.float64)


def _normalise(x: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return (x / np.maximum(n, 1e-12)).astype(np.float32)


def labels(n: int, num_classes: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if num_classes == 9:
        p = WM811K_PRIOR / WM811K_PRIOR.sum()
    else:
        p = np.full(num_classes, 1.0 / num_classes)
    return rng.choice(num_classes, size=n, p=p).astype(np.int64)


def gauss(n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _normalise(rng.standard_normal((n, d), dtype=np.float32))


def clustered(n: int, d: int, num_classes: int, seed: int, lab: np.ndarray | None = None,
              centroid_seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    cent = _normalise(np.random.default_rng(centroid_seed).standard_normal((num_classes, d)))
    if lab is None:
        lab = labels(n, num_classes, seed + 1)
    x = cent[lab] + 1.4 * rng.standard_normal((n, d), dtype=np.float32) / np.sqrt(d)
    return _normalise(x)


def relu(n: int, d: int, num_classes: int, seed: int, lab: np.ndarray | None = None,
         centroid_seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    cent = _normalise(np.random.default_rng(centroid_seed).standard_normal((num_classes, d)))
    if lab is None:
        lab = labels(n, num_classes, seed + 1)
    x = cent[lab] + 1.4 * rng.standard_normal((n, d), dtype=np.float32) / np.sqrt(d)
    return _normalise(np.maximum(x, 0.0))


def absgauss(n: int, d: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _normalise(np.abs(rng.standard_normal((n, d), dtype=np.float32)))


def make_case(name: str):
    """Named small cases shared by the golden generator and the tests.
    Returns dict(feature (B,D) f32, bank (D,N) f32 contiguous, labels (N,), C, k, t)."""
    cases = {
        # name: (kind, B, N, D, C, k)
        "gauss_small": ("gauss", 48, 4096, 512, 9, 200),
        "clustered_small": ("clustered", 48, 4096, 512, 9, 200),
        "clustered_d384": ("clustered", 33, 3000, 384, 9, 200),
        "mixed38": ("clustered", 40, 3000, 512, 38, 20),
        "k5": ("clustered", 64, 2500, 512, 9, 5),
        "ragged": ("gauss", 7, 1001, 72, 5, 10),
        "relu_small": ("relu", 48, 4096, 512, 9, 200),
        "absgauss_small": ("absgauss", 40, 4096, 512, 9, 200),
    }
    kind, B, N, D, C, k = cases[name]
    seed = 811 + sum(map(ord, name))
    lab = labels(N, C, seed + 2)
    if kind == "gauss":
        bank = gauss(N, D, seed)
        q = gauss(B, D, seed + 1)
    elif kind == "absgauss":
        bank = absgauss(N, D, seed)
        q = absgauss(B, D, seed + 1)
    elif kind == "relu":
        bank = relu(N, D, C, seed, lab)
        q = relu(B, D, C, seed + 1)
    else:
        bank = clustered(N, D, C, seed, lab)
        q = clustered(B, D, C, seed + 1)
    return dict(feature=q, bank=np.ascontiguousarray(bank.T), labels=lab, C=C, k=k, t=0.1)


CASE_NAMES = ("gauss_small", "clustered_small", "clustered_d384", "mixed38", "k5", "ragged")
# non-negative embeddings (post-ReLU features): parity cases added in round 2
RELU_CASE_NAMES = ("relu_small", "absgauss_small")
