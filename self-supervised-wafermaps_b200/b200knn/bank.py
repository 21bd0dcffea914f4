"""FeatureBank — the bank build of the reference's validation epoch, laid out for the kernels.

The reference builds its bank in ``KNNBenchmarkModule.on_validation_epoch_start``
(``src/ssl_wafermap/models/knn.py:67-81``): per dataloader batch ``F.normalize(feature, dim=1)``
(``:77``), then ``torch.cat(...).t().contiguous()`` (``:80``) — an (N,D) fp32 concatenation, an
(N,D) normalised copy and a (D,N) transposed copy, which ``knn_predict`` then has to re-lay-out
again for the tensor cores.  ``FeatureBank`` does the normalisation and the relayout in ONE
kernel (``b200knn_normalize_rows``: rows -> zero-padded (N, D_pad) fp32 rows) and hands the
result to the kernels as it is; the (D,N) tensor the reference interface expects is exposed as
a zero-copy view (``.bank``), so every ``b200knn`` entry point — and the reference's own
``knn_predict(feature, bank.bank, bank.labels, ...)`` call — works on it unchanged.

SURVEY.md §8(f1).  CUDA only, like the rest of the path.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch

from . import knn as K


class FeatureBank:
    def __init__(self, rows_padded: torch.Tensor, dim: int, labels: Optional[torch.Tensor] = None):
        if rows_padded.dim() != 2 or rows_padded.dtype != torch.float32 or not rows_padded.is_contiguous() \
                or rows_padded.shape[1] != K.padded_dim(dim):
            raise ValueError("rows_padded must be contiguous (N, padded_dim(dim)) fp32 with zero pad columns")
        self.rows = rows_padded                     # (N, D_pad) fp32: exact-mode operand and re-scoring rows
        self.dim = int(dim)
        self.bank = rows_padded[:, :dim].t()        # (D, N) view: the reference's feature_bank interface
        self.labels = labels
        K.register_padded_rows(self.bank, rows_padded)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_rows(cls, rows: torch.Tensor, labels: Optional[torch.Tensor] = None,
                  normalize: bool = True) -> "FeatureBank":
        """rows: (N, D) embeddings (fp32/fp16/bf16, CUDA).  normalize=True applies the reference's
        ``F.normalize(dim=1)`` (``knn.py:77``) in the same kernel that lays the rows out."""
        K._require_cuda("rows", rows)
        if rows.dim() != 2:
            raise RuntimeError("rows must be (N, D)")
        dim = rows.shape[1]
        if normalize:
            padded = K.normalize_rows(rows)._base
        else:
            padded = K.prepare_rows(rows, "f32rows", vectors_are_columns=False).hi
        if labels is not None:
            labels = labels.to(rows.device).long().contiguous().view(-1)
            if labels.numel() != rows.shape[0]:
                raise RuntimeError(f"labels has {labels.numel()} entries for {rows.shape[0]} rows")
        return cls(padded, dim, labels)

    @classmethod
    def from_batches(cls, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]],
                     normalize: bool = True) -> "FeatureBank":
        """The loop of ``on_validation_epoch_start`` (``knn.py:70-81``): (features, targets) per
        dataloader batch; each batch is normalised + laid out as it arrives."""
        rows: List[torch.Tensor] = []
        labs: List[torch.Tensor] = []
        dim = None
        for feat, target in batches:
            K._require_cuda("feature batch", feat)
            dim = feat.shape[1]
            rows.append(K.normalize_rows(feat)._base if normalize
                        else K.prepare_rows(feat, "f32rows", vectors_are_columns=False).hi)
            labs.append(target.to(feat.device).long().view(-1))
        if dim is None:
            raise RuntimeError("no batches")
        return cls(torch.cat(rows, 0), dim, torch.cat(labs, 0))

    # ------------------------------------------------------------------ the reference's calls
    @property
    def n_rows(self) -> int:
        return self.rows.shape[0]

    def _queries(self, feature: torch.Tensor, normalize: bool) -> torch.Tensor:
        return K.normalize_rows(feature) if normalize else feature  # knn.py:90

    def knn_predict(self, feature: torch.Tensor, num_classes: int, knn_k: int = 200, knn_t: float = 0.1,
                    normalize: bool = False) -> torch.Tensor:
        """``validation_step``'s kNN call (``knn.py:90-98``); normalize=True also fuses its
        ``F.normalize(feature, dim=1)``."""
        if self.labels is None:
            raise RuntimeError("this FeatureBank was built without labels")
        return K.knn_predict(self._queries(feature, normalize), self.bank, self.labels, num_classes, knn_k, knn_t)

    def knn_topk(self, feature: torch.Tensor, k: int, normalize: bool = False, mode: Optional[str] = None):
        return K.knn_topk(self._queries(feature, normalize), self.bank, k, mode)
