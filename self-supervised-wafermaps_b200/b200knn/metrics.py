"""On-device kNN validation metrics.

The reference scores its kNN predictions after every validation epoch with torchmetrics
(``src/ssl_wafermap/models/knn.py:52-53, 104-129``): ``MulticlassAccuracy(average="macro")``,
``MulticlassF1Score(average="macro")`` and ``MulticlassConfusionMatrix(normalize="true")``,
followed by a ``.cpu().numpy()`` of the matrix.  Everything follows from the (C,C) count matrix,
which one small kernel (``b200knn_confusion``) builds from ``pred_labels[:, 0]`` and the
targets; the derived numbers stay on the device until the caller reads them.

torchmetrics semantics restated [recalled — torchmetrics is not installed here]:
  macro accuracy = mean over classes of recall_c = tp_c / (tp_c + fn_c),
  macro F1       = mean over classes of 2 tp_c / (2 tp_c + fp_c + fn_c),
  both with 0/0 := 0 and classes that occur neither in the predictions nor in the targets left
  out of the mean; normalize="true" divides every row (true class) by its sum, 0/0 := 0.
The tests check them against scikit-learn's recall_score / f1_score / confusion_matrix.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from . import knn as K


def confusion_counts(pred: torch.Tensor, target: torch.Tensor, num_classes: int) -> torch.Tensor:
    """(C, C) int64 counts[t, p] = #(target == t and pred == p)."""
    K._require_cuda("pred", pred)
    K._require_cuda("target", target)
    pred = pred.long().contiguous().view(-1)
    target = target.to(pred.device).long().contiguous().view(-1)
    if pred.numel() != target.numel():
        raise RuntimeError(f"pred has {pred.numel()} entries, target {target.numel()}")
    C = int(num_classes)
    counts = torch.zeros((C, C), dtype=torch.int64, device=pred.device)
    if pred.numel():
        with torch.cuda.device(pred.device):
            flag = torch.zeros((1,), dtype=torch.int32, device=pred.device)
            _lib.check(_lib.load().b200knn_confusion(pred.data_ptr(), target.data_ptr(), pred.numel(), C,
                                                     counts.data_ptr(), flag.data_ptr(), K._stream()), "confusion")
            if int(flag.item()):
                raise RuntimeError(f"a prediction or target is outside [0, num_classes={C})")
    return counts


def metrics_from_counts(counts: torch.Tensor) -> Dict[str, torch.Tensor]:
    """macro accuracy, macro F1 (0-dim fp64 tensors) and the row-normalised confusion matrix."""
    c = counts.double()
    tp = c.diagonal()
    support = c.sum(1)            # tp + fn
    predicted = c.sum(0)          # tp + fp
    present = (support + predicted) > 0
    n_present = present.sum().clamp(min=1)
    recall = torch.where(support > 0, tp / support.clamp(min=1), torch.zeros_like(tp))
    denom = support + predicted   # 2 tp + fp + fn
    f1 = torch.where(denom > 0, 2 * tp / denom.clamp(min=1), torch.zeros_like(tp))
    return {
        "accuracy": (recall * present).sum() / n_present,
        "f1": (f1 * present).sum() / n_present,
        "confusion": torch.where(support.view(-1, 1) > 0, c / support.clamp(min=1).view(-1, 1), torch.zeros_like(c)),
        "counts": counts,
    }


def knn_metrics(pred: torch.Tensor, target: torch.Tensor, num_classes: int) -> Dict[str, torch.Tensor]:
    """``on_validation_epoch_end`` of the reference (``knn.py:104-129``) for predictions
    ``pred_labels[:, 0]`` and their targets."""
    return metrics_from_counts(confusion_counts(pred, target, num_classes))
