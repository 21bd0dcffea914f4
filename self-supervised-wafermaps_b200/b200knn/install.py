"""Make the reference scripts pick the B200 path up unchanged.

The reference binds the symbol at import time
(``from lightly.utils.benchmarking import knn_predict``,
``src/ssl_wafermap/models/knn.py:16``), so ``install()`` must

1. replace ``lightly.utils.benchmarking.knn_predict`` (and
   ``lightly.utils.benchmarking.knn.knn_predict`` in lightly versions that have
   that submodule) — effective for modules imported afterwards, e.g. when
   called before ``scripts/WM811k_benchmark.py`` imports ``ssl_wafermap``; and
2. rebind the module global of any already-imported consumer
   (``ssl_wafermap.models.knn`` and the script modules that define their own
   copy of ``KNNBenchmarkModule``).

Set ``B200KNN_AUTOINSTALL=1`` and import ``b200knn`` from ``sitecustomize`` /
a ``.pth`` file to do this with zero edits to the reference.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Tuple

_saved: List[Tuple[object, str, object]] = []

_PROVIDERS = ("lightly.utils.benchmarking", "lightly.utils.benchmarking.knn")
_CONSUMERS = ("ssl_wafermap.models.knn",)


def install(extra_consumers: Tuple[str, ...] = (), hooks: bool = False,
            hook_classes: Tuple[type, ...] = ()) -> Dict[str, bool]:
    """Rebind ``knn_predict`` everywhere the reference looks it up.  Returns which
    modules were patched.  Idempotent; ``uninstall()`` restores the originals.

    hooks=True additionally replaces the validation hooks of the reference's
    ``KNNBenchmarkModule`` / ``WandBKNNBenchmarkModule`` (``src/ssl_wafermap/models/knn.py:67-133``,
    ``:181-215``) by the fused bank build / query normalise / on-device metrics of
    ``b200knn.hooks``; hook_classes names further classes with the same
    hook interface (e.g. the copy ``scripts/WM811k_benchmark.py`` defines)."""
    from .knn import knn_predict

    done: Dict[str, bool] = {}
    for name in _PROVIDERS:
        mod = sys.modules.get(name)
        if mod is None:
            try:
                mod = __import__(name, fromlist=["knn_predict"])
            except Exception:
                done[name] = False
                continue
        if hasattr(mod, "knn_predict"):
            if getattr(mod, "knn_predict") is not knn_predict:
                _saved.append((mod, "knn_predict", getattr(mod, "knn_predict")))
                setattr(mod, "knn_predict", knn_predict)
            done[name] = True
        else:
            done[name] = False
    for name in _CONSUMERS + tuple(extra_consumers):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "knn_predict"):
            if getattr(mod, "knn_predict") is not knn_predict:
                _saved.append((mod, "knn_predict", getattr(mod, "knn_predict")))
                setattr(mod, "knn_predict", knn_predict)
            done[name] = True
        else:
            done[name] = False
    # scripts/WM811k_benchmark.py run as __main__ also does `from lightly... import knn_predict`
    main = sys.modules.get("__main__")
    if main is not None and getattr(main, "knn_predict", None) is not None \
            and getattr(main, "knn_predict") is not knn_predict \
            and getattr(getattr(main, "knn_predict"), "__module__", "").startswith("lightly"):
        _saved.append((main, "knn_predict", getattr(main, "knn_predict")))
        setattr(main, "knn_predict", knn_predict)
        done["__main__"] = True
    if hooks or hook_classes:
        from .hooks import install_hooks

        done.update(install_hooks(hook_classes))
    return done


def uninstall() -> None:
    while _saved:
        mod, attr, orig = _saved.pop()
        setattr(mod, attr, orig)
    from .hooks import uninstall_hooks

    uninstall_hooks()


if os.environ.get("B200KNN_AUTOINSTALL") == "1":  # pragma: no cover - exercised via subprocess test
    try:
        install(hooks=os.environ.get("B200KNN_HOOKS") == "1")
    except Exception:
        pass
