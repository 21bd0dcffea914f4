#!/usr/bin/env python
"""Pin the oracle the moment `lightly` is available (it is not in this image: DESIGN.md §2).

    python oracle/diff_against_lightly.py

Runs lightly.utils.benchmarking.knn_predict (the function the reference imports,
src/ssl_wafermap/models/knn.py:16) and the oracle's restatement knn_predict_r32 on every seeded
case of tests/datagen.py and on the reference's two real banks (tests/golden/real_*.npz), on CPU,
and requires identical (B, C) outputs.  Exit code 0 = identical everywhere (then "parity
unpinned" can be dropped from oracle/knn_oracle.py and DESIGN.md), 2 = lightly not importable,
1 = a difference (printed).  TEST INFRASTRUCTURE, like everything under oracle/."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
from oracle import knn_oracle as O  # noqa: E402


def main() -> int:
    try:
        from lightly.utils.benchmarking import knn_predict as lightly_knn_predict
    except Exception as e:  # noqa: BLE001
        print(f"lightly is not importable here ({type(e).__name__}: {e}); the oracle stays unpinned")
        return 2
    bad = 0
    cases = []
    for name in datagen.CASE_NAMES + datagen.RELU_CASE_NAMES:
        c = datagen.make_case(name)
        cases.append((name, c["feature"], c["bank"], c["labels"], c["C"], c["k"], c["t"]))
    for model in ("FastSiam", "SimSiam"):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"real_{model}.npz"))
        b = g["bank_rows_f16"].astype(np.float32)
        q = g["query_rows_f16"].astype(np.float32)
        b /= np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-12)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
        for k in (5, 200):
            cases.append((f"real_{model}_k{k}", q, np.ascontiguousarray(b.T), g["bank_labels"].astype(np.int64), 9, k, 0.1))
    for name, q, bank, lab, C, k, t in cases:
        tq, tb, tl = torch.from_numpy(q), torch.from_numpy(bank), torch.from_numpy(lab)
        want = lightly_knn_predict(tq, tb, tl, C, k, t)
        got = O.knn_predict_r32(tq, tb, tl, C, k, t)
        same = bool(torch.equal(want, got))
        print(f"{name}: {'identical' if same else 'DIFFERENT'}")
        bad += 0 if same else 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
