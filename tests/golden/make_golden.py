"""Generates the committed golden fixtures under tests/golden/.

Run in the BUILD container (needs /root/reference for the two real banks):
    python tests/golden/make_golden.py
Outputs (all produced by the oracle restatement, see oracle/knn_oracle.py —
lightly is not installable here, so these pin the restatement, not lightly):
  real_<model>.npz   inputs sampled from the reference's shipped embedding banks
                     data/interim/model_preds/<model>_preds_subset.pkl.xz
                     (12,449 x 512 fp16 + failureCode), plus R32 / O64 / SEQ outputs
  synth.npz          R32 / O64 / SEQ outputs of the seeded cases in tests/datagen.py
  synth_nonneg.npz   the same for the non-negative cases (datagen.RELU_CASE_NAMES, round 2)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datagen  # noqa: E402
from oracle import knn_oracle as O  # noqa: E402

REF = "/root/reference/data/interim/model_preds"


def outputs(q, bank, lab, C, k, t, prefix):
    out = {}
    tq, tb, tl = torch.from_numpy(q), torch.from_numpy(bank), torch.from_numpy(lab)
    pred, sims, idx, scores = O.knn_predict_r32_full(tq, tb, tl, C, k, t)
    out[prefix + "r32_pred"] = pred.numpy().astype(np.int16)
    out[prefix + "r32_idx"] = idx.numpy().astype(np.int32)
    s64, i64 = O.topk_o64(q, bank, min(k + 1, bank.shape[1]))
    out[prefix + "o64_idx"] = i64.astype(np.int32)
    out[prefix + "o64_sims"] = s64
    p64, sc64 = O.vote_o64(s64[:, :k], i64[:, :k], lab, C, t)
    out[prefix + "o64_pred"] = p64.astype(np.int16)
    out[prefix + "o64_scores"] = sc64
    ss, si = O.topk_seqfma(q, bank, k)
    out[prefix + "seq_sims"] = ss
    out[prefix + "seq_idx"] = si.astype(np.int32)
    ps, _ = O.vote_o64(ss, si, lab, C, t)
    out[prefix + "seq_pred"] = ps.astype(np.int16)
    return out


def real(model: str):
    import pandas as pd

    df = pd.read_pickle(os.path.join(REF, f"{model}_preds_subset.pkl.xz"))
    emb_cols = [c for c in df.columns if c not in ("waferMap", "failureType", "failureCode")]
    x = df[emb_cols].to_numpy().astype(np.float16)
    lab = df["failureCode"].to_numpy().astype(np.int64)
    rng = np.random.default_rng(20231018)
    perm = rng.permutation(len(x))
    bank_rows, q_rows = perm[:3000], perm[3000:3064]
    bank16, q16 = x[bank_rows], x[q_rows]
    res = dict(bank_rows_f16=bank16, query_rows_f16=q16, bank_labels=lab[bank_rows].astype(np.int16),
               query_labels=lab[q_rows].astype(np.int16))
    # raw (StandardScaler'd, un-normalised: knn_predict does not normalise) and L2-normalised
    for tag, norm in (("raw_", False), ("norm_", True)):
        b = bank16.astype(np.float32)
        q = q16.astype(np.float32)
        if norm:
            b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-12)
            q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
        bank = np.ascontiguousarray(b.T)
        # t=0.1 overflows exp() on the raw inputs (sims up to ~7e4) exactly as in the reference;
        # use t large enough on raw inputs that the vote stays finite
        t = 0.1 if norm else 1.0e4
        for k in (5, 200):
            res.update(outputs(q, bank, lab[bank_rows], 9, k, t, f"{tag}k{k}_"))
    np.savez_compressed(os.path.join(HERE, f"real_{model}.npz"), **res)
    print(model, {k: v.shape for k, v in res.items() if k.endswith("pred")})


def synth():
    res = {}
    for name in datagen.CASE_NAMES:
        c = datagen.make_case(name)
        res.update(outputs(c["feature"], c["bank"], c["labels"], c["C"], c["k"], c["t"], name + "_"))
    np.savez_compressed(os.path.join(HERE, "synth.npz"), **res)
    print("synth", len(res))


def synth_nonneg():
    """Round 2: non-negative (post-ReLU-like) embeddings, the reference's real input distribution."""
    res = {}
    for name in datagen.RELU_CASE_NAMES:
        c = datagen.make_case(name)
        res.update(outputs(c["feature"], c["bank"], c["labels"], c["C"], c["k"], c["t"], name + "_"))
    np.savez_compressed(os.path.join(HERE, "synth_nonneg.npz"), **res)
    print("synth_nonneg", len(res))


if __name__ == "__main__":
    if "--nonneg-only" in sys.argv:
        synth_nonneg()
        sys.exit(0)
    synth()
    synth_nonneg()
    for m in ("FastSiam", "SimSiam"):
        real(m)
