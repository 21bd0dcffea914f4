"""Small cases of every kernel for `compute-sanitizer --tool memcheck|racecheck` (scripts/gpu.sh
sanitizer TOOL).  The kernels use hand-rolled mbarrier protocols across CTA pairs, TMA bulk copies
and warp-cooperative list updates — exactly where the sanitizer pays off."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "self-supervised-wafermaps_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import b200knn  # noqa: E402
import datagen  # noqa: E402
from b200knn import knn as K  # noqa: E402

K.GRAPHS["enabled"] = False
dev = "cuda:0"
names = sys.argv[1:] or ["ragged", "k5", "mixed38"]
for name in names:
    c = datagen.make_case(name)
    f = torch.from_numpy(c["feature"]).to(dev)
    bank = torch.from_numpy(c["bank"]).to(dev)
    lab = torch.from_numpy(c["labels"]).to(dev)
    ref = None
    for mode in ("exact", "fp32", "bf16", "f16", "f16x2", "bf16x3", "tf32x3"):
        b200knn.set_default_mode(mode)
        pred = b200knn.knn_predict(f, bank, lab, c["C"], c["k"], c["t"])
        keys = b200knn.topk_keys(f, bank, c["k"], mode=mode)
        torch.cuda.synchronize()
        if mode == "exact":
            ref = (pred, keys)
        elif mode == "fp32":
            assert torch.equal(pred, ref[0]) and torch.equal(keys, ref[1])
        print(name, mode, "ok", flush=True)
    # a pair-kernel shape (B > 128) with a sampled threshold, merge of bank splits, FeatureBank + metrics
    fb = b200knn.FeatureBank.from_rows(torch.randn(3000, 200, device=dev), torch.randint(0, 5, (3000,), device=dev))
    b200knn.set_default_mode("fp32")
    p = fb.knn_predict(torch.randn(300, 200, device=dev), 5, 200, 0.1, normalize=True)
    m = b200knn.knn_metrics(p[:, 0], torch.randint(0, 5, (300,), device=dev), 5)
    torch.cuda.synchronize()
    print(name, "feature bank + metrics ok", float(m["accuracy"]), flush=True)
print("sanitizer cases done")
