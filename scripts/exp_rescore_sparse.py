"""What one shard of the 8-GPU fp32 step does in its re-scoring phase, on ONE GPU: Q queries whose
routed candidate lists (k_in slots wide) hold ~k_in/G of this shard's rows each, compacted to the
front.  Prints the time of b200knn_rescore on them and the achieved row bandwidth."""
import os
import sys
import time

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "self-supervised-wafermaps_b200"))
import torch  # noqa: E402

import b200knn  # noqa: E402
from b200knn import knn as K  # noqa: E402

dev = "cuda:0"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Q, N, D, k_in = 151552, 811457 // G, 512, 240
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1)
fb = b200knn.FeatureBank.from_rows(rows, normalize=False)
q = torch.nn.functional.normalize(torch.randn(Q, D, generator=g, device=dev), dim=1)
per_list = k_in // G
idx = torch.randint(0, N, (Q, per_list), generator=g, device=dev)
sim = torch.rand(Q, per_list, generator=g, device=dev)
u = sim.view(torch.int32).to(torch.int64) | 0x80000000
keys = (u << 32) | (0xFFFFFFFF - idx)
cand = torch.zeros((Q, k_in), dtype=torch.int64, device=dev)
cand[:, :per_list] = keys
for name, c in (("sparse (1/%d of the slots)" % G, cand),):
    for _ in range(3):
        out = K.rescore_sparse(q, fb.bank, c, "f16", 0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        out = K.rescore_sparse(q, fb.bank, c, "f16", 0)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    gb = Q * per_list * D * 4 / 1e9
    print(f"{name}: {ms:.3f} ms for {gb:.2f} GB of candidate rows = {gb / ms:.2f} TB/s", flush=True)
