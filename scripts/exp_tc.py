"""Experiment driver (not a test, not the bench): times the tensor-core kernel through the
debug entry point with pipeline stages switched off, to see which role bounds it.
flags: 1 no selection, 2 no TMEM loads, 4 no MMA issue, 8 no bank TMA loads."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "self-supervised-wafermaps_b200"))
import torch  # noqa: E402

import b200knn  # noqa: E402
from b200knn import _lib  # noqa: E402

dev = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 18944
N, D, k = 811457, 512, int(sys.argv[3]) if len(sys.argv) > 3 else 200
g = torch.Generator(device=dev).manual_seed(1)
bank = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1)
q = torch.nn.functional.normalize(torch.randn(Q, D, generator=g, device=dev), dim=1)
pb = b200knn.prepare_rows(bank.t(), mode, vectors_are_columns=True)
pq = b200knn.prepare_rows(q, mode, vectors_are_columns=False)
lib = _lib.load()
ws_bytes = lib.b200knn_topk_workspace_bytes(Q, N, D, k, _lib.MODES[mode])
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
keys = torch.empty((Q, k), dtype=torch.int64, device=dev)
diag = torch.zeros(4, dtype=torch.int32, device=dev)
print("plan", b200knn.plan_info(Q, N, D, k, mode))
flops = 2.0 * Q * N * D


def run(flags, reps=3):
    ts = []
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.b200knn_debug_topk_dump(_lib.MODES[mode], pq.hi.data_ptr(), None if pq.lo is None else pq.lo.data_ptr(),
                                         pb.hi.data_ptr(), None if pb.lo is None else pb.lo.data_ptr(), Q, N, D, k,
                                         keys.data_ptr(), ws.data_ptr(), ws_bytes, None, diag.data_ptr(), flags,
                                         torch.cuda.current_stream().cuda_stream)
        b.record()
        _lib.check(rc, "dbg")
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts[1:])
    print(f"flags={flags:2d}  {t:8.2f} ms   {flops / t / 1e9:8.1f} TFLOP/s-equivalent", flush=True)


for f in (0, 1, 3, 2 | 4, 4, 8, 8 | 1, 8 | 3, 4 | 8 | 3):
    run(f)
