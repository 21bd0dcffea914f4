"""torchrun --nproc-per-node G scripts/sharded_check.py : the bank-row-sharded mode over NCCL
must reproduce the single-GPU result bit for bit on every rank (SURVEY.md §8e invariant)."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "self-supervised-wafermaps_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import b200knn  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (N, D, B, k, C) in [(200000, 512, 300, 200, 9), (30000, 512, 130, 20, 38), (50001, 384, 65, 5, 9),
                        (120001, 512, 19001, 200, 9)]:  # the last one is large enough for the fused exchange
    g = torch.Generator(device=dev).manual_seed(811)  # same seed on every rank -> replicated inputs
    bank = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1).t().contiguous()
    q = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=dev), dim=1)
    lab = torch.randint(0, C, (N,), generator=g, device=dev)
    for mode in ("exact", "fp32", "bf16"):
        single = b200knn.topk_keys(q, bank, k, mode=mode)
        sb = b200knn.ShardedBank.from_full(bank, lab, mode=mode)
        sharded = sb.topk_keys(q, k)
        b200knn.set_default_mode(mode)
        p1 = b200knn.knn_predict(q, bank, lab, C, k, 0.1)
        p2 = sb.knn_predict(q, C, k, 0.1)                        # all-to-all by query slice (default)
        p3 = sb.knn_predict(q, C, k, 0.1, exchange="allgather")  # the literal all-gather of keys
        assert torch.equal(p2, p3), "exchange variants disagree"
        b200knn.ShardedBank.fused_exchange = False                 # NCCL all-to-all instead of P2P stores
        p4 = sb.knn_predict(q, C, k, 0.1)
        b200knn.ShardedBank.fused_exchange = True
        assert torch.equal(p2, p4), "fused (P2P) and NCCL exchanges disagree"
        # bf16: per-shard similarities are bitwise those of the unsharded run as well (fixed-order
        # accumulation, no split-K), so even the approximate mode is shard-count invariant
        same = bool(torch.equal(single, sharded)) and bool(torch.equal(p1, p2))
        ok &= same
        print(f"rank {rank} N={N} D={D} k={k} mode={mode}: sharded==single {same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
