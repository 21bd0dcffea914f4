"""torchrun --nproc-per-node G scripts/sharded_check.py : the bank-row-sharded mode over NCCL / NVLink
peer memory must reproduce the single-GPU result bit for bit on every rank (SURVEY.md §8e invariant).
Runs the comparisons of tests/test_gpu_multi.py (which pytest spawns on 2 GPUs) on G ranks and prints
one line per comparison; exit code 0 only if every rank passed everything."""
import os
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (ROOT, os.path.join(ROOT, "self-supervised-wafermaps_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from test_gpu_multi import run_cases  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
failures = run_cases(rank, dev, log=lambda s: print(s, flush=True))
flag = torch.tensor([len(failures)], device=dev)
dist.all_reduce(flag)
if rank == 0:
    print(f"world {world}: {'ALL BITWISE EQUAL' if int(flag.item()) == 0 else 'FAILURES'} {failures}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
