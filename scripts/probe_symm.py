"""Probe (2+ GPUs): does torch symmetric memory rendezvous work on this box, and do P2P stores land?"""
import os
import sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((world, 1024), dtype=torch.int64, device=dev)
    t.zero_()
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f"rank {rank}: rendezvous ok, ptrs {[hex(p) for p in hdl.buffer_ptrs]} multicast {hdl.has_multicast_support if hasattr(hdl,'has_multicast_support') else None}", flush=True)
    hdl.barrier()
    for peer in range(world):
        buf = hdl.get_buffer(peer, (world, 1024), torch.int64)
        buf[rank].fill_(100 * rank + peer)
    hdl.barrier()
    torch.cuda.synchronize()
    want = torch.tensor([100 * r + rank for r in range(world)], device=dev)
    ok = bool((t[:, 0] == want).all()) and bool((t[:, -1] == want).all())
    print(f"rank {rank}: p2p stores visible: {ok}", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"rank {rank}: symmetric memory FAILED: {type(e).__name__}: {e}", flush=True)
    ok = False
# CUDA IPC fallback probe
try:
    x = torch.zeros(1024, dtype=torch.int64, device=dev)
    info = x.untyped_storage()._share_cuda_()
    infos = [None] * world
    dist.all_gather_object(infos, info)
    peer = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*infos[peer])
    pt = torch.empty(0, dtype=torch.int64, device=st.device).set_(st)
    dist.barrier()
    pt[:4] = rank + 1
    torch.cuda.synchronize()
    dist.barrier()
    print(f"rank {rank}: ipc ok, got {x[:4].tolist()} on device {st.device}", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"rank {rank}: ipc FAILED: {type(e).__name__}: {e}", flush=True)
dist.destroy_process_group()
