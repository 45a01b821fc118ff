"""Probe (2+ GPUs, torchrun): which cross-process peer-memory mechanism works on this box --
torch symmetric memory (cuMem + fd exchange) and/or legacy CUDA IPC of caching-allocator tensors."""
import os
import sys
import time
import traceback

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
dist.barrier()


def log(*a):
    print(f"[rank{rank}]", *a, flush=True)


# ---- 1. symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    log("symm_mem ok: ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support(lr if False else 0, lr) if False else hdl.multicast_ptr)
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (1 << 20,), torch.float32)
    log("peer value", float(pt[123].item()), "expect", peer + 1)
    # bandwidth of a P2P read (torch copy kernel)
    big = symm_mem.empty(1 << 28, dtype=torch.float32, device=dev)  # 1 GiB
    h2 = symm_mem.rendezvous(big, dist.group.WORLD)
    src = h2.get_buffer(peer, (1 << 28,), torch.float32)
    dst = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    for _ in range(2):
        dst.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src)
    e1.record(); torch.cuda.synchronize()
    log("symm P2P read GB/s", 5 * (1 << 30) / 1e9 / (e0.elapsed_time(e1) / 1e3))
    # graph capture of barrier
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        hdl.barrier(channel=1)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g):
            hdl.barrier(channel=1)
            dst[:1 << 20].copy_(pt)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        log("symm barrier captured + replayed ok")
    except Exception as e:  # noqa: BLE001
        log("symm barrier capture FAILED", repr(e)[:300])
except Exception:  # noqa: BLE001
    log("symm_mem FAILED")
    traceback.print_exc()

dist.barrier()
# ---- 2. CUDA IPC of an ordinary caching-allocator tensor
try:
    from torch.multiprocessing.reductions import reduce_tensor
    x = torch.full((1 << 20,), float(10 + rank), device=dev)
    fn, args = reduce_tensor(x)
    objs = [None] * world
    dist.all_gather_object(objs, (args,))
    peer = (rank + 1) % world
    pargs = objs[peer][0]
    y = fn(*pargs)
    log("ipc tensor device", y.device, "value", float(y[5].item()), "expect", 10 + peer)
    z = y.to(dev)
    log("ipc copy to local ok", float(z[7].item()), "can_access_peer", torch.cuda.can_device_access_peer(lr, y.device.index))
except Exception:  # noqa: BLE001
    log("ipc FAILED")
    traceback.print_exc()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
