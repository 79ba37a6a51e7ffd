#!/usr/bin/env python
"""Aggregate device-to-host bandwidth of N ranks draining at once, for pinned host memory allocated three ways:
cudaHostAlloc (torch pin_memory), 4 KiB pages + cudaHostRegister, and 2 MiB transparent huge pages + cudaHostRegister.
usage: [torchrun ...] d2h_probe.py [MiB per copy]"""
import ctypes, mmap, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev.random_(0, 255)
rt = torch.cuda.cudart()
libc = ctypes.CDLL("libc.so.6", use_errno=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def registered(huge):
    mm = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    buf = (ctypes.c_char * (n + (2 << 20))).from_buffer(mm)
    addr = (ctypes.addressof(buf) + (2 << 20) - 1) & ~((2 << 20) - 1)
    if huge:
        libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(n), 14)  # MADV_HUGEPAGE
    else:
        libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(n), 15)  # MADV_NOHUGEPAGE
    arr = np.frombuffer((ctypes.c_uint8 * n).from_address(addr), dtype=np.uint8)
    arr[:] = 0   # first touch
    r = rt.cudaHostRegister(addr, n, 0)
    assert int(r) == 0, r
    t = torch.from_numpy(arr)
    return t, (mm, buf, addr)


def run(name, host):
    for _ in range(3):
        host.copy_(dev, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        host.copy_(dev, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    gbs = reps * n / dt / 1e9
    if world > 1:
        t = torch.tensor([gbs], device="cuda", dtype=torch.float64); dist.all_reduce(t); tot = float(t.item())
    else:
        tot = gbs
    if rank == 0:
        print(f"{name:32s} per rank {gbs:6.1f} GB/s   all {world} ranks {tot:7.1f} GB/s", flush=True)


thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
if rank == 0:
    print("THP:", thp, " copy size", mib, "MiB")
run("cudaHostAlloc (torch pin_memory)", torch.empty(n, dtype=torch.uint8, pin_memory=True))
h4, keep4 = registered(False); run("4 KiB pages + cudaHostRegister", h4)
h2, keep2 = registered(True); run("2 MiB THP + cudaHostRegister", h2)
if world > 1:
    dist.destroy_process_group()
