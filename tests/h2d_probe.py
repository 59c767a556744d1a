"""Aggregate host->device bandwidth of N ranks copying at once (torchrun), for the e2e leg's scaling question:
default pinned memory (torch) vs write-combined pinned memory (cudaHostAlloc), plus what the box says about its topology."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

NBYTES = 68_813_312
N_SETS, ITERS = 4, 24


def wc_pinned(nbytes):
    rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    assert rc == 0, rc
    buf = (ctypes.c_float * (nbytes // 4)).from_address(ptr.value)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.float32))


def measure(name, bufs):
    dst = torch.empty(NBYTES // 4, device=dev)
    for b in bufs:
        dst.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(ITERS):
        dst.copy_(bufs[i % N_SETS], non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([NBYTES * ITERS / (s.elapsed_time(e) / 1e3) / 1e9], device=dev)
    if world > 1:
        all_ = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(all_, gbs)
    else:
        all_ = [gbs]
    if rank == 0:
        vals = [float(v) for v in all_]
        print(f"{name}: per rank {[round(v, 1) for v in vals]} GB/s, aggregate {sum(vals):.1f} GB/s", flush=True)


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["sh", "-c", "ls /sys/devices/system/node/ | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            print("\n".join(out.splitlines()[:24]), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(cmd, ex)
src = torch.randn(NBYTES // 4)
pinned = [src.clone().pin_memory() for _ in range(N_SETS)]
measure("torch pinned", pinned)
try:
    wc = []
    for _ in range(N_SETS):
        t = wc_pinned(NBYTES)
        t.copy_(src)
        wc.append(t)
    measure("write-combined pinned", wc)
except Exception as ex:  # noqa: BLE001
    if rank == 0:
        print("write-combined allocation failed:", ex)
measure("torch pinned (again)", pinned)
if world > 1:
    dist.destroy_process_group()
