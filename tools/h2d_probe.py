#!/usr/bin/env python3
"""Host-to-device bandwidth with all ranks copying at once: ordinary pinned memory (torch pin_memory) against
write-combined pinned memory (cudaHostAlloc with cudaHostAllocWriteCombined), 2 GiB per rank -- what bounds the
end-to-end number of bench.py at 8 GPUs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py"""
import ctypes
import os

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 2 << 30
dst = torch.empty(NB, dtype=torch.uint8, device=dev)
rt = ctypes.CDLL("libcudart.so.12")


def timed(src_ptr, label, piece=NB):
    def copy():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for off in range(0, NB, piece):
            n = min(piece, NB - off)
            rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + off), ctypes.c_void_p(src_ptr + off),
                                    ctypes.c_size_t(n), 1, st)
            assert rc == 0, rc
    copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        copy()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 4], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        print(f"{label}: {ms:.1f} ms per 2 GiB per rank (max over {world} ranks) = {NB / ms / 1e6:.1f} GB/s per rank, "
              f"{world * NB / ms / 1e6:.0f} GB/s aggregate", flush=True)


pinned = torch.empty(NB, dtype=torch.uint8).pin_memory()
pinned.fill_(1)
timed(pinned.data_ptr(), "pinned (torch pin_memory)")
for mb in (256, 64, 16, 4):
    timed(pinned.data_ptr(), f"pinned, in pieces of {mb} MiB", mb << 20)
del pinned
p = ctypes.c_void_p()
rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), 4)  # cudaHostAllocWriteCombined
assert rc == 0, rc
ctypes.memset(p, 1, NB)
timed(p.value, "pinned + write-combined")
rt.cudaFreeHost(p)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
