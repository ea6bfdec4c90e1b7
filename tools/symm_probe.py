"""Probe: does torch symmetric memory (P2P-mapped peer buffers) work on this box, and how fast are copy-engine peer copies?
torchrun --nproc-per-node N tools/symm_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm

n = 50257 * 1024
buf = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok; multicast ptr:", hdl.multicast_ptr, "world", hdl.world_size, flush=True)
buf.fill_(float(rank + 1))
hdl.barrier(channel=0)
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (n,), torch.float32)
tmp = torch.empty(n // world, dtype=torch.float32, device=dev)
chunk = n // world
torch.cuda.synchronize()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tmp.copy_(pbuf[rank * chunk:(rank + 1) * chunk])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if rank == 0:
        print(f"pull {chunk * 4 / 1e6:.1f} MB from peer: {ms:.3f} ms = {chunk * 4 / ms / 1e6:.1f} GB/s", flush=True)
assert tmp[0].item() == float(peer + 1), tmp[0].item()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pbuf[rank * chunk:(rank + 1) * chunk].copy_(tmp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if rank == 0:
        print(f"push {chunk * 4 / 1e6:.1f} MB to peer: {ms:.3f} ms = {chunk * 4 / ms / 1e6:.1f} GB/s", flush=True)
t0 = time.perf_counter()
for _ in range(20):
    hdl.barrier(channel=0)
torch.cuda.synchronize()
if rank == 0:
    print(f"symm barrier: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
