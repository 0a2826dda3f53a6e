#!/usr/bin/env python3
"""Diagnostics: can this box give peer-mapped (symmetric) device memory to one-process-per-GPU ranks?  Allocates a symmetric
buffer, maps the peers', lets every rank write a pattern straight into its right neighbour's buffer and checks it."""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
size = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
t0 = time.perf_counter()
buf = symm_mem.empty(size, dtype=torch.uint8, device="cuda:%d" % rank)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
dt = time.perf_counter() - t0
ptrs = list(hdl.buffer_ptrs)
print("rank", rank, "rendezvous %.2f s" % dt, "ptrs", [hex(p) for p in ptrs], "multicast", hdl.has_multicast_support, flush=True)
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (size,), torch.uint8)
src = torch.full((size,), rank + 1, dtype=torch.uint8, device="cuda")
hdl.barrier(channel=0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
remote.copy_(src)
ev1.record()
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = bool((buf == ((rank - 1) % world) + 1).all().item())
print("rank", rank, "peer write ok:", ok, "%.1f GB/s" % (size / 1e9 / (ev0.elapsed_time(ev1) / 1e3)), flush=True)
dist.destroy_process_group()
