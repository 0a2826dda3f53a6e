#!/usr/bin/env python3
"""Diagnostics: bandwidth of torch.distributed.all_to_all_single (NCCL send/recv) for the row-exchange sizes of the
multi-GPU sample sort.  torchrun --nproc-per-node N tools/bench_a2a.py [GB per rank]"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 11.3
n = int(gb * 1e9) // world * world
if os.environ.get("A2A_ARENA") == "1":               # buffers from the library's VMM arena, viewed through __cuda_array_interface__
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from uq_b200.device import Context
    ctx = Context(int(os.environ["LOCAL_RANK"]))
    a_send, a_recv = ctx.alloc(n, 1), ctx.alloc(n, 1)
    send = torch.as_tensor(a_send, device="cuda")
    recv = torch.as_tensor(a_recv, device="cuda")
else:
    send = torch.empty(n, dtype=torch.uint8, device="cuda")
    recv = torch.empty(n, dtype=torch.uint8, device="cuda")
split = [n // world] * world
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    dist.all_to_all_single(recv, send, split, split)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        out = n * (world - 1) / world
        print("a2a %.1f GB per rank, %d ranks: %.1f ms, %.0f GB/s leaving each rank (env %s)" % (
            n / 1e9, world, dt * 1e3, out / dt / 1e9, {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}), flush=True)
dist.destroy_process_group()
