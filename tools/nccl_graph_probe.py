#!/usr/bin/env python
"""Probe (2+ GPUs, torchrun): can an NCCL all-reduce captured in a CUDA graph be replayed and then followed by eager
collectives -- on the same process group, on a dedicated one?  Prints one line per variant."""
import os
import sys
import traceback

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
extra = dist.new_group(backend="nccl")


def variant2(name, n):
    """two captured all-reduces (one on a forked side stream, one on the capture stream) of n floats each, like the
    train step's D / G gradient arenas, then eager collectives"""
    try:
        a, b = torch.ones(n, device=dev), torch.ones(n, device=dev)
        dist.all_reduce(a)
        dist.all_reduce(b)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            a.mul_(0.5)
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dist.all_reduce(a)
            b.mul_(0.5)
            main.wait_stream(side)
            a.add_(1.0)
            dist.all_reduce(b)
            b.add_(1.0)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        t = torch.tensor([float(rank)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bufs = [torch.empty_like(a[:1000]) for _ in range(dist.get_world_size())]
        dist.all_gather(bufs, a[:1000].contiguous())
        torch.cuda.synchronize()
        if rank == 0:
            print(f"PROBE {name}: OK", flush=True)
    except Exception as e:      # noqa: BLE001
        if rank == 0:
            print(f"PROBE {name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
        sys.exit(1)


def variant(name, group, eager_group, side_stream):
    try:
        buf = torch.ones(1 << 20, device=dev) * (rank + 1)
        dist.all_reduce(buf, group=group)                     # communicator warm-up, eager
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        static = torch.ones(1 << 20, device=dev)
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            static.mul_(2.0)
            if side_stream:
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    dist.all_reduce(static, group=group)
                main.wait_stream(side)
            else:
                dist.all_reduce(static, group=group)
            static.add_(1.0)
        for _ in range(3):
            static.fill_(1.0)
            g.replay()
        torch.cuda.synchronize()
        v = float(static[0])
        t = torch.tensor([float(rank)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=eager_group)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"PROBE {name}: OK replay value {v} (expect {2.0 * dist.get_world_size() + 1.0}), eager max {float(t)}", flush=True)
    except Exception as e:      # noqa: BLE001
        if rank == 0:
            print(f"PROBE {name}: FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
            traceback.print_exc()
        sys.exit(1)


which = sys.argv[1]
if which == "same":
    variant("captured + eager on the default group", None, None, False)
elif which == "same_side":
    variant("captured (side stream) + eager on the default group", None, None, True)
elif which == "two_small":
    variant2("two captured all-reduces, 1M floats", 1 << 20)
elif which == "two_large":
    variant2("two captured all-reduces, 32M floats", 1 << 25)
elif which == "dedicated":
    variant("captured on a dedicated group, eager on the default group", extra, None, True)
dist.destroy_process_group()
