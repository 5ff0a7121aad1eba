#!/usr/bin/env python
"""Multi-GPU consistency probe (run under torchrun): the same K steps through (a) the eager data-parallel step,
(b) the segmented CUDA-graph step with the flat all-reduce, (c) the segmented step WITHOUT gradient exchange.
Prints the losses per step and a parameter checksum per rank (replicas must stay identical in a, b).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_diag.py --config cfg1"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg  # noqa: E402
from saragan_b200 import comm, costmodel as C  # noqa: E402
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg1")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--bench-data", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl")
cfg = C.CONFIGS[args.config]
B, vol, alpha = cfg["batch"], C.volume(cfg["phase"]), 0.5


class NoExchange(comm.FlatAllReduce):
    def finish_tensors(self, grads):
        pass


def checksum(net):
    return float(sum(p.detach().double().abs().sum() for p in net.parameters()))


for mode in ("eager", "graph"):
    torch.manual_seed(0)
    g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    g_opt, d_opt = make_capturable_optimizers(g, d, world_size=world)
    xs = [torch.rand(B, 1, *vol, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * rank + i)) * 2
          for i in range(args.steps)]
    if args.bench_data:
        sys.path.insert(0, ROOT)
        import bench as BB
        xs = [BB.smooth_volumes(B, vol, seed=1234 + 17 * rank + (i % 4)).to(dev) for i in range(args.steps)]
    lines = []
    if mode.startswith("eager"):
        dp = (comm.FlatAllReduce(g, d) if mode == "eager_flat" else comm.DataParallel(g, d)) if world > 1 else None
        rng = torch.Generator(device=dev).manual_seed(1000 + rank)
        for i in range(args.steps + 2):
            x = xs[max(i - 2, 0)] if i >= 2 else torch.zeros_like(xs[0])
            dr = dict(noise=torch.randn((B, 1, *vol), device=dev, generator=rng),
                      z_d=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                      z_g=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                      eps=torch.rand((B, 1, 1, 1, 1), device=dev, generator=rng))
            o = sg.train_step(x, g, d, g_opt, d_opt, alpha, grad_sync=dp, **dr)
            lines.append((float(o["d_loss"]), float(o["gp"]), float(o["g_loss"])))
    else:
        dp = (NoExchange(g, d) if mode == "graph_noexchange" else comm.FlatAllReduce(g, d)) if world > 1 else None
        if mode == "graph_single_stream":
            import saragan_b200.train as T
            import saragan_b200.graph as GG
            _orig = T.d_phase
            GG.d_phase = lambda *a, **k: _orig(*a, **{**k, "overlap_gp": False})
        graphed = GraphedTrainStep(g, d, g_opt, d_opt, B, vol, alpha, warmup=2, seed=1000 + rank, grad_sync=dp)
        lines += [(float("nan"),) * 3] * 2
        for i in range(args.steps):
            if mode == "graph_then_eager_then_graph" and i == args.steps // 2:
                # what bench.py does after its timed region: two eager steps (per-kernel event timing), then replays again
                rng = torch.Generator(device=dev).manual_seed(77 + rank)
                for _ in range(2):
                    dr = dict(noise=torch.randn((B, 1, *vol), device=dev, generator=rng),
                              z_d=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                              z_g=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                              eps=torch.rand((B, 1, 1, 1, 1), device=dev, generator=rng))
                    o = sg.train_step(xs[i], g, d, g_opt, d_opt, alpha, grad_sync=dp, **dr)
                    lines.append((float(o["d_loss"]), float(o["gp"]), float(o["g_loss"])))
            if mode == "graph_hostinput":
                if i == 0:
                    hx = [x.cpu().pin_memory() for x in xs]
                o = graphed(hx[i])
            elif mode == "graph_devinput_repeat4":
                o = graphed(xs[i % 4])
            else:
                o = graphed(xs[i])
            lines.append((float(o["d_loss"]), float(o["gp"]), float(o["g_loss"])))
    torch.cuda.synchronize()
    cs = torch.tensor([checksum(g), checksum(d)], device=dev, dtype=torch.float64)
    allcs = [torch.zeros_like(cs) for _ in range(world)]
    if world > 1:
        dist.all_gather(allcs, cs)
    else:
        allcs = [cs]
    if rank == 0:
        print(f"== {mode}: (d_loss, gp, g_loss) per step incl. 2 warm-up steps")
        for i, l in enumerate(lines):
            print(f"   step {i:2d}  {l[0]:10.5f} {l[1]:10.5f} {l[2]:10.5f}")
        print("   param checksums (G, D) per rank:", [[round(float(v), 3) for v in c] for c in allcs])
if world > 1:
    dist.destroy_process_group()
