"""Worker of tests/test_nccl_gpu.py (run under torch.distributed.run, one process per GPU, NCCL).

Checks, on a small PGAN (phase 3 of 4, 4x16x16, per-rank batch 4, different reals and draws per rank):
 1. ONE step's averaged gradients of the eager bucketed/overlapped path (comm.DataParallel) and of the flat all-reduce
    (comm.FlatAllReduce) equal the mean over ranks of the gradients each rank computes alone;
 2. K optimiser steps through the eager bucketed path, through the segmented CUDA-graph path (flat all-reduce between
    four graph segments) and through the SINGLE-graph path (gradient arena, NCCL captured inside the graph) leave the
    replicas BIT-IDENTICAL across ranks, and the paths agree with each other to the run-to-run noise of the step.
Prints one JSON line on rank 0; exit code 1 on any violation."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg  # noqa: E402
from saragan_b200 import comm  # noqa: E402
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers  # noqa: E402

CFG = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
VOL, B, ALPHA, K = (4, 16, 16), 4, 0.5, 5


def build():
    torch.manual_seed(5)
    g = sg.Generator(CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
    d = sg.Discriminator(CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
    return g, d


def draws(rank, i, dev):
    gen = torch.Generator().manual_seed(1000 * rank + i)
    return dict(noise=torch.randn(B, 1, *VOL, generator=gen).to(dev), z_d=torch.randn(B, 64, generator=gen).to(dev),
                z_g=torch.randn(B, 64, generator=gen).to(dev), eps=torch.rand(B, 1, 1, 1, 1, generator=gen).to(dev))


def reals(rank, i, dev):
    return torch.rand(B, 1, *VOL, generator=torch.Generator().manual_seed(77 * rank + i)).to(dev) * 2


DEBUG = bool(os.environ.get("MGPU_DEBUG"))


def _dbg(msg):
    if DEBUG:
        torch.cuda.synchronize()
        print(f"[dbg rank {os.environ.get('RANK')}] {msg}", flush=True)


def gather_equal(t):
    """True when every rank holds bit-identical `t`."""
    world = dist.get_world_size()
    _dbg("gather_equal: enter")
    buf = [torch.empty_like(t) for _ in range(world)]
    _dbg("gather_equal: buffers allocated")
    dist.all_gather(buf, t.contiguous())
    _dbg("gather_equal: all_gather done")
    return all(torch.equal(buf[0], b) for b in buf[1:])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    res, ok = {"world": world}, True
    if DEBUG:
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("MGPU_DUMP_AFTER", "90")), exit=True)

    # ---- 1. averaged gradients == mean of the per-rank gradients.  In the exact fp32 mode: two runs of the same step
    # then agree to the order of the weight-gradient atomics (~2e-6), so the 1e-4 bound tests the exchange and nothing
    # else (in the reduced-precision modes the GP double backward amplifies that noise, see tools/determinism_probe.py)
    sg.set_precision("fp32")
    g, d = build()
    g_opt, d_opt = sg.make_optimizers(g, d)
    sg.train_step(reals(rank, 0, dev), g, d, g_opt, d_opt, ALPHA, apply=False, **draws(rank, 0, dev))
    worst = {}
    for net, tag in (() if os.environ.get("MGPU_SKIP_PART1") else ((d, "d"), (g, "g"))):
        own = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        mean = {}
        for k, v in own.items():
            m = v.clone()
            dist.all_reduce(m)
            mean[k] = m / world
        for sync_name, make in (("bucketed", lambda a, b: comm.DataParallel(a, b, bucket_bytes=1 << 20)),
                                ("flat", comm.FlatAllReduce)):
            g2, d2 = build()
            o2 = sg.make_optimizers(g2, d2)
            sg.train_step(reals(rank, 0, dev), g2, d2, *o2, ALPHA, apply=False, grad_sync=make(g2, d2), **draws(rank, 0, dev))
            net2 = d2 if tag == "d" else g2
            err = max(float((p.grad - mean[k]).abs().max() / (mean[k].abs().max() + 1e-30))
                      for k, p in net2.named_parameters() if p.grad is not None)
            worst[f"{tag}.{sync_name}"] = err
            same = all(gather_equal(p.grad) for p in net2.parameters() if p.grad is not None)
            ok &= err < 1e-4 and same
            res[f"grads_identical_across_ranks.{tag}.{sync_name}"] = same
    res["avg_grad_vs_mean_of_rank_grads_max_rel"] = worst
    sg.set_precision("bf16")

    # ---- 2. K steps: eager bucketed vs segmented graph; replicas bit-identical
    finals = {}
    modes = os.environ.get("MGPU_MODES", "eager_bucketed,graph_segments,graph_arena").split(",")
    for mode in modes:
        g, d = build()
        g_opt, d_opt = make_capturable_optimizers(g, d, world_size=world)
        if mode == "eager_bucketed":
            dp = comm.DataParallel(g, d, bucket_bytes=1 << 20)
            for i in range(K):
                o = sg.train_step(reals(rank, i, dev), g, d, g_opt, d_opt, ALPHA, grad_sync=dp, **draws(rank, i, dev))
        else:
            dp = comm.FlatAllReduce(g, d) if mode == "graph_segments" else comm.ArenaAllReduce(g, d)
            graphed = GraphedTrainStep(g, d, g_opt, d_opt, B, VOL, ALPHA, warmup=2, seed=1, grad_sync=dp)
            _dbg(f"{mode}: captured")
            for i in range(K):
                _dbg(f"{mode}: replay {i}")
                dr = draws(rank, i, dev)
                graphed.draw = lambda dr=dr: [getattr(graphed, k).copy_(v) for k, v in dr.items()]   # replay the same draws
                o = graphed(reals(rank, i, dev))
            o = {k: v.clone() for k, v in o.items()}
            graphed.close()
        torch.cuda.synchronize()
        _dbg(f"{mode}: steps done")
        if DEBUG:
            tt = torch.ones(4, device=dev)
            dist.all_reduce(tt)
            _dbg(f"{mode}: small eager all_reduce ok {tt.tolist()}")
        flat = torch.cat([p.detach().reshape(-1) for p in list(g.parameters()) + list(d.parameters())])
        _dbg(f"{mode}: flat built {flat.numel()}")
        same = gather_equal(flat)
        res[f"replicas_bit_identical.{mode}"] = same
        ok &= same
        finals[mode] = flat
        res[f"losses.{mode}"] = [float(o[k]) for k in ("d_loss", "gp", "g_loss")]
    for other in [m for m in modes if m != "eager_bucketed" and "eager_bucketed" in modes]:
        diff = float((finals["eager_bucketed"] - finals[other]).abs().mean())
        res[f"mean_abs_weight_diff_eager_vs_{other}"] = diff
        # Adam with beta1 = 0 turns rounding-level gradient differences into +-lr steps: after K steps the two paths
        # may differ by a fraction of K*lr per weight, never by more
        ok &= diff < 0.25 * K * 1e-3 * world ** 0.5
        for k in range(3):
            a, b = res["losses.eager_bucketed"][k], res[f"losses.{other}"][k]
            ok &= abs(a - b) < 2e-2 * max(1.0, abs(a))
    res["ok"] = bool(ok)
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
