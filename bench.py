#!/usr/bin/env python
"""bench.py -- G+D train images/s of the 3D-PGAN step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one per-GPU batch: train.py:133-190 (one D update
with the WGAN-GP double backward + one G update, Adam applied).  Workload at every N: BASELINE
configs[2] = 3D PGAN 'small' final phase 32x128x128, per-GPU batch 4 (weak scaling; the
configuration the metric is quoted on, it fits one GPU).  Synthetic CT-like volumes, random-init
weights.

Prints ONE JSON line (rank 0).  `value` = images/s with the inputs already resident in HBM;
`e2e` = the same metric through the public API with host (pinned) inputs, H2D copy and a D2H
read of the losses inside the timed region; `roofline` = the dominant conv kernel timed with
CUDA events inside the timed region against the measured bf16 peak; `cpu_baseline` = the
oracle (port of the reference step) timed on this box's host cores.
`--impl reference` times the reference's CPU implementation of the path (oracle port; the
reference itself is Python/torch and /root/reference does not exist on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--network", default="network", choices=["network", "network_dict"],
                    help="network: the pgan_pytorch/network.py API (BASELINE north star); network_dict: the network_dict.py "
                         "variant main.py imports (LeakyReLU 0.3, He gain, no minibatch-stddev, top-level fade-in; SURVEY 8f row 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as CUDA graph(s) (on several GPUs: three "
                    "segments with the NCCL all-reduce between them); 0: eager with the bucketed, overlapped all-reduce")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm=p.get("hbm_gbs"),
                    src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(bf16=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def smooth_volumes(n, vol, seed):
    """Synthetic CT-like reals: clip(1024 + 350*smooth(N(0,1)), 0, 3072)/1024."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 1, *vol, generator=gen)
    k = torch.ones(1, 1, 3, 3, 3) / 27
    for _ in range(2):
        x = torch.nn.functional.conv3d(x, k, padding=1)
    x = x / x.std()
    return (torch.clamp(1024 + 350 * x, 0, 3072) / 1024).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self, t0=None, t1=None):
        """Samples that arrived inside the timed region [t0, t1]; a region shorter than nvidia-smi's sampling
        period may hold none, then the samples of the warm-up steps right before it (same load) are used."""
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        inside = [r for r in rows if t0 is not None and t0 <= r[-1] <= t1]
        window = "timed region"
        if not inside:
            inside, window = [r for r in rows if t1 is None or r[-1] <= t1][-5:], "warm-up + timed region (timed region shorter than the sampling period)"
        sm = [float(r[0]) for r in inside]
        mx = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i in range(4) if r[2 + i] == "Active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm), window=window)


def dict_oracle_state(cfg):
    """network_dict.py parameters (reference names and distributions: weight ~ N(0,1), bias ~ U(+-1/sqrt(fan_in)))
    for oracle/pgan_dict_oracle.py, laid out like the reference's state_dict at cfg's phase."""
    from saragan_b200 import costmodel as C
    from saragan_b200.network import num_filters
    f = lambda i: int(num_filters(i, cfg["num_phases"], cfg["base_dim"]))     # noqa: E731
    ph, bd, ld, vol0 = cfg["phase"], cfg["base_dim"], cfg["latent_dim"], int(np.prod(C.BASE_SHAPE[1:]))
    g = {"generator_in.0": (vol0 * bd, ld), "generator_in.3": (bd, bd, 3, 3, 3), "torgb_current.conv": (1, f(ph), 1, 1, 1)}
    d = {"fromrgb_current.fromrgb.0": (f(ph), 1, 1, 1, 1), "discriminator_out.0": (bd, bd, 3, 3, 3),
         "discriminator_out.3": (ld, vol0 * bd), "discriminator_out.5": (1, ld)}
    if ph > 1:
        g["torgb_prev.conv"] = (1, f(ph - 1), 1, 1, 1)
        d["fromrgb_prev.fromrgb.0"] = (f(ph - 1), 1, 1, 1, 1)
    for i in range(2, ph + 1):
        g[f"blocks.block_phase_{i}.conv1"] = (f(i), f(i - 1), 3, 3, 3)
        g[f"blocks.block_phase_{i}.conv2"] = (f(i), f(i), 3, 3, 3)
        d[f"blocks.block_phase_{i}.conv1"] = (f(i), f(i), 3, 3, 3)
        d[f"blocks.block_phase_{i}.conv2"] = (f(i - 1), f(i), 3, 3, 3)
    gen = torch.Generator().manual_seed(0)
    out = []
    for shapes in (g, d):
        p = {}
        for name, shp in shapes.items():
            w = torch.empty(shp).normal_(0, 1, generator=gen)
            bound = 1.0 / float(np.sqrt(w[0].numel()))
            p[name + ".weight"], p[name + ".bias"] = w, torch.empty(shp[0]).uniform_(-bound, bound, generator=gen)
        out.append(p)
    return out


def cpu_step_rate(cfg, batch, steps, warmup, alpha, threads=None, network="network"):
    """images/s of the oracle (CPU restatement of the reference step) on the host cores."""
    from oracle import pgan_oracle as O
    from saragan_b200 import costmodel as C
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    if network == "network_dict":
        from oracle import pgan_dict_oracle as OD
        pg, pd = dict_oracle_state(cfg)
        st = OD.DictTrainState(pg, pd, cfg["phase"], "leaky_relu", 0.3)
    else:
        gen = torch.Generator().manual_seed(0)
        pg = O.init_params("g", cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE, gen)
        pd = O.init_params("d", cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE, gen)
        st = O.TrainState(pg, pd, cfg["phase"], cfg["num_phases"])
    vol = C.volume(cfg["phase"])
    x = smooth_volumes(batch, vol, seed=1)
    times = []
    for i in range(warmup + steps):
        noise, z1, z2 = torch.randn_like(x), torch.randn(batch, cfg["latent_dim"]), torch.randn(batch, cfg["latent_dim"])
        eps = torch.rand(batch, 1, 1, 1, 1)
        t0 = time.perf_counter()
        st.step(x, noise, z1, eps, z2, alpha)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times[warmup:]))
    return batch / t, t, threads


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    b = args.cpu_batch
    rate, t, threads = cpu_step_rate(cfg, b, args.steps, args.warmup, args.alpha, network=args.network)
    sample = (f"{args.steps} full train steps (after {args.warmup} warm-up) at batch {b} of {args.config} "
              f"({args.network}.py), fp32, torch CPU")
    line = {"impl": "reference", "metric": "G+D train images/s", "value": rate, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "per_gpu_batch": b, "name": args.config, "network": args.network + ".py"},
            "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    from saragan_b200 import costmodel as C
    cfg = dict(C.CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch.distributed as dist
    import saragan_b200 as sg
    from saragan_b200 import _lib, comm, kernels
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sg.set_precision(args.precision)
    B, alpha = cfg["batch"], args.alpha
    vol = C.volume(cfg["phase"])

    torch.manual_seed(0)
    if args.network == "network_dict":
        from saragan_b200 import network_dict as nd
        margs = (cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE, "leaky_relu")
        g, d = nd.Generator(*margs, param=0.3), nd.Discriminator(*margs, param=0.3)     # main.py:203-204 defaults
    else:
        g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
        d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    use_graph = args.graph == 1
    if use_graph:
        from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers
        g_opt, d_opt = make_capturable_optimizers(g, d, world_size=world)
    else:
        from saragan_b200.graph import make_capturable_optimizers
        g_opt, d_opt = make_capturable_optimizers(g, d, world_size=world)      # fused Adam in eager mode too
    if world > 1:
        dp = comm.FlatAllReduce(g, d) if use_graph else comm.DataParallel(g, d)
    else:
        dp = None

    n_pool = 4
    host_pool = [smooth_volumes(B, vol, seed=1234 + 17 * rank + i).pin_memory() for i in range(n_pool)]
    dev_pool = [x.to(dev) for x in host_pool]
    rng = torch.Generator(device=dev).manual_seed(1000 + rank)

    def draws():
        return dict(noise=torch.randn((B, 1, *vol), device=dev, generator=rng),
                    z_d=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                    z_g=torch.randn((B, cfg["latent_dim"]), device=dev, generator=rng),
                    eps=torch.rand((B, 1, 1, 1, 1), device=dev, generator=rng))

    # nvidia-smi needs a few hundred ms before its first sample: start it before the capture's own warm-up steps
    clocks = ClockSampler(local)
    clocks.__enter__()
    if use_graph:
        graphed = GraphedTrainStep(g, d, g_opt, d_opt, B, vol, alpha, warmup=2, seed=1000 + rank, grad_sync=dp)

        def step(x):
            return graphed(x)
    else:
        def step(x):
            return sg.train_step(x, g, d, g_opt, d_opt, alpha, grad_sync=dp, **draws())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernels reported against the roofline: the top D block's convolutions (the launch shapes with
    # the largest time shares in profiles/): wgrad of conv2 (k_wgrad_tc is the kernel with the largest
    # share of the step), fprop of conv2, and conv2's dgrad
    layers = C.conv_layers("d", cfg["phase"], cfg["num_phases"], cfg["base_dim"])
    name, ci, co, v = max(layers, key=lambda l: l[1] * l[2] * int(np.prod(l[3])))
    flops_per_launch = 2.0 * B * int(np.prod(v)) * ci * co * 27
    probe_keys = {"wgrad": ("wgrad", B, ci, co, *v), "fprop": ("fprop", B, ci, co, *v), "dgrad": ("fprop", B, co, ci, *v)}
    probe = kernels.ConvProbe(probe_keys.values())
    # dram__bytes_read+write per launch from the committed `ncu --set full` capture (profiles/), cfg3 only
    ncu_traffic = {"wgrad": 407.1e6, "fprop": 349.9e6, "dgrad": 381.5e6} if (args.config == "cfg3" and B == 4) else {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for i in range(args.warmup):
            step(dev_pool[i % n_pool])
        barrier()
        kernels.conv_probe = probe
        launches0 = _lib.launch_count()
        t_region0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            out = step(dev_pool[i % n_pool])
        e1.record()
        barrier()
        t_region1 = time.perf_counter()
    finally:
        clocks.__exit__(None, None, None)
    if os.environ.get("SARAGAN_BENCH_DEBUG") and rank == 0:
        print("debug: losses after the timed region", [float(out[k]) for k in ("d_loss", "g_loss", "gp")], file=sys.stderr)
    kernels.conv_probe = None
    if use_graph:
        # a graph replay cannot carry per-kernel events: time the dominant kernel in two eager
        # passes of the same step (same shapes, same in-step cache state) right after the timed region
        kernels.conv_probe = probe
        for i in range(2):
            # single-stream order: with the gradient-penalty chain on its second stream the events around a launch
            # would also time whatever the other stream runs meanwhile
            sg.train_step(dev_pool[i % n_pool], g, d, g_opt, d_opt, alpha, grad_sync=dp, overlap_gp=False, **draws())
        barrier()
        kernels.conv_probe = None
        if os.environ.get("SARAGAN_BENCH_DEBUG") and rank == 0:
            print("debug: eager probe steps done", file=sys.stderr)
    launches = graphed.launches_per_step * args.steps if use_graph else _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    value = world * B / (ms_step * 1e-3)

    # end-to-end: pinned host batch -> H2D -> step -> D2H of the three losses, every step
    def e2e_step(i):
        # graph mode copies the pinned host batch straight into the graph's static input buffer
        return step(host_pool[i % n_pool] if use_graph else host_pool[i % n_pool].to(dev, non_blocking=True))

    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        o = e2e_step(i)
        losses = torch.stack([o["d_loss"], o["g_loss"], o["gp"]]).cpu()
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e = world * B / (float(ms2) / args.steps * 1e-3)

    if rank == 0:
        pk = peaks()
        step_flops = (C.step_flops_per_image_dict if args.network == "network_dict" else C.step_flops_per_image)(**cfg)
        line = {
            "metric": "G+D train images/s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "name": args.config, "network": args.network + ".py",
                       "per_gpu_batch": B, "global_batch": B * world,
                       "alpha": alpha, "parallelism": f"dp{world}",
                       "launch": ("cuda-graph replay of the whole step" if world == 1 else
                                  "4 cuda-graph segments + eager NCCL all-reduce (D-gradient all-reduce overlapped with the generator forward)") if use_graph else "eager",
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; 4 rotating input batches",
                       "step_gflop_per_image": step_flops / 1e9,
                       "step_tensor_frac": step_flops * value / world / (pk["bf16"] * 1e12)},
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * int(np.prod(vol)) * 4,
                    "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "cuda_core_conv_fallbacks_per_step": graphed.cuda_core_conv_fallbacks if use_graph else None,
            "clocks": clocks.summary(t_region0, t_region1),
            "losses": [float(v) for v in losses],
        }
        def roof(kind, label):
            kd_ = probe.durations_ms(probe_keys[kind])
            if not kd_:
                return None
            dur = float(np.mean(kd_)) * 1e-3
            ach = flops_per_launch / dur / 1e12
            return {"bound": "tensor", "kernel": label, "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16"], "traffic": ncu_traffic.get(kind), "launches_timed": len(kd_),
                    "ms_per_launch": dur * 1e3, "flop_per_launch": flops_per_launch, "peak_source": pk["src"]}
        shape = f"{ci}->{co} @{'x'.join(map(str, v))} B={B}"
        r = roof("wgrad", f"k_wgrad_tc: conv3d wgrad {name} {shape}")
        if r:
            line["roofline"] = r
            line["roofline_more"] = [x for x in (roof("fprop", f"k_conv_tc_res: conv3d fprop {name} {shape}"),
                                                 roof("dgrad", f"k_conv_tc_res: conv3d dgrad {name} {shape}")) if x]
        if world == 1 and not args.no_cpu_baseline:
            rate, t, threads = cpu_step_rate(cfg, args.cpu_batch, 1, 0, alpha, network=args.network)
            line["cpu_baseline"] = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                                    "sample": f"1 full train step at batch {args.cpu_batch} of {args.config} "
                                              f"({t:.1f} s), fp32 torch CPU oracle"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
