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
`e2e` = the same metric through the public API from HOST memory: the raw uint16 volumes in pinned
memory -> H2D -> sg_prepare_real (cast, /1024) -> step -> D2H read of the losses, every step;
`roofline` = the kernel (entry point + shape) with the largest share of the step, timed with CUDA
events on its launching stream in two EAGER passes of the same step right after the timed region (a
graph replay cannot carry per-kernel events), against the measured BURST peaks (a single kernel timed
alone between events); `roofline_more` = the other convolution variants of that layer and the
memory-bound kernels with the largest shares against the measured HBM peak; `traffic` = dram bytes
of that launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json, null if the
launch is not in it); `parity_check` = one step on seeded inputs against the stored fp32 oracle
losses (tests/golden/fullsize_*.json, oracle/pin_fullsize.py) -- the run aborts if it fails;
`cpu_baseline` = the oracle (port of the reference step) timed on this box's host cores.
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
    ap.add_argument("--phase", type=int, default=0, help="growth phase (default: the config's top phase); the per-phase "
                                                         "table of a progression (cfg2) is one run per phase")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--network", default="network", choices=["network", "network_dict"],
                    help="network: the pgan_pytorch/network.py API (BASELINE north star); network_dict: the network_dict.py "
                         "variant main.py imports (LeakyReLU 0.3, He gain, no minibatch-stddev, top-level fade-in; SURVEY 8f row 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU arm (default: the per-GPU batch)")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as CUDA graph(s) (on several GPUs: three "
                    "segments with the NCCL all-reduce between them); 0: eager with the bucketed, overlapped all-reduce")
    return ap.parse_args()


def peaks():
    """Measured peaks (driver-written MEASURED_PEAKS.json): burst = a kernel timed alone, sustained = inside a long
    step; fallback = the figures of B200_PROFILING.md."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p.get("bf16_tflops"), bf16_sustained=p.get("bf16_tflops_sustained", p.get("bf16_tflops")),
                    hbm=p.get("hbm_gbs"), src="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def smooth_volumes(n, vol, seed):
    from saragan_b200.data import synthetic_reals
    return synthetic_reals(n, vol, seed)


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
    b = args.cpu_batch or cfg["batch"]
    rate, t, threads = cpu_step_rate(cfg, b, args.steps, args.warmup, args.alpha, network=args.network)
    sample = (f"{args.steps} full train steps (after {args.warmup} warm-up) at batch {b} of {args.config} "
              f"({args.network}.py), fp32, torch CPU")
    line = {"impl": "reference", "metric": "G+D train images/s", "value": rate, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"].rsplit(", B=", 1)[0] + f", B={b}", "per_gpu_batch": b, "name": args.config,
                       "network": args.network + ".py"},
            "cpu_baseline": {"value": rate, "unit": "img/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def parity_check(args, cfg, g, d, g_opt, d_opt, dev):
    """One train step (no optimiser update) on the seeded inputs of tests/golden/fullsize_<cfg>_b<B>.json against the
    fp32 oracle losses stored there (oracle/pin_fullsize.py).  Aborts the run when the CUDA path disagrees."""
    import saragan_b200 as sg
    from saragan_b200 import costmodel as C
    from saragan_b200.data import step_draws, synthetic_reals
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_{args.config}_b{cfg['batch']}.json")
    if args.network != "network" or args.phase or not os.path.exists(path):
        return None
    with open(path) as f:
        ref = json.load(f)
    vol = C.volume(cfg["phase"])
    x = synthetic_reals(cfg["batch"], vol, ref["x_seed"])
    dr = step_draws(cfg["batch"], vol, cfg["latent_dim"], ref["draw_seed"])
    out = sg.train_step(x, g, d, g_opt, d_opt, ref["alpha"], apply=False, **dr)
    got = {k: float(out[k]) for k in ("d_loss", "gp", "g_loss")}
    tol = {"bf16": 2e-3, "tf32": 1e-3, "fp32": 1e-4}[args.precision]
    err = {k: abs(got[k] - ref["losses"][k]) / max(abs(ref["losses"][k]), 1.0 if k == "g_loss" else 1e-12) for k in got}
    for net in (g, d):
        for p in net.parameters():
            p.grad = None
    res = {"fixture": os.path.relpath(path, ROOT), "cuda": got, "oracle": ref["losses"], "rel_err": err, "tol": tol,
           "ok": all(e < tol for e in err.values())}
    if not res["ok"]:
        print(json.dumps({"parity_check": res}), file=sys.stderr, flush=True)
        raise SystemExit("bench.py: the CUDA step disagrees with the stored oracle losses -- not benchmarking a wrong kernel")
    return res


def _dt(code):
    return 2 if code == 0 else 4       # element size of SG_BF16 / SG_F32


def launch_work(name, ints, flags):
    """(bound, algorithmic flops or bytes, label) of one ABI call from its integer arguments (include/saragan_b200.h;
    DESIGN.md lists the per-unit figures), or None for calls that are not reported."""
    if name == "sg_conv3d_fprop":
        dt, n, ci, co, d, h, w = ints[:7]
        return "tensor", 2.0 * n * d * h * w * ci * co * 27, f"conv3d fprop/dgrad {ci}->{co} @{d}x{h}x{w} B={n} ({'bf16' if dt == 0 else 'fp32 storage'})"
    if name == "sg_conv3d_wgrad":
        dt, n, ci, co, d, h, w = ints[:7]
        return "tensor", 2.0 * n * d * h * w * ci * co * 27, f"conv3d wgrad {ci}->{co} @{d}x{h}x{w} B={n} ({'bf16' if dt == 0 else 'fp32 storage'})"
    if name == "sg_up2":
        di, do, vec, p, d, h, w = ints[:7]
        v = p * d * h * w * vec
        has_mask = len(flags) > 2 and flags[2]
        return "hbm", v * _dt(di) + 8 * v * _dt(do) * (2 if has_mask else 1), f"up2{'+mask' if has_mask else ''} [{p}]x{d}x{h}x{w}x{vec}"
    if name == "sg_down2":
        di, do, vec, p, d, h, w = ints[:7]
        v = p * d * h * w * vec
        return "hbm", v * _dt(di) + v // 8 * _dt(do), f"down2 [{p}]x{d}x{h}x{w}x{vec}"
    if name in ("sg_pw_expand", "sg_pw_expand_masked"):
        dt, n, c, v = ints[:4]
        cp = 16 * ((c + 15) // 16)
        return "hbm", n * v * (4 + cp * _dt(dt) * (2 if name.endswith("masked") else 1)), f"FromRGB 1->{c} B={n} V={v}"
    if name == "sg_pw_reduce":
        dt, n, c, v = ints[:4]
        return "hbm", n * v * (16 * ((c + 15) // 16) * _dt(dt) + 4), f"ToRGB {c}->1 B={n} V={v}"
    if name == "sg_pw_wgrad":
        dt, n, c, v = ints[:4]
        return "hbm", n * v * (16 * ((c + 15) // 16) * _dt(dt) + (4 if flags[1] else 0)), f"1x1x1 wgrad C={c} B={n} V={v}"
    if name == "sg_mask_mul":
        dt, n = ints[:2]
        return "hbm", 3 * n * _dt(dt), f"mask_mul n={n}"
    if name == "sg_lincomb":
        dt, n = ints[:2]
        return "hbm", (3 if flags[1] else 2) * n * _dt(dt), f"lincomb n={n}"
    if name in ("sg_pixelnorm_fwd", "sg_pixelnorm_bwd"):
        dt, n, c, v = ints[:4]
        cp = 16 * ((c + 15) // 16)
        return "hbm", n * v * cp * _dt(dt) * (2 if name.endswith("fwd") else 3), f"{name[3:]} C={c} B={n} V={v}"
    return None


def roofline_entries(table, pk, n_passes):
    """table: (entry point, int args, pointer flags) -> [ms per launch].  First entry = the launch shape with the
    largest time share of the step (always a convolution), then the other conv variants with the next shares and the
    memory-bound kernels with the largest shares."""
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    rows = []
    for (name, ints, flags), ms in table.items():
        wk = launch_work(name, ints, flags)
        if wk is None:
            continue
        bound, work, label = wk
        dur = float(np.mean(ms)) * 1e-3
        peak = pk["bf16_burst"] if bound == "tensor" else pk["hbm"]
        ach = work / dur / (1e12 if bound == "tensor" else 1e9)
        key = name + ("+mask" if name == "sg_up2" and len(flags) > 2 and flags[2] else "") + ":" + ",".join(map(str, ints[:7]))
        rows.append({"bound": bound, "kernel": f"{name}: {label}", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": ach / peak,
                     "traffic": traffic.get(key, {}).get("dram_bytes"), "launches_per_step": len(ms) / n_passes,
                     "ms_per_launch": dur * 1e3, "ms_per_step": float(np.sum(ms)) / n_passes,
                     ("flop_per_launch" if bound == "tensor" else "bytes_per_launch"): work,
                     "peak_source": pk["src"] + (", burst bf16 (kernel timed alone between events)" if bound == "tensor" else ", copy bandwidth"),
                     "timed": "CUDA events on the launching stream, eager single-stream passes after the timed region"})
    rows.sort(key=lambda r: -r["ms_per_step"])
    conv = [r for r in rows if r["bound"] == "tensor"]
    # bandwidth entries: launches that move >= 32 MB (a smaller launch is bound by launch latency, and in these eager
    # passes its event-to-event time also holds the host's call overhead)
    mem = [r for r in rows if r["bound"] == "hbm" and r["bytes_per_launch"] >= 32e6]
    return conv[:4] + mem[:6]


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    from saragan_b200 import costmodel as C
    cfg = dict(C.CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    if args.phase:
        cfg["phase"] = args.phase
        if not args.batch:
            # the reference's per-phase batch rule, main.py:99 (128 // resolution), floored at 2 (a batch of 1
            # zeroes the minibatch-stddev feature): 32, 16, 8, 4, 2 for the resolutions 4 ... 64 of the xs progression
            cfg["batch"] = max(2, 128 // (C.BASE_SHAPE[2] * 2 ** (args.phase - 1)))
        cfg["desc"] = cfg["desc"].replace(", B=", f" [run at phase {args.phase}: {'x'.join(map(str, C.volume(args.phase)))}], B=")
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch.distributed as dist
    import saragan_b200 as sg
    from saragan_b200 import _lib, comm, data
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
    ap_exchange = os.environ.get("SARAGAN_EXCHANGE", "arena")     # arena | flat (4 graph segments + eager NCCL) -- A/B switch
    if world > 1 or os.environ.get("SARAGAN_FORCE_ARENA") == "1":     # (the switch: the N > 1 step structure on one GPU, for profiling)
        dp = (comm.ArenaAllReduce(g, d) if ap_exchange == "arena" else comm.FlatAllReduce(g, d)) if use_graph else comm.DataParallel(g, d)
    else:
        dp = None

    # ---- parity gate: one step (no optimiser update) on seeded inputs against the stored fp32 oracle losses
    parity = parity_check(args, cfg, g, d, g_opt, d_opt, dev) if rank == 0 else None

    n_pool = 4
    host_pool = [smooth_volumes(B, vol, seed=1234 + 17 * rank + i) for i in range(n_pool)]
    # the dataset's raw form: uint16 voxels (create_lidc_idri_dataset.py:185-212), value*1024 -- what e2e feeds
    host_raw = [(x * 1024).round().to(torch.uint16).pin_memory() for x in host_pool]
    dev_pool = [r.to(dev).float().div_(1024).reshape(B, 1, *vol) for r in host_raw]
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
            if x.dtype == torch.uint16:
                x = data.prepare_real(x.to(dev, non_blocking=True), None)
            return sg.train_step(x, g, d, g_opt, d_opt, alpha, grad_sync=dp, **draws())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for i in range(args.warmup):
            step(dev_pool[i % n_pool])
        barrier()
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
    launches = graphed.launches_per_step * args.steps if use_graph else _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end-to-end: pinned uint16 host batch -> H2D -> sg_prepare_real -> step -> D2H of the three losses, every step
    for i in range(2):
        step(host_raw[i % n_pool])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        o = step(host_raw[i % n_pool])
        losses = torch.stack([o["d_loss"], o["g_loss"], o["gp"]]).cpu()
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e = world * B / (float(ms2) / args.steps * 1e-3)

    # ---- per-kernel roofline: two eager, single-stream passes of the same step (same shapes, same in-step cache
    # state) with CUDA events around every ABI call on its launching stream (with the gradient-penalty chain on its
    # second stream the events around a launch would also time whatever the other stream runs meanwhile)
    table = {}
    if rank == 0 or world > 1:
        _lib.PROFILE = []
        for i in range(2):
            sg.train_step(dev_pool[i % n_pool], g, d, g_opt, d_opt, alpha, grad_sync=dp, overlap_gp=False, **draws())
        barrier()
        prof, _lib.PROFILE = _lib.PROFILE, None
        for name, ints, a, b, flags in prof:
            table.setdefault((name, ints, flags), []).append(a.elapsed_time(b))

    if rank == 0:
        pk = peaks()
        step_flops = (C.step_flops_per_image_dict if args.network == "network_dict" else C.step_flops_per_image)(**cfg)
        line = {
            "metric": "G+D train images/s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": cfg["desc"].rsplit(", B=", 1)[0] + f", B={B}", "name": args.config,
                       "network": args.network + ".py", "per_gpu_batch": B, "global_batch": B * world,
                       "alpha": alpha, "parallelism": f"dp{world}",
                       "precision_policy": sg.config.describe(),
                       "launch": ("cuda-graph replay of the whole step" if world == 1 else graphed.launch_desc) if use_graph else "eager",
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; 4 rotating input batches",
                       "step_gflop_per_image": step_flops / 1e9,
                       "step_tensor_frac_of_burst": step_flops * value / world / (pk["bf16_burst"] * 1e12),
                       "step_tensor_frac_of_sustained": step_flops * value / world / (pk["bf16_sustained"] * 1e12)},
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * int(np.prod(vol)) * 2,
                    "d2h_bytes_per_step": 12, "input": "raw uint16 volumes in pinned host memory (data.prepare_real on the device)"},
            "gpu_launches": int(launches),
            "cuda_core_conv_fallbacks_per_step": graphed.cuda_core_conv_fallbacks if use_graph else None,
            "clocks": clocks.summary(t_region0, t_region1),
            "losses": [float(v) for v in losses],
            "parity_check": parity,
        }
        roofs = roofline_entries(table, pk, n_passes=2)
        if roofs:
            line["roofline"] = roofs[0]
            line["roofline_more"] = roofs[1:]
        if world == 1 and not args.no_cpu_baseline:
            cb = args.cpu_batch or B
            rate, t, threads = cpu_step_rate(cfg, cb, 1, 0, alpha, network=args.network)
            line["cpu_baseline"] = {"value": rate, "unit": "img/s", "cores": threads, "kind": "port",
                                    "sample": f"1 full train step at batch {cb} of {args.config} "
                                              f"({t:.1f} s), fp32 torch CPU oracle"}
        print(json.dumps(line), flush=True)
    if world > 1:
        if use_graph:
            graphed.close()                    # NCCL work captured in the graph must go before the communicator does
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
