#!/usr/bin/env python
"""Row 8f.4 measurement: get_metrics (KS + sliced Wasserstein) on the GPU against the numpy/scipy oracle on the host.

  python tools/metrics_bench.py [--batch 4] [--shape 32 128 128] [--cpu-sample]

Prints one JSON line: per-call time of saragan_b200.metrics.get_metrics with device random numbers (CUDA events,
after warm-up), per-kernel times from CUDA events around every ABI call, the achieved GB/s of the projection kernel
(algorithmic bytes = the directions matrix K x 128 x 4 + the descriptor matrix 2B x K x 4, read once) against the
measured HBM peak, and -- with --cpu-sample -- the oracle (the reference's algorithm, numpy/scipy) timed on the host
cores on a bounded sample (batch 2 at 16x64x64) next to the GPU on the same sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import _lib  # noqa: E402
from saragan_b200 import metrics as M  # noqa: E402


def volumes(b, shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    real = torch.nn.functional.avg_pool3d(torch.randn(b, 1, *shape, device="cuda", generator=g), 3, 1, 1) * 0.9
    fake = torch.randn(b, 1, *shape, device="cuda", generator=g) * 0.4 + 0.05
    return real, fake


def time_gpu(real, fake, reps):
    for _ in range(2):
        M.get_metrics(real, fake)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        M.get_metrics(real, fake)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--shape", type=int, nargs=3, default=[32, 128, 128])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-sample", action="store_true")
    args = ap.parse_args()
    real, fake = volumes(args.batch, args.shape, 0)
    ms = time_gpu(real, fake, args.reps)
    _lib.PROFILE = []
    M.get_metrics(real, fake)
    torch.cuda.synchronize()
    per = {}
    for name, _, a, b, *_r in _lib.PROFILE:
        t = per.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += a.elapsed_time(b)
    _lib.PROFILE = None
    k = 128 * args.batch * 243
    proj_bytes = (k * 128 + 2 * args.batch * k) * 4
    n_proj, t_proj = per["sg_swd_project"]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6555.8))
    out = {"metric": "get_metrics (KS + 3-level sliced Wasserstein) per call", "unit": "ms", "value": ms,
           "config": {"batch": args.batch, "shape": args.shape, "rng": "device"},
           "kernels_ms": {n: {"calls": c, "total_ms": round(t, 4)} for n, (c, t) in sorted(per.items())},
           "roofline": {"bound": "hbm", "kernel": "k_swd_project", "achieved": proj_bytes * n_proj / (t_proj * 1e-3) / 1e9,
                        "peak": hbm, "unit": "GB/s", "frac": proj_bytes * n_proj / (t_proj * 1e-3) / 1e9 / hbm,
                        "bytes_per_launch": proj_bytes, "ms_per_launch": t_proj / n_proj}}
    if args.cpu_sample:
        from oracle import metrics_oracle as O
        sb, sshape = 2, (16, 64, 64)
        r, f = volumes(sb, sshape, 1)
        gpu_ms = time_gpu(r, f, args.reps)
        rn, fn = r.cpu().numpy(), f.cpu().numpy()
        t0 = time.perf_counter()
        want = O.get_metrics(rn.copy(), fn.copy(), np.random.RandomState(5))
        cpu_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        got = M.get_metrics(r, f, rng=np.random.RandomState(5))
        torch.cuda.synchronize()
        host_rng_s = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_s * 1e3, "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"one get_metrics call, batch {sb}, {sshape}, numpy/scipy oracle",
                               "gpu_same_sample_ms": gpu_ms, "gpu_same_sample_host_rng_ms": host_rng_s * 1e3,
                               "parity_rel": {k_: abs(got[k_] - want[k_]) / max(abs(want[k_]), 1e-30) for k_ in want}}
    print(json.dumps(out, default=float))


if __name__ == "__main__":
    main()
