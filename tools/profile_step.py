#!/usr/bin/env python
"""Per-entry-point time breakdown of one train step, measured in situ with CUDA events around
every C-ABI call (no profiler attached).  Usage: python tools/profile_step.py [--config cfg3]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg  # noqa: E402
from saragan_b200 import _lib, costmodel as C  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--top", type=int, default=90)
    args = ap.parse_args()
    cfg = dict(C.CONFIGS[args.config])
    B = args.batch or cfg["batch"]
    vol = C.volume(cfg["phase"])
    torch.manual_seed(0)
    g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    g_opt, d_opt = sg.make_optimizers(g, d)
    x = torch.rand(B, 1, *vol, device="cuda") * 2
    for _ in range(2):
        sg.train_step(x, g, d, g_opt, d_opt, 0.5)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sg.train_step(x, g, d, g_opt, d_opt, 0.5, overlap_gp=False)   # one stream: events time one kernel at a time
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / args.steps
    agg = collections.defaultdict(lambda: [0, 0.0])
    byname = collections.defaultdict(lambda: [0, 0.0])
    for name, key, a, b, *_r in _lib.PROFILE:
        ms = a.elapsed_time(b) / args.steps
        agg[(name, key)][0] += 1
        agg[(name, key)][1] += ms
        byname[name][0] += 1
        byname[name][1] += ms
    _lib.PROFILE = None
    covered = sum(v[1] for v in byname.values())
    print(f"config {args.config} B={B}: {total:.2f} ms/step; ABI calls cover {covered:.2f} ms "
          f"({100 * covered / total:.1f} %); {len(agg)} distinct (entry point, shape) pairs")
    print("\n-- by entry point")
    for name, (n, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:9.3f} ms {100 * ms / total:5.1f} %  n/step={n / args.steps:6.1f}  {name}")
    print("\n-- by (entry point, integer args)")
    for (name, key), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"{ms:9.3f} ms {100 * ms / total:5.1f} %  n/step={n / args.steps:5.1f}  {ms / n * args.steps * 1e3:9.1f} us/call  {name} {key}")
    print("\n-- conv calls by shape: TFLOP/s and time above a 1100 TFLOP/s kernel")
    rows = []
    for (name, key), (n, ms) in agg.items():
        if name not in ("sg_conv3d_fprop", "sg_conv3d_wgrad"):
            continue
        _, nb, ci, co, d, h, w = key[:7]
        fl = 2.0 * nb * ci * co * 27 * d * h * w * n / args.steps
        rows.append((ms - fl / 1100e9, ms, fl, n / args.steps, name, key[:7]))
    for ex, ms, fl, n, name, key in sorted(rows, reverse=True):
        print(f"excess {ex:7.3f} ms  total {ms:7.3f} ms  n={n:4.1f}  {fl / ms / 1e9:7.1f} TF/s  {name[10:]} {key}")
    print(f"sum excess {sum(r[0] for r in rows):.3f} ms of {sum(r[1] for r in rows):.3f} ms conv time")
    print("\n-- memory-bound entry points: algorithmic bytes / event time (eager; small calls are launch-latency bound), "
          "peak = 6555.8 GB/s measured copy bandwidth")

    def pad(c):
        return 16 * ((c + 15) // 16)

    def esz(code):
        return 2 if code == 0 else 4

    def nbytes(name, k):
        if name == "sg_up2":          # dtype_in, dtype_out, N, C, D, H, W of the INPUT (mask_ref read not counted)
            v = k[2] * pad(k[3]) * k[4] * k[5] * k[6]
            return v * esz(k[0]) + 8 * v * esz(k[1])
        if name == "sg_down2":
            v = k[2] * pad(k[3]) * k[4] * k[5] * k[6]
            return v * esz(k[0]) + v // 8 * esz(k[1])
        if name == "sg_mask_mul":
            return 3 * k[1] * esz(k[0])
        if name == "sg_lincomb":      # b may be null: lower bound
            return 2 * k[1] * esz(k[0])
        if name == "sg_pw_expand":
            return k[1] * k[3] * 4 + k[1] * pad(k[2]) * k[3] * esz(k[0])
        if name in ("sg_pw_reduce", "sg_pw_wgrad"):
            return k[1] * k[3] * 4 + k[1] * pad(k[2]) * k[3] * esz(k[0])
        if name == "sg_pixelnorm_fwd":
            return 3 * k[1] * pad(k[2]) * k[3] * esz(k[0])
        if name == "sg_pixelnorm_bwd":
            return 5 * k[1] * pad(k[2]) * k[3] * esz(k[0])
        return None

    mem = []
    for (name, key), (n, ms) in agg.items():
        b = nbytes(name, key)
        if b is not None:
            mem.append((ms, n / args.steps, b, name, key))
    tot_ms = sum(m[0] for m in mem)
    tot_b = sum(m[2] * m[1] for m in mem)
    print(f"all of them: {tot_ms:.3f} ms/step for {tot_b / 1e9:.2f} GB/step = {tot_b / tot_ms / 1e6:.0f} GB/s average")
    for ms, n, b, name, key in sorted(mem, reverse=True)[:24]:
        print(f"{ms:8.3f} ms  n={n:4.1f}  {b / 1e6:8.1f} MB/call  {b * n / ms / 1e6:7.0f} GB/s  {name} {key}")


if __name__ == "__main__":
    main()
