#!/usr/bin/env python
"""Memory-bound entry points alone at the top-level shapes of a BASELINE config: CUDA events around each launch, L2
flushed between launches, achieved GB/s against the measured HBM peak (algorithmic bytes as in DESIGN.md 4.4).
    python tools/membound_bench.py [--config cfg3] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import costmodel as C, kernels as K  # noqa: E402
from saragan_b200.network import num_filters  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    cfg = C.CONFIGS[args.config]
    B, ph = cfg["batch"], cfg["phase"]
    d, h, w = C.volume(ph)
    V = d * h * w
    c_top = int(num_filters(ph, cfg["num_phases"], cfg["base_dim"]))          # D's top block input channels
    c_out = int(num_filters(ph - 1, cfg["num_phases"], cfg["base_dim"]))      # ... and its output channels
    peak = 6555.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    BF = torch.bfloat16
    dev = "cuda"
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    img = torch.randn(B, 1, d, h, w, device=dev)
    x_top = K.plain_to_act(torch.randn(B, c_top, d, h, w, device=dev), BF)
    y_top = K.plain_to_act(torch.randn(B, c_out, d, h, w, device=dev), BF)
    p_top = K.plain_to_act(torch.randn(B, c_out, d // 2, h // 2, w // 2, device=dev), BF)
    wv, bv = torch.randn(c_top, device=dev), torch.randn(c_top, device=dev)
    nparam = 32_000_000
    from saragan_b200.optim import FusedAdam
    prm = torch.nn.Parameter(torch.randn(nparam, device=dev))
    prm.grad = torch.randn(nparam, device=dev)
    opt = FusedAdam([prm], lr=1e-3)
    opt.step()
    e2 = 2
    cases = [
        ("pw_expand (FromRGB 1->%d)" % c_top, lambda: K.pw_expand(img, wv, bv, BF, c_top, 0.1, True), B * V * (4 + c_top * e2)),
        ("pw_expand_masked", lambda: K.pw_expand(img, wv, None, BF, c_top, 0.1, False, x_top), B * V * (4 + 2 * c_top * e2)),
        ("pw_reduce (ToRGB %d->1)" % c_top, lambda: K.pw_reduce(x_top, wv, bv[:1], c_top, 0.1), B * V * (c_top * e2 + 4)),
        ("pw_wgrad C=%d" % c_top, lambda: K.pw_wgrad(x_top, img, c_top, 0.1, True, True), B * V * (c_top * e2 + 4)),
        ("chan_sum C=%d" % c_out, lambda: K.pw_wgrad(y_top, None, c_out, 1.0, False, True), B * V * c_out * e2),
        ("down2 C=%d" % c_out, lambda: K.down2(y_top, 0.125), B * V * c_out * e2 * 9 // 8),
        ("up2+mask C=%d" % c_out, lambda: K.up2(p_top, 0.125, BF, y_top), B * V * c_out * e2 * (2 + 1 / 8)),
        ("up2 C=%d" % c_out, lambda: K.up2(p_top, 1.0, BF), B * V * c_out * e2 * (1 + 1 / 8)),
        ("mask_mul C=%d" % c_out, lambda: K.mask_mul(y_top, y_top), 3 * B * V * c_out * e2),
        ("lincomb C=%d" % c_out, lambda: K.lincomb(y_top, y_top, 0.5, 0.5), 3 * B * V * c_out * e2),
        ("pixelnorm_fwd C=%d" % c_top, lambda: K.pixelnorm_fwd(x_top, c_top, True), 2 * B * V * c_top * e2),
        ("pixelnorm_bwd C=%d" % c_top, lambda: K.pixelnorm_bwd(x_top, x_top, c_top, True, True), 3 * B * V * c_top * e2),
        ("adam 32M params", lambda: opt.step(), nparam * 20),
    ]
    print(f"{args.config}: B={B} top level {d}x{h}x{w}, D top block {c_top}->{c_out}; HBM peak {peak:.0f} GB/s")
    for name, fn, nbytes in cases:
        ts = []
        for i in range(args.reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        t = min(ts)
        print(f"{name:32s} {t * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / t / 1e6:7.0f} GB/s  {nbytes / t / 1e6 / peak:5.2f} of peak")


if __name__ == "__main__":
    main()
