#!/usr/bin/env python
"""Repeat the fp32 golden step with the caching allocator poisoned (NaN-filled blocks freed before every
run) and print the worst gradient per run: exposes reads of uninitialised scratch.  python tools/flaky_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg  # noqa: E402
from tests.util import build_pair, golden_tensors, load_golden, rel_err, run_step  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_p3"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
z, cfg = load_golden(name)
inp = {k: torch.from_numpy(z["in." + k]) for k in ("x_real", "noise", "z_d", "z_g", "eps")}
for it in range(30):
    junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22)
            for _ in range(4)]
    del junk
    with sg.use_precision(prec):
        g, d = build_pair(cfg)
        out = run_step(g, d, inp, cfg["alpha"])
    worst = ("", 0.0)
    for kind, mod in (("d_grads", d), ("g_grads", g)):
        want = golden_tensors(z, f"ref.{kind}.")
        for k, p in mod.named_parameters():
            e = rel_err(p.grad, want[k]) if want[k].numel() > 1 else 0.0
            if not (e <= worst[1]):
                worst = (kind + "." + k, e)
    print(it, "d_loss %.6f gp %.6f g_loss %.6f" % (float(out["d_loss"]), float(out["gp"]), float(out["g_loss"])), worst)
