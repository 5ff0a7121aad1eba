#!/usr/bin/env python
"""One eager train step of a BASELINE config inside a cudaProfilerStart/Stop range (for
`ncu --profile-from-start off` launch lists).  python tools/one_step.py [--config cfg3]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg  # noqa: E402
from saragan_b200 import costmodel as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
args = ap.parse_args()
cfg = C.CONFIGS[args.config]
vol = C.volume(cfg["phase"])
torch.manual_seed(0)
g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
from saragan_b200 import graph as G  # noqa: E402
g_opt, d_opt = G.make_capturable_optimizers(g, d)[:2]
x = torch.rand(cfg["batch"], 1, *vol, device="cuda") * 2
sg.train_step(x, g, d, g_opt, d_opt, 0.5)      # warm-up (lazy init, kernel attributes)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = sg.train_step(x, g, d, g_opt, d_opt, 0.5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("d_loss", float(out["d_loss"]), "g_loss", float(out["g_loss"]))
