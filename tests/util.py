"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = {k: int(z[k]) for k in ("phase", "num_phases", "base_dim", "latent_dim", "batch")}
    cfg["alpha"] = float(z["alpha"])
    cfg["base_shape"] = (1, 1, 4, 4)
    return z, cfg


def golden_tensors(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / ||b||  (the tolerance metric of BASELINE.json)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_pair(cfg, device=None, seed=0):
    import saragan_b200 as sg
    torch.manual_seed(seed)
    g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], cfg["base_shape"])
    d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], cfg["base_shape"])
    return g, d


def draw_inputs(cfg, seed=123):
    """Same draw as oracle/pin_against_reference.py::draw_inputs."""
    gen = torch.Generator().manual_seed(seed)
    b, p = cfg["batch"], cfg["phase"]
    bs = cfg["base_shape"]
    vol = (bs[1] * 2 ** (p - 1), bs[2] * 2 ** (p - 1), bs[3] * 2 ** (p - 1))
    return dict(
        x_real=torch.randn(b, 1, *vol, generator=gen),
        noise=torch.randn(b, 1, *vol, generator=gen),
        z_d=torch.randn(b, cfg["latent_dim"], generator=gen),
        z_g=torch.randn(b, cfg["latent_dim"], generator=gen),
        eps=torch.rand(b, 1, 1, 1, 1, generator=gen),
    )


def run_step(g, d, inp, alpha, apply=False, opts=None):
    import saragan_b200 as sg
    g_opt, d_opt = opts if opts is not None else sg.make_optimizers(g, d)
    out = sg.train_step(inp["x_real"], g, d, g_opt, d_opt, alpha, noise=inp["noise"], z_d=inp["z_d"],
                        z_g=inp["z_g"], eps=inp["eps"], apply=apply)
    return out
