"""Helpers for the network_dict.py parity tests (fixtures: oracle/pin_dict_against_reference.py)."""
import os

import numpy as np
import torch

from tests.util import GOLDEN

DICT_CASES = ["dict_p3_lrelu", "dict_p2_relu", "dict_p1_lrelu", "dict_grow_p2to3"]
BASE_SHAPE = (1, 1, 4, 4)


def load_dict_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = {k: int(z[k]) for k in ("phase", "grow_from", "num_phases", "base_dim", "latent_dim", "batch")}
    cfg["alpha"] = float(z["alpha"])
    cfg["nonlinearity"] = str(z["nonlinearity"])
    cfg["param"] = None if float(z["param"]) < 0 else float(z["param"])
    return z, cfg


def build_dict_pair(cfg, seed=0):
    """torch.manual_seed(0); Generator(...); Discriminator(...); grow() -- the order the fixtures were minted in."""
    from saragan_b200 import network_dict as nd
    torch.manual_seed(seed)
    args = (cfg["grow_from"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], BASE_SHAPE, cfg["nonlinearity"])
    g = nd.Generator(*args, param=cfg["param"])
    d = nd.Discriminator(*args, param=cfg["param"])
    for _ in range(cfg["phase"] - cfg["grow_from"]):
        g.grow()
        d.grow()
    return g, d


def golden_inputs(z):
    return {k: torch.from_numpy(z["in." + k]) for k in ("x_real", "noise", "z_d", "z_g", "eps")}
