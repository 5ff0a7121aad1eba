"""``saragan_b200.train_epoch`` against the UNMODIFIED reference ``train.py::train_epoch`` (train.py:126-198) run as
written on network_dict.py modules (fixture: oracle/pin_epoch_against_reference.py): same seed, same batches, three
optimiser steps -- same signature, same return tuple, same values.  Runs on the emulated kernels: every random
number of the epoch then comes from torch's CPU generator in train.py's order (noise, z, eps, z), which this
implementation keeps."""
import os

import numpy as np
import torch

import saragan_b200 as sg
from tests.util import GOLDEN


def test_train_epoch_matches_reference_as_written(cpu_kernels):
    from saragan_b200 import network_dict as nd
    z = np.load(os.path.join(GOLDEN, "dict_epoch.npz"))
    cfg = {k: int(z[k]) for k in ("phase", "num_phases", "base_dim", "latent_dim", "batch", "n_batches", "seed")}
    alpha = float(z["alpha"])
    with sg.use_precision("fp32"):
        torch.manual_seed(0)
        args = (cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], (1, 1, 4, 4), str(z["nonlinearity"]))
        g, d = nd.Generator(*args, param=float(z["param"])), nd.Discriminator(*args, param=float(z["param"]))
        for prefix, mod in (("g.", g), ("d.", d)):
            for k, v in mod.state_dict().items():
                assert torch.equal(v, torch.from_numpy(z[prefix + k])), k          # the reference's initial weights
        g_opt, d_opt = sg.make_optimizers(g, d)
        gen = torch.Generator().manual_seed(77)
        vol = tuple(s * 2 ** (cfg["phase"] - 1) for s in (1, 4, 4))
        loader = [1.0 + 0.35 * torch.randn(cfg["batch"], 1, *vol, generator=gen) for _ in range(cfg["n_batches"])]
        torch.manual_seed(cfg["seed"])
        out = sg.train_epoch(loader, g, d, g_opt, d_opt, alpha)
    assert isinstance(out, tuple) and len(out) == 6
    x_fake, x_real, d_loss, g_loss, distance, gp = out
    assert x_fake.device.type == "cpu" and x_real.device.type == "cpu" and not x_fake.requires_grad
    for got, key in ((d_loss, "d_loss"), (g_loss, "g_loss"), (distance, "distance"), (gp, "gp")):
        assert isinstance(got, np.floating)
        assert abs(float(got) - float(z["ref." + key])) < 1e-5 * max(1.0, abs(float(z["ref." + key]))), key      # measured 3e-7
    assert torch.allclose(x_real, torch.from_numpy(z["ref.x_real"]), atol=1e-6)     # the noisy reals of the last batch
    assert torch.allclose(x_fake, torch.from_numpy(z["ref.x_fake"]), atol=2e-3)                 # measured 1e-6
    # weights after three Adam(beta1 = 0) steps: a +-lr step per update, so compare at the scale of 3 * lr
    worst = 0.0
    for prefix, mod in (("after.g.", g), ("after.d.", d)):
        for k, v in mod.state_dict().items():
            worst = max(worst, float((v - torch.from_numpy(z[prefix + k])).abs().max()))
    assert worst < 3.5e-3, worst                                                    # measured 4.5e-6
    assert all(p.requires_grad for p in list(g.parameters()) + list(d.parameters()))       # train.py:192-196
