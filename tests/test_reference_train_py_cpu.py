"""Drop-in at the module level, proven with the reference's OWN step code: the unmodified
``/root/reference/pgan_pytorch/train.py`` is imported with its sibling imports (``loss``, ``metrics``, ``utils``)
resolved to shims that re-export saragan_b200, and its ``train_epoch`` (train.py:126-198, as written) drives
saragan_b200's network_dict modules on the emulated kernels.  The result must equal what the same function produced
on the reference's own modules (tests/golden/dict_epoch.npz, oracle/pin_epoch_against_reference.py).

Needs the reference tree, i.e. runs in the build container only (skipped elsewhere)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

import saragan_b200 as sg
from tests.util import GOLDEN

REF = "/root/reference/pgan_pytorch"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train.py")), reason="reference tree not present")
def test_unmodified_reference_train_epoch_drives_our_modules(cpu_kernels, monkeypatch):
    from saragan_b200 import loss as our_loss
    from saragan_b200 import metrics as our_metrics
    from saragan_b200 import network_dict as nd
    # train.py:1-10: `from loss import ...`, `from utils import write_summary, count_parameters`,
    # `import matplotlib.pyplot`, `import horovod.torch`, `from metrics import ...`
    shims = {name: types.ModuleType(name) for name in ("loss", "utils", "metrics", "matplotlib", "matplotlib.pyplot",
                                                        "horovod", "horovod.torch")}
    shims["loss"].wasserstein_loss = our_loss.wasserstein_loss
    shims["loss"].compute_gradient_penalty = our_loss.compute_gradient_penalty
    shims["metrics"].kolmogorov_smirnov_distance = our_metrics.kolmogorov_smirnov_distance
    shims["metrics"].sliced_wasserstein_distance = our_metrics.sliced_wasserstein_distance
    shims["utils"].write_summary = lambda *a, **k: None
    shims["utils"].count_parameters = lambda m: sum(p.numel() for p in m.parameters() if p.requires_grad)
    shims["matplotlib"].pyplot = shims["matplotlib.pyplot"]
    shims["horovod"].torch = shims["horovod.torch"]
    for name, mod in shims.items():
        monkeypatch.setitem(sys.modules, name, mod)
    spec = importlib.util.spec_from_file_location("reference_train", os.path.join(REF, "train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)

    z = np.load(os.path.join(GOLDEN, "dict_epoch.npz"))
    cfg = {k: int(z[k]) for k in ("phase", "num_phases", "base_dim", "latent_dim", "batch", "n_batches", "seed")}
    with sg.use_precision("fp32"):
        torch.manual_seed(0)
        args = (cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], (1, 1, 4, 4), str(z["nonlinearity"]))
        g, d = nd.Generator(*args, param=float(z["param"])), nd.Discriminator(*args, param=float(z["param"]))
        g_opt = torch.optim.Adam(g.parameters(), lr=1e-3, betas=(0.0, 0.99))
        d_opt = torch.optim.Adam(d.parameters(), lr=1e-3, betas=(0.0, 0.99))
        gen = torch.Generator().manual_seed(77)
        vol = tuple(s * 2 ** (cfg["phase"] - 1) for s in (1, 4, 4))
        loader = [1.0 + 0.35 * torch.randn(cfg["batch"], 1, *vol, generator=gen) for _ in range(cfg["n_batches"])]
        torch.manual_seed(cfg["seed"])
        x_fake, x_real, d_loss, g_loss, distance, gp = ref_train.train_epoch(loader, g, d, g_opt, d_opt, float(z["alpha"]))
    for got, key in ((d_loss, "d_loss"), (g_loss, "g_loss"), (distance, "distance"), (gp, "gp")):
        assert abs(float(got) - float(z["ref." + key])) < 1e-5 * max(1.0, abs(float(z["ref." + key]))), key
    assert torch.allclose(x_real, torch.from_numpy(z["ref.x_real"]), atol=1e-6)
    assert torch.allclose(x_fake, torch.from_numpy(z["ref.x_fake"]), atol=2e-3)
    worst = max(float((v - torch.from_numpy(z[p + k])).abs().max())
                for p, m in (("after.g.", g), ("after.d.", d)) for k, v in m.state_dict().items())
    assert worst < 3.5e-3, worst
    # the whole train() (train.py:30-123: alpha from 1 to 0 over the mixing epochs, LambdaLR on G, get_metrics after
    # every stabilising epoch) on our modules, against the digests of the weights it ends with on the reference's
    monkeypatch.setattr(our_metrics, "_device", lambda: torch.device("cpu"))
    with sg.use_precision("fp32"):
        torch.manual_seed(0)
        g, d = nd.Generator(*args, param=float(z["param"])), nd.Discriminator(*args, param=float(z["param"]))
        g_opt = torch.optim.Adam(g.parameters(), lr=1e-3, betas=(0.0, 0.99))
        d_opt = torch.optim.Adam(d.parameters(), lr=1e-3, betas=(0.0, 0.99))
        sched = torch.optim.lr_scheduler.LambdaLR(g_opt, lambda epoch: .99 ** epoch)
        dl = torch.utils.data.DataLoader(torch.cat(loader), batch_size=cfg["batch"], shuffle=False)
        torch.manual_seed(cfg["seed"])
        ref_train.train(g, d, g_opt, d_opt, sched, dl, int(z["train.mixing_epochs"]), int(z["train.stabilizing_epochs"]),
                        cfg["phase"], None)
    assert abs(g_opt.param_groups[0]["lr"] - float(z["train.g_lr"])) < 1e-12
    for prefix, mod in (("train.g.", g), ("train.d.", d)):
        for k, v in mod.state_dict().items():
            want_sum, want_abs = z[prefix + k]
            # 9 Adam(beta1 = 0) steps of +-lr per element: digests agree far below one step's worth of change
            assert abs(float(v.double().abs().sum()) - want_abs) < 1e-4 * want_abs + 1e-6, k
            assert abs(float(v.double().sum()) - want_sum) < 1e-4 * want_abs + 1e-6, k

    # the reference's get_metrics (train.py:12-27) on top of our metrics module, called the way train.py:76,99,120
    # call it (numpy arrays, no generator argument): exact KS distance, the reference's labels
    from tests.test_metrics_oracle_cpu import load_metrics_golden
    zm, real, fake, _ = load_metrics_golden("metrics_w64")
    torch.manual_seed(3)
    m = ref_train.get_metrics(real, fake)
    assert set(m) == {"mean_swd", "swd_64", "kms"} and float(m["kms"]) == float(zm["ref.kms"])
    assert np.allclose([m["swd_64"], m["mean_swd"]], zm["ref.swd"], rtol=0.25)      # device-style random stream
