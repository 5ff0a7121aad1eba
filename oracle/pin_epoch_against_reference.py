"""Run the UNMODIFIED reference ``train.py::train_epoch`` (train.py:126-198) on CPU and mint
tests/golden/dict_epoch.npz: the epoch's return values and the weights after it, for a fixed seed.

train.py is imported as written; only its two unavailable third-party imports (matplotlib, horovod -- neither is used
by train_epoch) are stubbed in sys.modules, and numpy.product = numpy.prod (NumPy 2).  The networks are the reference's
network_dict.py (the variant main.py imports, whose generator returns a tensor so that train.py:146 runs as written).
Every random number of the epoch comes from torch's global CPU generator in the order train.py draws them
(instance noise, z, the gradient penalty's eps, z), so ``torch.manual_seed`` fixes the whole epoch and the CUDA-path
``saragan_b200.train_epoch`` -- which draws in the same order -- can be compared from the same seed
(tests/test_epoch_cpu.py, on the emulated kernels).  Build container only.

    python oracle/pin_epoch_against_reference.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pgan_pytorch"

CFG = dict(phase=2, num_phases=3, base_dim=32, latent_dim=32, batch=4, n_batches=3, alpha=0.5, seed=11,
           nonlinearity="leaky_relu", param=0.3)
BASE_SHAPE = (1, 1, 4, 4)
TRAIN = dict(mixing_epochs=2, stabilizing_epochs=1)


def batches(cfg):
    gen = torch.Generator().manual_seed(77)
    vol = tuple(s * 2 ** (cfg["phase"] - 1) for s in BASE_SHAPE[1:])
    return [1.0 + 0.35 * torch.randn(cfg["batch"], 1, *vol, generator=gen) for _ in range(cfg["n_batches"])]


def main():
    np.product = np.prod
    for name in ("matplotlib", "matplotlib.pyplot", "horovod", "horovod.torch"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["horovod"].torch = sys.modules["horovod.torch"]
    sys.path.insert(0, REF)
    import network_dict as net      # noqa: E402  (the reference's)
    import train as ref_train       # noqa: E402  (the reference's, unmodified)

    torch.set_num_threads(8)
    cfg = CFG
    torch.manual_seed(0)
    args = (cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], BASE_SHAPE, cfg["nonlinearity"])
    g, d = net.Generator(*args, param=cfg["param"]), net.Discriminator(*args, param=cfg["param"])
    init = {"g." + k: v.clone().numpy() for k, v in g.state_dict().items()}
    init.update({"d." + k: v.clone().numpy() for k, v in d.state_dict().items()})
    g_opt = torch.optim.Adam(g.parameters(), lr=1e-3, betas=(0.0, 0.99))      # main.py:141-142 ((0, .99) there)
    d_opt = torch.optim.Adam(d.parameters(), lr=1e-3, betas=(0.0, 0.99))
    torch.manual_seed(cfg["seed"])
    x_fake, x_real, d_loss, g_loss, dist, gp = ref_train.train_epoch(batches(cfg), g, d, g_opt, d_opt, cfg["alpha"])
    print(f"reference train_epoch: d_loss {d_loss:.6f} g_loss {g_loss:.6f} distance {dist:.6f} gp {gp:.6f}")
    out = {k: np.int64(v) for k, v in cfg.items() if isinstance(v, int)}
    out.update(alpha=np.float64(cfg["alpha"]), param=np.float64(cfg["param"]), nonlinearity=np.array(cfg["nonlinearity"]))
    out.update(init)
    out.update({"ref.x_fake": x_fake.numpy(), "ref.x_real": x_real.numpy(), "ref.d_loss": np.float64(d_loss),
                "ref.g_loss": np.float64(g_loss), "ref.distance": np.float64(dist), "ref.gp": np.float64(gp)})
    out.update({"after.g." + k: v.numpy() for k, v in g.state_dict().items()})
    out.update({"after.d." + k: v.numpy() for k, v in d.state_dict().items()})
    # the whole train() of train.py:30-123 (alpha schedule 1 -> 0 over the mixing epochs, LambdaLR on G, get_metrics
    # after every stabilising epoch; writer=None, horovod=False): per-parameter digests of the weights it ends with
    torch.manual_seed(0)
    g, d = net.Generator(*args, param=cfg["param"]), net.Discriminator(*args, param=cfg["param"])
    g_opt = torch.optim.Adam(g.parameters(), lr=1e-3, betas=(0.0, 0.99))
    d_opt = torch.optim.Adam(d.parameters(), lr=1e-3, betas=(0.0, 0.99))
    sched = torch.optim.lr_scheduler.LambdaLR(g_opt, lambda epoch: .99 ** epoch)          # main.py:144-145
    loader = torch.utils.data.DataLoader(torch.cat(batches(cfg)), batch_size=cfg["batch"], shuffle=False)
    torch.manual_seed(cfg["seed"])
    ref_train.train(g, d, g_opt, d_opt, sched, loader, TRAIN["mixing_epochs"], TRAIN["stabilizing_epochs"], cfg["phase"], None)
    for prefix, mod in (("train.g.", g), ("train.d.", d)):
        for k, v in mod.state_dict().items():
            out[prefix + k] = np.array([float(v.double().sum()), float(v.double().abs().sum())])
    out["train.g_lr"] = np.float64(g_opt.param_groups[0]["lr"])
    out.update({"train." + k: np.int64(v) for k, v in TRAIN.items()})
    print(f"reference train(): {TRAIN}, final G lr {g_opt.param_groups[0]['lr']:.6g}")
    path = os.path.join(ROOT, "tests", "golden", "dict_epoch.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
