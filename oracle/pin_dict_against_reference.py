"""Pin oracle/pgan_dict_oracle.py against the UNMODIFIED reference pgan_pytorch/network_dict.py + loss.py and mint
tests/golden/dict_*.npz.  Build container only (needs /root/reference); tests read the committed fixtures.

    python oracle/pin_dict_against_reference.py [--check]

Shims (no edits to the reference): numpy.product = numpy.prod (network_dict.py:216,339).  The reference modules are
driven through train.py:133-190 (train.py itself needs horovod + matplotlib and is not imported).
"""
import argparse
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import pgan_dict_oracle as O  # noqa: E402

REF = "/root/reference/pgan_pytorch"


def load_reference():
    np.product = np.prod
    mods = {}
    for name in ("network_dict", "loss"):
        spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["network_dict"], mods["loss"]


CONFIGS = {
    "dict_p3_lrelu": dict(phase=3, grow_from=3, num_phases=3, base_dim=32, latent_dim=32, batch=4, alpha=0.5,
                          nonlinearity="leaky_relu", param=0.3),
    "dict_p2_relu": dict(phase=2, grow_from=2, num_phases=3, base_dim=32, latent_dim=32, batch=2, alpha=0.25,
                         nonlinearity="relu", param=None),
    "dict_p1_lrelu": dict(phase=1, grow_from=1, num_phases=3, base_dim=32, latent_dim=32, batch=4, alpha=0.0,
                          nonlinearity="leaky_relu", param=0.3),
    # built at phase 2, then grow() once (main.py:54-58): names, RNG stream and arithmetic after growing
    "dict_grow_p2to3": dict(phase=3, grow_from=2, num_phases=4, base_dim=32, latent_dim=32, batch=3, alpha=0.75,
                            nonlinearity="leaky_relu", param=0.2),
}
BASE_SHAPE = (1, 1, 4, 4)


def build(net, cfg, seed=0):
    torch.manual_seed(seed)
    args = (cfg["grow_from"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], BASE_SHAPE, cfg["nonlinearity"])
    g = net.Generator(*args, param=cfg["param"])
    d = net.Discriminator(*args, param=cfg["param"])
    for _ in range(cfg["phase"] - cfg["grow_from"]):
        g.grow()
        d.grow()
    return g, d


def draw_inputs(cfg, seed=123):
    gen = torch.Generator().manual_seed(seed)
    b, p = cfg["batch"], cfg["phase"]
    vol = tuple(s * 2 ** (p - 1) for s in BASE_SHAPE[1:])
    return dict(x_real=torch.randn(b, 1, *vol, generator=gen), noise=torch.randn(b, 1, *vol, generator=gen),
                z_d=torch.randn(b, cfg["latent_dim"], generator=gen), z_g=torch.randn(b, cfg["latent_dim"], generator=gen),
                eps=torch.rand(b, 1, 1, 1, 1, generator=gen))


def reference_step(loss_mod, g, d, inp, alpha):
    for p_ in g.parameters():
        p_.requires_grad = False
    for p_ in d.parameters():
        p_.requires_grad = True
    x_real = inp["x_real"] + inp["noise"] * 1e-2
    x_fake = g(inp["z_d"], alpha).detach()                      # train.py:146 as written
    d_real, d_fake = d(x_real, alpha), d(x_fake, alpha)
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: inp["eps"]                     # loss.py:11
    try:
        gp = loss_mod.compute_gradient_penalty(d, x_real, x_fake, alpha)
    finally:
        torch.rand = orig_rand
    d_loss = -loss_mod.wasserstein_loss(d_real) + loss_mod.wasserstein_loss(d_fake) + gp + 1e-3 * (d_real ** 2).mean()
    d.zero_grad()
    d_loss.backward()
    d_grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in d.named_parameters()}
    for p_ in g.parameters():
        p_.requires_grad = True
    for p_ in d.parameters():
        p_.requires_grad = False
    img = g(inp["z_g"], alpha)
    g_loss = -loss_mod.wasserstein_loss(d(img, alpha))
    g.zero_grad()
    g_loss.backward()
    g_grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in g.named_parameters()}
    return dict(d_loss=d_loss.detach(), gp=gp.detach(), g_loss=g_loss.detach(), d_real=d_real.detach(),
                d_fake=d_fake.detach(), img=img.detach(), d_grads=d_grads, g_grads=g_grads)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(8)
    net, loss_mod = load_reference()
    for name, cfg in CONFIGS.items():
        g, d = build(net, cfg)
        inp = draw_inputs(cfg)
        ref = reference_step(loss_mod, g, d, inp, cfg["alpha"])
        pg = {k: v.detach().clone() for k, v in g.state_dict().items()}
        pd = {k: v.detach().clone() for k, v in d.state_dict().items()}
        st = O.DictTrainState(pg, pd, cfg["phase"], cfg["nonlinearity"], cfg["param"])
        got = st.step(inp["x_real"], inp["noise"], inp["z_d"], inp["eps"], inp["z_g"], cfg["alpha"], apply=False)
        worst = max(abs(float(ref[k]) - got[k]) for k in ("d_loss", "gp", "g_loss"))
        worst = max(worst, float((ref["img"] - got["img"]).abs().max()))
        for kind in ("d_grads", "g_grads"):
            for k, v in ref[kind].items():
                o = got[kind][k]
                assert (v is None) == (o is None), k
                if v is not None:
                    worst = max(worst, float((v - o).abs().max()) / (float(v.abs().max()) + 1e-30))
        print(f"{name}: d_loss {float(ref['d_loss']):+.6f} gp {float(ref['gp']):.6f} g_loss {float(ref['g_loss']):+.6f}; "
              f"worst deviation oracle vs reference {worst:.3e}")
        assert worst == 0.0, "oracle does not reproduce the reference bit for bit"
        if args.check:
            continue
        out = {"alpha": np.float64(cfg["alpha"]), "nonlinearity": np.array(cfg["nonlinearity"]),
               "param": np.float64(-1.0 if cfg["param"] is None else cfg["param"])}
        for k in ("phase", "grow_from", "num_phases", "base_dim", "latent_dim", "batch"):
            out[k] = np.int64(cfg[k])
        for k in ("d_loss", "gp", "g_loss"):
            out["ref." + k] = np.float64(float(ref[k]))
        for k, v in inp.items():
            out["in." + k] = v.numpy()
        for k in ("d_real", "d_fake", "img"):
            out["ref." + k] = ref[k].numpy()
        for k, v in pg.items():
            out["g." + k] = v.numpy()
        for k, v in pd.items():
            out["d." + k] = v.numpy()
        for kind in ("d_grads", "g_grads"):
            for k, v in ref[kind].items():
                if v is not None:
                    out[f"ref.{kind}.{k}"] = v.numpy()
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"  wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    print("network_dict oracle pinned against reference: OK")


if __name__ == "__main__":
    main()
