"""Pin oracle/pgan_oracle.py against the UNMODIFIED reference and mint tests/golden/*.npz.

Runs only in the build container (needs /root/reference).  Nothing in tests/, smoke() or
bench.py imports this file; they read the committed .npz fixtures instead.

    python oracle/pin_against_reference.py            # check + write fixtures
    python oracle/pin_against_reference.py --check    # check only

In-process shims (no edits to the reference): numpy.product = numpy.prod,
stdout of the debug prints swallowed.  train.py is not imported (needs horovod and
matplotlib); its step body train.py:133-190 is what oracle.TrainState restates, and this
script drives the reference *modules* through the same sequence for the comparison.
"""
import argparse
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import pgan_oracle as O  # noqa: E402

REF = "/root/reference/pgan_pytorch"


def load_reference():
    np.product = np.prod  # NumPy 2 shim (network.py:161,247)
    mods = {}
    for name in ("network", "loss"):
        spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["network"], mods["loss"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def build(net, cfg, seed):
    torch.manual_seed(seed)
    g = quiet(net.Generator, cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"],
              cfg["base_shape"])
    d = quiet(net.Discriminator, cfg["phase"], cfg["num_phases"], cfg["base_dim"],
              cfg["latent_dim"], cfg["base_shape"])
    return g, d


def draw_inputs(cfg, seed):
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


def reference_step(net, loss_mod, g, d, inp, alpha):
    """The reference modules driven through train.py:133-190 (no optimiser step)."""
    for p_ in g.parameters():
        p_.requires_grad = False
    for p_ in d.parameters():
        p_.requires_grad = True
    x_real = inp["x_real"] + inp["noise"] * 1e-2
    x_fake = quiet(g, inp["z_d"], alpha)[-1].detach()
    d_real = quiet(d, x_real, alpha)
    d_fake = quiet(d, x_fake, alpha)
    # loss.py:11 draws eps with torch.rand; feed ours by patching torch.rand for one call
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: inp["eps"]
    try:
        gp = quiet(loss_mod.compute_gradient_penalty, d, x_real, x_fake, alpha)
    finally:
        torch.rand = orig_rand
    d_loss = -loss_mod.wasserstein_loss(d_real) + loss_mod.wasserstein_loss(d_fake) + gp \
        + 1e-3 * (d_real ** 2).mean()
    d.zero_grad()
    d_loss.backward()
    d_grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in d.named_parameters()}
    for p_ in g.parameters():
        p_.requires_grad = True
    for p_ in d.parameters():
        p_.requires_grad = False
    imgs = quiet(g, inp["z_g"], alpha)
    d_fake2 = quiet(d, imgs[-1], alpha)
    g_loss = -loss_mod.wasserstein_loss(d_fake2)
    g.zero_grad()
    g_loss.backward()
    g_grads = {k: (None if v.grad is None else v.grad.clone()) for k, v in g.named_parameters()}
    return dict(d_loss=d_loss.detach(), gp=gp.detach(), g_loss=g_loss.detach(),
                d_real=d_real.detach(), d_fake=d_fake.detach(), imgs=[i.detach() for i in imgs],
                d_grads=d_grads, g_grads=g_grads)


def maxdiff(a, b):
    return float((a - b).abs().max())


CONFIGS = {
    # tiny: full tensors are stored (weights, inputs, every output and gradient)
    "tiny_p3": dict(phase=3, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4),
                    batch=4, alpha=0.5, full=True),
    "tiny_p2_b8": dict(phase=2, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4),
                       batch=8, alpha=0.25, full=True),
    "tiny_p1": dict(phase=1, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4),
                    batch=4, alpha=0.0, full=True),
    # edge cases of the fade-in and of the minibatch-stddev group rule (network.py:119-124): alpha at both ends of
    # its schedule (train.py:33,82), batch 6 -> one group of 6, batch 3 -> one group of 3
    "tiny_p3_b6_a1": dict(phase=3, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4),
                          batch=6, alpha=1.0, full=True),
    "tiny_p2_b3_a0": dict(phase=2, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4),
                          batch=3, alpha=0.0, full=True),
    # BASELINE cfg1 (xs, 4x16x16, B=4): weights regenerated from the seed by the oracle's own
    # modules is not bit-identical to the reference's stream, so only scalars + grad norms of
    # the reference run are stored together with the reference's initial state digest.
    "cfg1_xs_p3": dict(phase=3, num_phases=6, base_dim=256, latent_dim=256,
                       base_shape=(1, 1, 4, 4), batch=4, alpha=0.5, full=False),
}


def run_config(name, cfg, net, loss_mod, write):
    g, d = build(net, cfg, seed=0)
    inp = draw_inputs(cfg, seed=123)
    alpha = cfg["alpha"]
    ref = reference_step(net, loss_mod, g, d, inp, alpha)

    pg = {k: v.detach().clone() for k, v in g.state_dict().items()}
    pd = {k: v.detach().clone() for k, v in d.state_dict().items()}
    st = O.TrainState(pg, pd, cfg["phase"], cfg["num_phases"])
    got = st.step(inp["x_real"], inp["noise"], inp["z_d"], inp["eps"], inp["z_g"], alpha, apply=False)

    worst = 0.0
    for k in ("d_loss", "gp", "g_loss"):
        dlt = abs(float(ref[k]) - got[k])
        worst = max(worst, dlt)
        print(f"  {name}: {k}: ref {float(ref[k]):+.8f} oracle {got[k]:+.8f} |d|={dlt:.2e}")
    imgs = O.generator_forward(st.pg, inp["z_g"], alpha, cfg["phase"])
    for a, b in zip(ref["imgs"], imgs):
        worst = max(worst, maxdiff(a, b.detach()))
    n_none = 0
    for kind in ("d_grads", "g_grads"):
        for k, v in ref[kind].items():
            o = got[kind][k]
            if v is None:
                assert o is None, f"{k}: reference grad None, oracle not"
                n_none += 1
                continue
            assert o is not None, f"{k}: oracle grad None, reference not"
            worst = max(worst, maxdiff(v, o) / (float(v.abs().max()) + 1e-30))
    act_d = set(O.active_names("d", cfg["phase"], cfg["num_phases"]))
    act_g = set(O.active_names("g", cfg["phase"], cfg["num_phases"]))
    assert act_d == {k for k, v in ref["d_grads"].items() if v is not None}, "active D set"
    assert act_g == {k for k, v in ref["g_grads"].items() if v is not None}, "active G set"
    print(f"  {name}: worst rel/abs deviation oracle vs reference = {worst:.3e} "
          f"({n_none} inactive params)")
    assert worst < 1e-6, "oracle does not reproduce the reference"

    # init_params must reproduce names and shapes
    for kind, sd in (("g", pg), ("d", pd)):
        mine = O.init_params(kind, cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"],
                             cfg["base_shape"])
        assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in sd.items()}, kind

    if not write:
        return
    out = {"alpha": np.float64(alpha)}
    for k in ("phase", "num_phases", "base_dim", "latent_dim", "batch"):
        out[k] = np.int64(cfg[k])
    for k in ("d_loss", "gp", "g_loss"):
        out["ref." + k] = np.float64(float(ref[k]))
    for k, v in inp.items():
        out["in." + k] = v.numpy()
    out["ref.d_real"] = ref["d_real"].numpy()
    out["ref.d_fake"] = ref["d_fake"].numpy()
    if cfg["full"]:
        for k, v in pg.items():
            out["g." + k] = v.numpy()
        for k, v in pd.items():
            out["d." + k] = v.numpy()
        for i, im in enumerate(ref["imgs"]):
            out[f"ref.img{i}"] = im.numpy()
        for kind in ("d_grads", "g_grads"):
            for k, v in ref[kind].items():
                if v is not None:
                    out[f"ref.{kind}.{k}"] = v.numpy()
        out["ref.gp_grad"] = got["gp_grad"].numpy()
    else:
        for kind in ("d_grads", "g_grads"):
            for k, v in ref[kind].items():
                if v is not None:
                    out[f"ref.{kind}.norm.{k}"] = np.float64(float(v.double().norm()))
        out["ref.img_last.sum"] = np.float64(float(ref["imgs"][-1].double().sum()))
        out["ref.img_last.abs"] = np.float64(float(ref["imgs"][-1].double().abs().sum()))
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(8)
    net, loss_mod = load_reference()
    # num_filters / mbstd group rule against the reference's own functions
    for base in (32, 64, 256, 512, 1024):
        for nph in range(int(np.log2(base / 16)) + 2, 9):
            for ph in range(1, nph + 2):
                assert net.num_filters(ph, nph, base) == O.num_filters(ph, nph, base)
    print("num_filters: ok")
    for name, cfg in CONFIGS.items():
        run_config(name, cfg, net, loss_mod, write=not args.check)
    print("oracle pinned against reference: OK")


if __name__ == "__main__":
    main()
