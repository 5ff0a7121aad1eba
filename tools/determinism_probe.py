"""Run-to-run determinism of one train step (one GPU): the same step twice on freshly built, identically seeded
networks, per precision mode; reports which gradients differ and by how much, and whether the forward passes differ."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import saragan_b200 as sg

CFG = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
VOL, B = (4, 16, 16), 4
if len(sys.argv) > 2 and sys.argv[2] == "mid":
    CFG = dict(phase=4, num_phases=5, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4)); VOL = (8, 32, 32)
dev = torch.device("cuda")

def build():
    torch.manual_seed(5)
    a = (CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
    return sg.Generator(*a).to(dev), sg.Discriminator(*a).to(dev)

def draws():
    gen = torch.Generator().manual_seed(11)
    return dict(noise=torch.randn(B, 1, *VOL, generator=gen).to(dev), z_d=torch.randn(B, 64, generator=gen).to(dev),
                z_g=torch.randn(B, 64, generator=gen).to(dev), eps=torch.rand(B, 1, 1, 1, 1, generator=gen).to(dev))

x = (torch.rand(B, 1, *VOL, generator=torch.Generator().manual_seed(3)) * 2).to(dev)
from saragan_b200 import ops
KEEP, VARIANT = [], os.environ.get("PROBE_VARIANT", "default")
if VARIANT == "single":
    ops.packs_settled = lambda net: False
for mode in sys.argv[1].split(","):
    sg.set_precision(mode)
    runs = []
    for r in range(3):
        g, d = build()
        if VARIANT == "keep":
            KEEP.append((g, d))
        print("settled at start:", ops.packs_settled(d), flush=True)
        opts = sg.make_optimizers(g, d)
        with torch.no_grad():
            fake = g(draws()["z_d"], 0.5)
            score = d(x, 0.5)
        o = sg.train_step(x, g, d, *opts, 0.5, apply=False, **draws())
        torch.cuda.synchronize()
        runs.append(dict(fake=[f.clone() for f in fake] if isinstance(fake, (list, tuple)) else [fake.clone()], score=score.clone(),
                         losses=[float(o[k]) for k in ("d_loss", "gp", "g_loss")],
                         grads={("g." if net is g else "d.") + k: p.grad.clone() for net in (g, d)
                                for k, p in net.named_parameters() if p.grad is not None}))
    for q, r in ((0, 1), (0, 2), (1, 2)):
        a, b = runs[q], runs[r]
        fwd_g = all(torch.equal(u, v) for u, v in zip(a["fake"], b["fake"]))
        fwd_d = torch.equal(a["score"], b["score"])
        worst = sorted(((float((a["grads"][k] - b["grads"][k]).abs().max() / (a["grads"][k].abs().max() + 1e-30)), k)
                        for k in a["grads"]), reverse=True)
        nz = sum(1 for e, _ in worst if e > 0)
        print(f"{VARIANT} {mode} run{q} vs run{r}: G fwd identical {fwd_g}, D fwd identical {fwd_d}, losses {a['losses']} / {b['losses']}; "
              f"{nz}/{len(worst)} gradients differ; worst: " + ", ".join(f"{k} {e:.2e}" for e, k in worst[:6]), flush=True)
