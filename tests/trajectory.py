"""100-step loss-trajectory protocol of BASELINE.json's north star ("G/D/GP loss trajectories over 100 steps within
1 %"), SURVEY.md 0.9 / 7.4: a free-running GAN trajectory is chaotic (Adam with beta1 = 0 turns a near-zero
gradient of either sign into a +-lr step), so rounding differences of ANY implementation -- including two fp32 runs
with different summation orders -- separate over 100 steps.  The protocol therefore has two parts:

* teacher-forced: before every step the CUDA arm is given the oracle's current weights and Adam state, both arms
  take the step (same batch, same random draws, optimisers applied), the three losses of that step are compared:
  100 independent one-step parity checks along the oracle's own trajectory;
* free-running: both arms run the 100 steps on their own; the deviation is normalised by the range the oracle's
  trajectory covers.
"""
import numpy as np
import torch

import saragan_b200 as sg
from oracle import pgan_oracle as O
from tests.util import build_pair, draw_inputs

CFG = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4), batch=4)
ALPHA = 0.5
KEYS = ("d_loss", "gp", "g_loss")


def _batches(n, cfg):
    """a pool of CT-like batches cycled like a data loader would"""
    vol = tuple(s * 2 ** (cfg["phase"] - 1) for s in cfg["base_shape"][1:])
    out = []
    for i in range(n):
        x = torch.randn(cfg["batch"], 1, *vol, generator=torch.Generator().manual_seed(300 + i))
        k = torch.ones(1, 1, 3, 3, 3) / 27
        for _ in range(3):
            x = torch.nn.functional.conv3d(x, k, padding=1)
        out.append(torch.clamp(1024 + 350 * x / x.std(), 0, 3072) / 1024)
    return out


def _inputs(step, pool, cfg):
    inp = draw_inputs(cfg, seed=1000 + step)
    inp["x_real"] = pool[step % len(pool)]
    return inp


def _load_optimizer(dst, src):
    """torch.optim state by parameter index (both optimisers were built over the same parameter order)"""
    sd = src.state_dict()
    dst.load_state_dict({"state": {k: {n: (t.clone() if torch.is_tensor(t) else t) for n, t in v.items()}
                                   for k, v in sd["state"].items()}, "param_groups": dst.state_dict()["param_groups"]})


def run(precision, steps=100, cfg=CFG, alpha=ALPHA):
    """returns (oracle, teacher_forced, free_running) arrays of shape (steps, 3) in KEYS order"""
    pool = _batches(4, cfg)
    with sg.use_precision(precision):
        g, d = build_pair(cfg, seed=3)
        gf, df = build_pair(cfg, seed=3)                      # the free-running arm
        st = O.TrainState({k: v.detach().cpu() for k, v in g.state_dict().items()},
                          {k: v.detach().cpu() for k, v in d.state_dict().items()}, cfg["phase"], cfg["num_phases"])
        g_opt, d_opt = sg.make_optimizers(g, d)
        free_opts = sg.make_optimizers(gf, df)
        oracle, forced, free = [], [], []
        for step in range(steps):
            inp = _inputs(step, pool, cfg)
            # teacher forcing: the oracle's state BEFORE this step
            with torch.no_grad():
                for mod, ref in ((g, st.pg), (d, st.pd)):
                    for name, p in mod.named_parameters():
                        p.copy_(ref[name])
                        torch.autograd.graph.increment_version(p)
            if step > 0:
                _load_optimizer(g_opt, st.g_opt)
                _load_optimizer(d_opt, st.d_opt)
            want = st.step(inp["x_real"], inp["noise"], inp["z_d"], inp["eps"], inp["z_g"], alpha)
            got = sg.train_step(inp["x_real"], g, d, g_opt, d_opt, alpha, noise=inp["noise"], z_d=inp["z_d"],
                                z_g=inp["z_g"], eps=inp["eps"])
            fr = sg.train_step(inp["x_real"], gf, df, *free_opts, alpha, noise=inp["noise"], z_d=inp["z_d"],
                               z_g=inp["z_g"], eps=inp["eps"])
            oracle.append([want[k] for k in KEYS])
            forced.append([float(got[k]) for k in KEYS])
            free.append([float(fr[k]) for k in KEYS])
    return np.array(oracle), np.array(forced), np.array(free)


def summarize(oracle, forced, free):
    """per loss: worst / median teacher-forced deviation relative to max(|loss|, 5 % of the largest |loss| of the
    trajectory) -- g_loss crosses zero -- and the free-running deviation normalised by the oracle trajectory's range"""
    rng = oracle.max(0) - oracle.min(0)
    scale = np.maximum(np.abs(oracle), 0.05 * np.abs(oracle).max(0))     # g_loss crosses zero: floor the denominator
    tf = (np.abs(forced - oracle) / scale).max(0)
    tf_med = np.median(np.abs(forced - oracle) / scale, axis=0)
    fr = np.abs(free - oracle).max(0) / rng
    fr_final = np.abs(free[-10:] - oracle[-10:]).mean(0) / rng
    return {k: dict(teacher_forced_rel=float(tf[i]), teacher_forced_median_rel=float(tf_med[i]), free_running_range_norm=float(fr[i]),
                    free_running_last10_range_norm=float(fr_final[i]), oracle_first=float(oracle[0, i]),
                    oracle_last=float(oracle[-1, i]), oracle_range=float(rng[i])) for i, k in enumerate(KEYS)}
