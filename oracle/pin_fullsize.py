"""TEST INFRASTRUCTURE.  Full-size parity fixtures: one train step (train.py:133-190, no optimiser update) of the CPU
oracle at a BASELINE configuration's real shape, on the weights the drop-in modules draw from seed 0 (the reference's
own RNG stream, tests/test_wiring_cpu.py) and on inputs every implementation can regenerate from seeds
(saragan_b200.data.synthetic_reals / step_draws).  Stored: the three losses and the norm of every parameter gradient.

    python oracle/pin_fullsize.py cfg3 2      # -> tests/golden/fullsize_cfg3_b2.json   (tests/test_fullsize_gpu.py)
    python oracle/pin_fullsize.py cfg3 4      # -> tests/golden/fullsize_cfg3_b4.json   (bench.py's parity check)

The oracle itself is pinned bit-for-bit against the unmodified reference by oracle/pin_against_reference.py."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

X_SEED, DRAW_SEED, ALPHA = 4321, 8765, 0.5


def main(name, batch):
    import saragan_b200 as sg
    from oracle import pgan_oracle as O
    from saragan_b200 import costmodel as C
    from saragan_b200.data import step_draws, synthetic_reals
    cfg = C.CONFIGS[name]
    vol = C.volume(cfg["phase"])
    torch.manual_seed(0)
    g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
    pg = {k: v.detach().cpu() for k, v in g.state_dict().items()}
    pd = {k: v.detach().cpu() for k, v in d.state_dict().items()}
    x = synthetic_reals(batch, vol, X_SEED)
    dr = step_draws(batch, vol, cfg["latent_dim"], DRAW_SEED)
    t0 = time.time()
    want = O.TrainState(pg, pd, cfg["phase"], cfg["num_phases"]).step(x, dr["noise"], dr["z_d"], dr["eps"], dr["z_g"],
                                                                     ALPHA, apply=False)
    out = {"config": name, "batch": batch, "alpha": ALPHA, "x_seed": X_SEED, "draw_seed": DRAW_SEED,
           "weights_seed": 0, "torch": torch.__version__, "oracle_seconds": round(time.time() - t0, 1),
           "losses": {k: float(want[k]) for k in ("d_loss", "gp", "g_loss")},
           "d_grad_norms": {k: float(v.double().norm()) for k, v in want["d_grads"].items() if v is not None},
           "g_grad_norms": {k: float(v.double().norm()) for k, v in want["g_grads"].items() if v is not None}}
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_{name}_b{batch}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, out["losses"], f"{out['oracle_seconds']} s")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
