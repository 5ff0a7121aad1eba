"""Full-size parity: one train step of the CUDA path at the SHAPE of the headline configuration (BASELINE cfg3, 'small'
final phase, 32x128x128) against the fp32 CPU oracle's stored results (tests/golden/fullsize_cfg3_b2.json, minted by
oracle/pin_fullsize.py from seeds both sides can regenerate), and the top-level block of cfg5 (64x256x256) layer-locally
against the oracle run on the box's host cores."""
import json
import os

import numpy as np
import pytest
import torch

import saragan_b200 as sg
from oracle import pgan_oracle as O
from saragan_b200 import costmodel as C
from saragan_b200.data import step_draws, synthetic_reals
from tests.util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _need(precision):
    if precision not in getattr(sg.config, "PRECISIONS", ("bf16", "fp32")):
        pytest.skip(f"precision mode {precision} not built")


def _run(precision, name="cfg3", batch=2):
    with open(os.path.join(GOLDEN, f"fullsize_{name}_b{batch}.json")) as f:
        ref = json.load(f)
    cfg = C.CONFIGS[name]
    vol = C.volume(cfg["phase"])
    with sg.use_precision(precision):
        torch.manual_seed(ref["weights_seed"])
        g = sg.Generator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
        d = sg.Discriminator(cfg["phase"], cfg["num_phases"], cfg["base_dim"], cfg["latent_dim"], C.BASE_SHAPE)
        g_opt, d_opt = sg.make_optimizers(g, d)
        dr = step_draws(batch, vol, cfg["latent_dim"], ref["draw_seed"])
        out = sg.train_step(synthetic_reals(batch, vol, ref["x_seed"]), g, d, g_opt, d_opt, ref["alpha"], apply=False, **dr)
    torch.cuda.synchronize()
    return ref, out, g, d


# loss tolerance / bounds on |norm - oracle norm| / oracle norm of every parameter gradient: median and maximum over
# the WEIGHT tensors, maximum over the BIAS gradients.  The per-element gradient error of a reduced-precision step
# against an fp32 run is dominated by LeakyReLU mask flips and the real/fake cancellation at initialisation (DESIGN.md
# "Precision": even TF32 operands everywhere give 1.4 % median), so the NORMS are what a whole-step check can pin;
# elementwise parity is checked layer-locally.  Bias gradients are signed sums over every voxel of a level (the ToRGB
# ones over the whole image gradient): |sum| << sum of |terms|, a handful of flips moves them by 10-30 %, and they move
# from run to run with the order of the fp32 atomics.  Measured on B200 over three runs: bf16 policy weights median
# 0.9-1.7e-2, worst weight 4-12e-2, worst bias 0.16-0.32; tf32 mode 0.5-1.4e-2 / 2-6e-2 / 0.02-0.15.
@pytest.mark.parametrize("precision,ltol,med,mx,mxb", [("bf16", 2e-3, 3e-2, 0.2, 0.6), ("tf32", 1e-3, 2e-2, 0.12, 0.4)])
def test_cfg3_shaped_step_against_oracle(precision, ltol, med, mx, mxb):
    _need(precision)
    ref, out, g, d = _run(precision)
    for k in ("d_loss", "gp"):
        want = ref["losses"][k]
        assert abs(float(out[k]) - want) < ltol * abs(want), (k, float(out[k]), want)
    assert abs(float(out["g_loss"]) - ref["losses"]["g_loss"]) < ltol, float(out["g_loss"])
    errs, errs_b = {}, {}
    for mod, key in ((d, "d_grad_norms"), (g, "g_grad_norms")):
        got = {k: p.grad for k, p in mod.named_parameters() if p.grad is not None}
        assert set(got) == set(ref[key]), key                      # exactly the active levels' parameters have gradients
        for k, want in ref[key].items():
            assert torch.isfinite(got[k]).all(), k
            (errs_b if k.endswith(".bias") else errs)[key[0] + "." + k] = abs(float(got[k].double().norm()) - want) / want
    e = np.array(list(errs.values()))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(f"\n[fullsize cfg3 B=2 {precision}] gradient-norm error over {len(e)} weight tensors: median {np.median(e):.3e} "
          f"max {e.max():.3e} {worst}; bias gradients: max {max(errs_b.values()):.3e}")
    assert np.median(e) < med and e.max() < mx and max(errs_b.values()) < mxb, (np.median(e), e.max(), worst, errs_b)


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("tf32", 1e-3)])
def test_cfg5_top_level_block_layer_local(precision, tol):
    """D's top block of BASELINE cfg5 (xs 64x256x256: conv 16->16, conv 16->32, avg-pool; activation-memory-bound) on
    one sample against the oracle: output activation within the tolerance."""
    _need(precision)
    torch.manual_seed(11)
    blk = sg.DiscriminatorBlock(16, 32).cuda()
    x = torch.randn(1, 16, 64, 256, 256, generator=torch.Generator().manual_seed(3))
    p = {k: v.detach().cpu() for k, v in blk.named_parameters()}
    with torch.no_grad():
        want = O.pool2(O.lrelu(O.eq_conv3d(O.lrelu(O.eq_conv3d(x, p["conv1.weight"], p["conv1.bias"], 1)),
                                           p["conv2.weight"], p["conv2.bias"], 1)))
    with sg.use_precision(precision), torch.no_grad():
        got = blk(x.cuda())
    assert rel_err(got, want) < tol, rel_err(got, want)
