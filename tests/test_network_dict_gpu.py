"""SURVEY 8f row 3 on the GPU: the ``network_dict.py`` drop-in (LeakyReLU 0.3 / ReLU, He gain in the equalized
learning rate, no minibatch-stddev, top-level fade-in, grow()) through the C ABI against the fixtures minted from
the unmodified reference, and every kernel that carries the LeakyReLU slope against its CPU restatement at the
slopes network_dict.py uses."""
import numpy as np
import pytest
import torch

import saragan_b200 as sg
from saragan_b200 import _lib
from saragan_b200 import kernels as K
from tests import cpu_emul as E
from tests.dict_util import DICT_CASES, build_dict_pair, golden_inputs, load_dict_golden
from tests.test_kernels_gpu import act, close, rnd
from tests.util import golden_tensors, rel_err, run_step

pytestmark = pytest.mark.gpu


def _set_slope(slope):
    K.set_leaky_slope(slope)
    E.set_leaky_slope(slope)
    assert abs(K.get_leaky_slope() - slope) < 1e-7


@pytest.mark.parametrize("slope", [0.3, 0.0])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_kernels_follow_the_slope(slope, dtype):
    _set_slope(slope)
    n, cin, cout, d, h, w = 2, 32, 32, 4, 16, 16
    wt, bias = rnd(cout, cin, 3, 3, 3, seed=1), rnd(cout, seed=3)
    xa, xg = act(n, cin, d, h, w, dtype, seed=2)
    ma, mg = act(n, cout, d, h, w, dtype, seed=4)
    impls = [_lib.IMPL_DIRECT] + ([_lib.IMPL_TCGEN05] if dtype == torch.bfloat16 else [])
    for impl in impls:           # epilogue LeakyReLU and the dgrad-epilogue mask, CUDA-core and tcgen05 kernels
        for lrelu, mask in ((True, False), (False, True)):
            want = E.conv3d_fprop(xa, E.pack_conv_weight(wt, dtype, False), bias, ma if mask else None, cin, cout, 0.11, lrelu)
            got = K.conv3d_fprop(xg, K.pack_conv_weight(wt.cuda(), dtype, False), bias.cuda(), mg if mask else None,
                                 cin, cout, 0.11, lrelu, impl)
            close(got, want, dtype, f"conv impl={impl} lrelu={lrelu} mask={mask} slope={slope}")
    close(K.lrelu_fwd(xg), E.lrelu_fwd(xa), dtype, "lrelu_fwd")
    close(K.mask_mul(xg, mg), E.mask_mul(xa, ma), dtype, "mask_mul")
    close(K.pixelnorm_fwd(xg, cin, True), E.pixelnorm_fwd(xa, cin, True), dtype, "pixelnorm_fwd")
    close(K.pixelnorm_bwd(xg, mg, cin, True, True), E.pixelnorm_bwd(xa, ma, cin, True, True), dtype, "pixelnorm_bwd")
    la, lg = act(n, cin, d // 2, h // 2, w // 2, dtype, seed=5)
    close(K.up2(lg, 0.125, dtype, xg), E.up2(la, 0.125, dtype, xa), dtype, "up2 mask")
    img = rnd(n, 1, d, h, w, seed=6)
    close(K.pw_expand(img.cuda(), rnd(cin, seed=7).cuda(), bias.cuda(), dtype, cin, 0.7, True),
          E.pw_expand(img, rnd(cin, seed=7), bias, dtype, cin, 0.7, True), dtype, "pw_expand")
    x, wl = rnd(4, 96, seed=8), rnd(40, 96, seed=9)
    close(K.linear_fwd(x.cuda(), wl.cuda(), None, 0.1, True), E.linear_fwd(x, wl, None, 0.1, True), torch.float32, "linear")
    # the values really depend on the slope (guards against a constant that never reaches the device)
    neg = -torch.ones(1, 2, 1, 1, 8, 8, device="cuda", dtype=dtype)
    assert abs(float(K.lrelu_fwd(neg).float().mean()) + slope) < 1e-2


@pytest.mark.parametrize("name", DICT_CASES)
def test_dict_golden_step_fp32(name):
    z, cfg = load_dict_golden(name)
    with sg.use_precision("fp32"):
        g, d = build_dict_pair(cfg)
        for prefix, mod in (("g.", g), ("d.", d)):        # same weights as the reference drew
            want = golden_tensors(z, prefix)
            assert all(torch.equal(want[k], v.cpu()) for k, v in mod.state_dict().items())
        out = run_step(g, d, golden_inputs(z), cfg["alpha"])
    for k in ("d_loss", "gp", "g_loss"):
        ref = float(z["ref." + k])
        assert abs(float(out[k]) - ref) < 1e-4 * max(1.0, abs(ref)), (k, float(out[k]), ref)
    assert rel_err(out["x_fake"], torch.from_numpy(z["ref.img"])) < 1e-3
    for kind, mod in (("d_grads", d), ("g_grads", g)):
        want = golden_tensors(z, f"ref.{kind}.")
        got = {k: p.grad for k, p in mod.named_parameters()}
        assert {k for k, v in got.items() if v is not None} == set(want), kind
        for k, v in want.items():
            tol = 5e-3 if k.endswith(".bias") else 1e-3          # as tests/test_step_gpu.py::test_golden_step_fp32
            if v.numel() == 1:
                tol += 1e-6 / max(float(v.abs().max()), 1e-12)
            assert rel_err(got[k], v) < tol, (kind, k, rel_err(got[k], v))


@pytest.mark.parametrize("name", DICT_CASES)
def test_dict_golden_step_bf16(name):
    """bf16 tier: losses within 2e-3; gradients are bounded by LeakyReLU/ReLU mask flips near zero (DESIGN.md
    'Precision' item 1; no minibatch-stddev amplification in this variant)."""
    z, cfg = load_dict_golden(name)
    with sg.use_precision("bf16"):
        g, d = build_dict_pair(cfg)
        out = run_step(g, d, golden_inputs(z), cfg["alpha"])
    for k in ("d_loss", "gp"):
        ref = float(z["ref." + k])
        assert abs(float(out[k]) - ref) < 5e-3 * abs(ref), (k, float(out[k]), ref)
    assert abs(float(out["g_loss"]) - float(z["ref.g_loss"])) < 5e-3
    errs = []
    for kind, mod in (("d_grads", d), ("g_grads", g)):
        for k, v in golden_tensors(z, f"ref.{kind}.").items():
            got = dict(mod.named_parameters())[k].grad
            assert got is not None and torch.isfinite(got).all(), k
            errs.append(rel_err(got, v))
    assert np.median(errs) < 0.1 and max(errs) < 0.4, (np.median(errs), max(errs))


def test_dict_graph_step_matches_eager():
    """the whole step of a network_dict.py pair (capturable fused Adam) in a CUDA graph == the eager step from the
    same state and the same draws (losses; the weight-level noise floor is tests/test_graph_gpu.py's subject)"""
    from saragan_b200.graph import make_capturable_optimizers
    from tests.test_graph_gpu import _Recording
    z, cfg = load_dict_golden("dict_p3_lrelu")
    vol, b, alpha, warm, n = (4, 16, 16), 4, 0.5, 2, 2
    x = [torch.rand(b, 1, *vol, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) for i in range(n)]
    g1, d1 = build_dict_pair(cfg, seed=5)
    graphed = _Recording(g1, d1, *make_capturable_optimizers(g1, d1), b, vol, alpha, warmup=warm, seed=7)
    assert graphed.launches_per_step > 100 and graphed.cuda_core_conv_fallbacks == 0
    for xi in x:
        o = graphed(xi)
    loss_graph = [float(o[k]) for k in ("d_loss", "gp", "g_loss")]
    g2, d2 = build_dict_pair(cfg, seed=5)
    g_opt2, d_opt2 = make_capturable_optimizers(g2, d2)
    # (the graph's warm-up steps are undone when it is built: the eager arm starts from the fresh networks too)
    history = [(xi, graphed.draws[warm + 1 + i]) for i, xi in enumerate(x)]
    for xi, dr in history:
        oe = sg.train_step(xi, g2, d2, g_opt2, d_opt2, alpha, noise=dr["noise"], z_d=dr["z_d"], z_g=dr["z_g"], eps=dr["eps"])
    for k, got in zip(("d_loss", "gp", "g_loss"), loss_graph):
        assert np.isfinite(got) and abs(float(oe[k]) - got) < 5e-3 * max(1.0, abs(got)), (k, float(oe[k]), got)
