"""SURVEY 8f row 3 -- the ``network_dict.py`` API variant: the restatement in oracle/pgan_dict_oracle.py against
the fixtures minted from the unmodified reference, and the drop-in modules' host logic (names, RNG stream at
construction incl. grow(), autograd wiring through the GP double backward) on the emulated kernels."""
import pytest
import torch

import saragan_b200 as sg
from oracle import pgan_dict_oracle as OD
from tests.dict_util import DICT_CASES, build_dict_pair, golden_inputs, load_dict_golden
from tests.util import golden_tensors, rel_err, run_step


@pytest.mark.parametrize("name", DICT_CASES)
def test_dict_oracle_matches_reference_goldens(name):
    z, cfg = load_dict_golden(name)
    st = OD.DictTrainState(golden_tensors(z, "g."), golden_tensors(z, "d."), cfg["phase"], cfg["nonlinearity"], cfg["param"])
    inp = golden_inputs(z)
    got = st.step(inp["x_real"], inp["noise"], inp["z_d"], inp["eps"], inp["z_g"], cfg["alpha"], apply=False)
    for k in ("d_loss", "gp", "g_loss"):
        assert got[k] == float(z["ref." + k]), k
    assert torch.equal(got["img"], torch.from_numpy(z["ref.img"]))
    for kind in ("d_grads", "g_grads"):
        want = golden_tensors(z, f"ref.{kind}.")
        assert {k for k, v in got[kind].items() if v is not None} == set(want)
        for k, v in want.items():
            assert torch.equal(got[kind][k], v), (kind, k)


@pytest.mark.parametrize("name", DICT_CASES)
def test_dict_init_matches_reference_rng_stream(name, cpu_kernels):
    """same state_dict names, shapes and VALUES as the reference modules built (and grown) from the same seed"""
    z, cfg = load_dict_golden(name)
    g, d = build_dict_pair(cfg)
    for prefix, mod in (("g.", g), ("d.", d)):
        want = golden_tensors(z, prefix)
        got = mod.state_dict()
        assert set(want) == set(got)
        for k in want:
            assert torch.equal(want[k], got[k].cpu()), k
    assert g.phase == d.phase == cfg["phase"]


@pytest.mark.parametrize("name", DICT_CASES)
def test_dict_step_matches_reference_fp32(name, cpu_kernels):
    z, cfg = load_dict_golden(name)
    with sg.use_precision("fp32"):
        g, d = build_dict_pair(cfg)
        assert abs(cpu_kernels.get_leaky_slope() - (0.0 if cfg["nonlinearity"] == "relu" else 0.3)) < 1e-7
        out = run_step(g, d, golden_inputs(z), cfg["alpha"])
    for k in ("d_loss", "gp", "g_loss"):
        ref = float(z["ref." + k])
        assert abs(float(out[k]) - ref) < 2e-5 * max(1.0, abs(ref)), k
    assert rel_err(out["x_fake"], torch.from_numpy(z["ref.img"])) < 1e-5
    for kind, mod in (("d_grads", d), ("g_grads", g)):
        want = golden_tensors(z, f"ref.{kind}.")
        got = {k: p.grad for k, p in mod.named_parameters()}
        assert {k for k, v in got.items() if v is not None} == set(want), kind
        for k, v in want.items():
            tol = 1e-4 if v.numel() > 1 else 1e-4 + 2e-6 / max(float(v.abs().max()), 1e-12)
            assert rel_err(got[k], v) < tol, (kind, k, rel_err(got[k], v))


def test_dict_swish_is_unsupported_like_in_the_reference(cpu_kernels):
    """network_dict.py's 'swish' cannot be constructed: calculate_gain('swish') raises ValueError."""
    from saragan_b200 import network_dict as nd
    with pytest.raises(ValueError, match="Unsupported nonlinearity"):
        nd.Generator(1, 3, 32, 32, (1, 1, 4, 4), "swish")


def test_slope_switches_back_for_network_py(cpu_kernels):
    """the slope is process-wide in the kernel library: a network.py model after a network_dict.py one sees 0.2"""
    from saragan_b200 import network_dict as nd
    nd.Discriminator(1, 3, 32, 32, (1, 1, 4, 4), "leaky_relu", param=0.3)
    assert abs(cpu_kernels.get_leaky_slope() - 0.3) < 1e-7
    with sg.use_precision("fp32"):
        d = sg.Discriminator(1, 3, 32, 32, (1, 1, 4, 4))
        d(torch.randn(4, 1, 1, 4, 4), 0.0)
    assert abs(cpu_kernels.get_leaky_slope() - 0.2) < 1e-7
