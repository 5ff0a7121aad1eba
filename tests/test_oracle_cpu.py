"""The CPU oracle (oracle/pgan_oracle.py) against the golden fixtures minted from the unmodified reference
(oracle/pin_against_reference.py): bit-for-bit on losses, images and every parameter gradient -- the check that
travels to boxes without /root/reference."""
import numpy as np
import pytest
import torch

from oracle import pgan_oracle as O
from tests.util import golden_tensors, load_golden

FULL = ["tiny_p3", "tiny_p2_b8", "tiny_p1", "tiny_p3_b6_a1", "tiny_p2_b3_a0"]


@pytest.mark.parametrize("name", FULL)
def test_oracle_reproduces_reference_goldens(name):
    torch.set_num_threads(8)            # the fixtures were minted with 8 threads (oneDNN reductions depend on it)
    z, cfg = load_golden(name)
    st = O.TrainState(golden_tensors(z, "g."), golden_tensors(z, "d."), cfg["phase"], cfg["num_phases"])
    inp = {k: torch.from_numpy(z["in." + k]) for k in ("x_real", "noise", "z_d", "z_g", "eps")}
    got = st.step(inp["x_real"], inp["noise"], inp["z_d"], inp["eps"], inp["z_g"], cfg["alpha"], apply=False)
    for k in ("d_loss", "gp", "g_loss"):
        assert abs(got[k] - float(z["ref." + k])) <= 1e-6 * max(1.0, abs(float(z["ref." + k]))), k
    imgs = O.generator_forward(st.pg, inp["z_g"], cfg["alpha"], cfg["phase"])
    assert len(imgs) == cfg["phase"]
    for i, im in enumerate(imgs):
        assert np.allclose(im.detach().numpy(), z[f"ref.img{i}"], rtol=0, atol=1e-6), i
    for kind, names in (("d_grads", O.active_names("d", cfg["phase"], cfg["num_phases"])),
                        ("g_grads", O.active_names("g", cfg["phase"], cfg["num_phases"]))):
        want = golden_tensors(z, f"ref.{kind}.")
        assert set(want) == set(names) == {k for k, v in got[kind].items() if v is not None}
        for k, v in want.items():
            assert float((got[kind][k] - v).abs().max()) <= 1e-6 * float(v.abs().max()) + 1e-12, (kind, k)


def test_mbstd_group_rule():
    """network.py:119-124: min(4, B), bumped to the next divisor of B"""
    assert [O.mbstd_group(b) for b in (1, 2, 3, 4, 5, 6, 8, 9, 12, 16)] == [1, 2, 3, 4, 5, 6, 4, 9, 4, 4]


def test_step_flop_model_matches_survey():
    """4*Gf + 14*Df per image = 2507 GFLOP at cfg3, 55.9 at cfg1"""
    assert abs(O.step_flops_per_image(6, 7, 512, 512) / 1e9 - 2507) < 1
    assert abs(O.step_flops_per_image(3, 6, 256, 256) / 1e9 - 55.9) < 0.1


def test_flop_models_agree():
    """saragan_b200.costmodel (what bench.py reports `step_tensor_frac` with) against the oracle's count, and the
    network_dict.py model against a brute-force count over that variant's parameter shapes."""
    import sys
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    from saragan_b200 import costmodel as C
    for name, cfg in C.CONFIGS.items():
        assert abs(C.step_flops_per_image(**cfg) - O.step_flops_per_image(cfg["phase"], cfg["num_phases"], cfg["base_dim"],
                                                                          cfg["latent_dim"])) < 1.0, name
        ph = cfg["phase"]

        def level(k):
            if k.startswith("blocks.block_phase_"):
                return int(k.split("_")[2].split(".")[0])
            return ph if "_current" in k else ph - 1 if "_prev" in k else 1

        def fwd(params):
            return sum(2.0 * w.numel() * (1 if w.dim() == 2 else int(np.prod(C.volume(level(k)))))
                       for k, w in params.items() if k.endswith(".weight"))
        pg, pd = bench.dict_oracle_state(cfg)
        assert abs(C.step_flops_per_image_dict(**cfg) - (4 * fwd(pg) + 14 * fwd(pd))) < 1.0, name
