"""100-step loss trajectories of the CUDA step against the fp32 CPU oracle (BASELINE.json north star: "G/D/GP loss
trajectories over 100 steps within 1 %"), protocol in tests/trajectory.py.

Measured on a B200 (tools/trajectory_probe.py, profiles/r1_trajectory_100steps.json), 3D PGAN phase 3 of 4, base_dim 64,
4x16x16, B = 4, lr 1e-3, alpha 0.5; the oracle's d_loss runs 9.8 -> 7.5, gp 9.8 -> 4.2, g_loss -0.06 -> -5.9:

                      teacher-forced, worst of 100 steps        free-running, max |dev| / range of the trajectory
    fp32 mode         6e-7 (d_loss), 2e-7 (gp), 6e-7 (g_loss)   6 %, 10 %, 7 %   (last 10 steps: 2 %, 1 %, 3 %)
    bf16 mode         1.0 %,         2.6 %,     0.4 %           3 %,  5 %, 5 %   (last 10 steps: 1.4 %, 0.9 %, 1.6 %)

Teacher-forced, every step of the fp32 mode reproduces the oracle to fp32 rounding and the bf16 mode stays at the
1 % level (its worst single step is 2.6 % on the gradient penalty, a difference of squares of per-sample gradient
norms).  Free-running, even the fp32 mode -- whose every step is exact to 6e-7 -- drifts by up to 10 % of the range:
that drift is the dynamics amplifying rounding noise, not an implementation error, which is why the
assertions on it are loose sanity bounds and the tight ones are on the teacher-forced arm."""
import pytest

from tests import trajectory as T

pytestmark = pytest.mark.gpu


# teacher-forced bounds: north star "within 1 %" for the benchmarked policy and the TF32 mode (at 4x16x16 both run
# TF32 tensor cores + the exact base level); the exact fp32 mode reproduces the oracle to rounding
@pytest.mark.parametrize("precision,worst,median", [("fp32", 1e-4, 1e-5), ("tf32", 1e-2, 3e-3), ("bf16", 1e-2, 3e-3)])
def test_100_step_trajectory(precision, worst, median):
    s = T.summarize(*T.run(precision, steps=100))
    print(precision, s)
    for k, v in s.items():
        assert v["teacher_forced_rel"] < worst, (k, v)
        assert v["teacher_forced_median_rel"] < median, (k, v)
        assert v["free_running_range_norm"] < 0.3 and v["free_running_last10_range_norm"] < 0.15, (k, v)
    # the trajectory actually goes somewhere (a frozen model would pass everything above)
    assert s["gp"]["oracle_range"] > 2.0 and abs(s["g_loss"]["oracle_last"] - s["g_loss"]["oracle_first"]) > 1.0


def test_100_step_trajectory_with_a_bf16_level():
    """The same protocol one growth phase up (phase 4 of 4: 8x32x32, 16 -> 32 channels at the top), where the top
    level of both networks runs the bf16 kernels under the benchmarked policy: teacher-forced, all three losses within
    1 % at every one of the 100 steps."""
    cfg = dict(T.CFG, phase=4)
    s = T.summarize(*T.run("bf16", steps=100, cfg=cfg))
    print("bf16 phase 4", s)
    for k, v in s.items():
        assert v["teacher_forced_rel"] < 1e-2 and v["teacher_forced_median_rel"] < 3e-3, (k, v)
        # (free-running, this configuration separates further than phase 3 -- measured 0.5 of the range on d_loss with
        # every single step exact to 2e-3: the dynamics, not the arithmetic; a loose sanity bound only)
        assert v["free_running_range_norm"] < 0.8, (k, v)
