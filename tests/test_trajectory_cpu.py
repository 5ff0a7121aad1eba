"""The trajectory-parity protocol (tests/trajectory.py) itself, on the emulated kernels: with teacher forcing the
emulated fp32 step must reproduce the oracle's losses at every step, including the Adam state hand-over."""
import numpy as np

from tests import trajectory as T


def test_teacher_forced_protocol_on_emulated_kernels(cpu_kernels):
    cfg = dict(T.CFG, base_dim=32, latent_dim=32, num_phases=3, phase=2)
    oracle, forced, free = T.run("fp32", steps=4, cfg=cfg)
    s = T.summarize(oracle, forced, free)
    assert np.abs(forced - oracle).max() < 2e-5 * np.abs(oracle).max(), s
    assert np.abs(free - oracle).max() < 1e-3 * np.abs(oracle).max(), s       # 4 steps: no time to diverge
