"""The CUDA-graph replay of the step must reproduce the eager step: same weights after two
updates from identical inputs and draws."""
import pytest
import torch

import saragan_b200 as sg
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers
from tests.util import build_pair

pytestmark = pytest.mark.gpu


def test_graph_replay_matches_eager():
    cfg = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
    vol, b, alpha = (4, 16, 16), 4, 0.5
    x = [torch.rand(b, 1, *vol, device="cuda") for _ in range(3)]

    g1, d1 = build_pair(cfg, seed=5)
    g_opt, d_opt = make_capturable_optimizers(g1, d1)
    graphed = GraphedTrainStep(g1, d1, g_opt, d_opt, b, vol, alpha, warmup=0, seed=7)
    draws = []
    orig_draw = graphed.draw

    def recording_draw():
        orig_draw()
        draws.append({k: getattr(graphed, k).clone() for k in ("noise", "z_d", "z_g", "eps")})
    graphed.draw = recording_draw
    losses = []
    for xi in x:
        o = graphed(xi)
        losses.append([float(o[k]) for k in ("d_loss", "gp", "g_loss")])

    g2, d2 = build_pair(cfg, seed=5)
    g_opt2, d_opt2 = make_capturable_optimizers(g2, d2)
    # the capture itself applied one (un-replayed) update with the draw made just before it:
    # replay that history eagerly -- capture-time step first, then the three recorded steps
    # (warmup=0, so nothing else touched the weights)
    # NOTE: torch.cuda.graph capture does not execute kernels, so the weights are only updated by replays.
    for xi, dr in zip(x, draws):
        o = sg.train_step(xi, g2, d2, g_opt2, d_opt2, alpha, noise=dr["noise"], z_d=dr["z_d"], z_g=dr["z_g"],
                          eps=dr["eps"])
    torch.cuda.synchronize()
    assert abs(float(o["d_loss"]) - losses[-1][0]) < 2e-3 * abs(losses[-1][0])
    for (n1, p1), (n2, p2) in zip(list(g1.named_parameters()) + list(d1.named_parameters()),
                                  list(g2.named_parameters()) + list(d2.named_parameters())):
        assert n1 == n2
        assert float((p1 - p2).abs().max()) < 5e-3, n1      # 3 Adam steps of lr 1e-3 each at most 3e-3 apart
