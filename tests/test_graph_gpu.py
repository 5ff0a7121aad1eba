"""The CUDA-graph replay of the step must reproduce the eager step: same weights after the same
updates from identical inputs and draws.  Building the graph must not train: its warm-up steps run on
the zero input buffer and are undone (weights, Adam state and step count restored)."""
import pytest
import torch

import saragan_b200 as sg
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers
from tests.util import build_pair

pytestmark = pytest.mark.gpu


class _Recording(GraphedTrainStep):
    """records every random draw (the warm-up steps of __init__ included)"""

    def draw(self):
        super().draw()
        if not hasattr(self, "draws"):
            self.draws = []
        self.draws.append({k: getattr(self, k).clone() for k in ("noise", "z_d", "z_g", "eps")})


CFG = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
VOL, B, ALPHA, WARM = (4, 16, 16), 4, 0.5, 2


def _eager_arm(x, draws, alphas=None, lrs=None):
    g2, d2 = build_pair(CFG, seed=5)
    g_opt2, d_opt2 = make_capturable_optimizers(g2, d2)
    # the graph's warm-up steps are undone, so the eager arm starts from the fresh networks too
    history = [(xi, draws[WARM + 1 + i]) for i, xi in enumerate(x)]
    for i, (xi, dr) in enumerate(history):
        if lrs is not None:
            for opt in (g_opt2, d_opt2):
                opt.param_groups[0]["lr"] = lrs[i]
        o = sg.train_step(xi, g2, d2, g_opt2, d_opt2, ALPHA if alphas is None else alphas[i], noise=dr["noise"],
                          z_d=dr["z_d"], z_g=dr["z_g"], eps=dr["eps"])
    torch.cuda.synchronize()
    return o, [p.detach() for p in list(g2.parameters()) + list(d2.parameters())]


class _NoSync:
    """stands in for comm.FlatAllReduce on one GPU: selects the segmented capture (what several GPUs run)"""
    world = 1

    def arm(self, module):
        pass

    def finish(self, module):
        pass

    def finish_tensors(self, grads):
        for g in grads:          # touch the buffers on the comm stream like an all-reduce would
            g.mul_(1.0)


@pytest.mark.parametrize("n_replays,segmented", [(1, False), (3, False), (3, True)])
def test_graph_replay_matches_eager(n_replays, segmented):
    """Replays against eager steps from the same state and draws.  Two EAGER runs of this step
    already differ: fp32-atomic summation order perturbs gradients at the 1e-7 level and Adam
    with beta1 = 0 turns a near-zero gradient of either sign into a +-lr step.  The graph must
    sit inside that run-to-run noise floor, measured here with a second eager arm."""
    x = [torch.rand(B, 1, *VOL, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i))
         for i in range(n_replays)]
    g1, d1 = build_pair(CFG, seed=5)
    g_opt, d_opt = make_capturable_optimizers(g1, d1)
    # warm-up steps run eagerly on the (zero) static input buffer, create the Adam state outside the graph and are
    # undone; the draw made right before the capture is consumed by no executed step
    w0 = [p.detach().clone() for p in list(g1.parameters()) + list(d1.parameters())]
    graphed = _Recording(g1, d1, g_opt, d_opt, B, VOL, ALPHA, warmup=WARM, seed=7,
                         grad_sync=_NoSync() if segmented else None)
    assert (graphed.segments is not None) == segmented
    for a, b in zip(w0, list(g1.parameters()) + list(d1.parameters())):
        assert torch.equal(a, b)                     # building the graph did not train
    assert int(d_opt._step_dev[0]) == 0 and int(g_opt._step_dev[0]) == 0
    for xi in x:
        o = graphed(xi)
    loss_graph = [float(o[k]) for k in ("d_loss", "gp", "g_loss")]
    assert len(graphed.draws) == WARM + 1 + len(x)
    pa = [p.detach() for p in list(g1.parameters()) + list(d1.parameters())]
    ob, pb = _eager_arm(x, graphed.draws)
    oc, pc = _eager_arm(x, graphed.draws)
    for k, got in zip(("d_loss", "gp", "g_loss"), loss_graph):
        assert abs(float(ob[k]) - got) < 2e-3 * max(1.0, abs(got)), k
    graph_vs_eager = sum(float((a - b).abs().mean()) for a, b in zip(pa, pb)) / len(pa)
    eager_vs_eager = sum(float((b - c).abs().mean()) for b, c in zip(pb, pc)) / len(pa)
    # single samples of a heavy-tailed quantity (sign flips of +-lr steps): allow a generous factor
    assert graph_vs_eager < 6 * eager_vs_eager + 1e-4, (graph_vs_eager, eager_vs_eager)


def test_one_capture_follows_alpha_and_lr_schedules():
    """alpha (train.py:33,63) and the learning rate (main.py:145 LambdaLR) change between replays of ONE captured
    graph -- both live in device scalars the kernels read -- and the result equals eager steps with those values."""
    n = 3
    alphas, lrs = [1.0, 0.6, 0.2], [1e-3, 7e-4, 4e-4]
    x = [torch.rand(B, 1, *VOL, device="cuda", generator=torch.Generator(device="cuda").manual_seed(10 + i))
         for i in range(n)]
    g1, d1 = build_pair(CFG, seed=5)
    g_opt, d_opt = make_capturable_optimizers(g1, d1)
    graphed = _Recording(g1, d1, g_opt, d_opt, B, VOL, alphas[0], warmup=WARM, seed=7)
    outs = []
    for i, xi in enumerate(x):
        graphed.set_alpha(alphas[i])
        for opt in (g_opt, d_opt):
            opt.param_groups[0]["lr"] = lrs[i]
        o = graphed(xi)
        outs.append([float(o[k]) for k in ("d_loss", "gp", "g_loss")])
    pa = [p.detach() for p in list(g1.parameters()) + list(d1.parameters())]
    ob, pb = _eager_arm(x, graphed.draws, alphas, lrs)
    oc, pc = _eager_arm(x, graphed.draws, alphas, lrs)
    wrong, _ = _eager_arm(x, graphed.draws, [alphas[0]] * n, [lrs[0]] * n)       # what a graph with baked-in values would do
    for k, got in zip(("d_loss", "gp", "g_loss"), outs[-1]):
        assert abs(float(ob[k]) - got) < 2e-3 * max(1.0, abs(got)), k
    assert abs(float(wrong["d_loss"]) - outs[-1][0]) > 10 * abs(float(ob["d_loss"]) - outs[-1][0]) + 1e-4
    graph_vs_eager = sum(float((a - b).abs().mean()) for a, b in zip(pa, pb)) / len(pa)
    eager_vs_eager = sum(float((b - c).abs().mean()) for b, c in zip(pb, pc)) / len(pa)
    assert graph_vs_eager < 6 * eager_vs_eager + 1e-4, (graph_vs_eager, eager_vs_eager)
    graphed.close()                        # releases the captured graph(s); the weights it trained stay
    assert graphed.graph is None and torch.isfinite(pa[0]).all()
