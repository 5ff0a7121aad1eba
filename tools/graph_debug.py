import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import saragan_b200 as sg
from saragan_b200 import costmodel as C
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers

class Noop:
    def arm(self, m): pass
    def finish(self, m): pass
    def finish_tensors(self, t): pass

ap = argparse.ArgumentParser()
ap.add_argument("--phase", type=int); ap.add_argument("--num-phases", type=int); ap.add_argument("--base-dim", type=int)
ap.add_argument("--latent", type=int, default=256); ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--fused", type=int, default=1); ap.add_argument("--segments", type=int, default=1)
a = ap.parse_args()
vol = C.volume(a.phase)
torch.manual_seed(0)
g = sg.Generator(a.phase, a.num_phases, a.base_dim, a.latent, C.BASE_SHAPE)
d = sg.Discriminator(a.phase, a.num_phases, a.base_dim, a.latent, C.BASE_SHAPE)
go, do = make_capturable_optimizers(g, d, fused=bool(a.fused))
gs = GraphedTrainStep(g, d, go, do, a.batch, vol, 0.5, warmup=2, seed=1, grad_sync=Noop() if a.segments else None)
x = torch.rand(a.batch, 1, *vol, device="cuda")
gs.x.copy_(x); gs.draw()
try:
  for rep in range(3):
    if gs.segments is None:
        gs.graph.replay(); torch.cuda.synchronize(); print("single graph ok", rep)
    else:
        for i, seg in enumerate(gs.segments):
            seg.replay(); torch.cuda.synchronize(); print("segment", i, "ok", rep)
  if True:
    print("losses", float(gs.out["d_loss"]), float(gs.out["g_loss"]))
    pool = [torch.rand(a.batch, 1, *vol).pin_memory() for _ in range(2)]
    for i in range(4):
        o = gs(pool[i % 2])
    torch.cuda.synchronize(); print("call path ok", float(o["d_loss"]))
    for i in range(2):     # eager steps on the same modules/optimisers, then replays again (bench.py's flow)
        sg.train_step(x, g, d, go, do, 0.5)
    import gc; gc.collect()
    junk = [torch.empty(1 << 20).pin_memory() for _ in range(8)]
    for i in range(3):
        o = gs(pool[i % 2])
    torch.cuda.synchronize(); print("replay after eager steps ok", float(o["d_loss"]))
except Exception as e:
    print("FAILED:", str(e)[:200])
