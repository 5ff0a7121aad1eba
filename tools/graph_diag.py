import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import saragan_b200 as sg
from saragan_b200.graph import make_capturable_optimizers
from tests.util import build_pair
from tests.test_graph_gpu import _Recording

cfg = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
vol, b, alpha, warm = (4, 16, 16), 4, 0.5, 2
x = torch.rand(b, 1, *vol, device="cuda")
g1, d1 = build_pair(cfg, seed=5)
go, do = make_capturable_optimizers(g1, d1)
graphed = _Recording(g1, d1, go, do, b, vol, alpha, warmup=warm, seed=7)
snapA = {k: v.detach().clone() for k, v in list(g1.named_parameters()) + list(d1.named_parameters())}
oA = graphed(x)
torch.cuda.synchronize()
draws = graphed.draws

def eager_arm():
    g2, d2 = build_pair(cfg, seed=5)
    g2o, d2o = make_capturable_optimizers(g2, d2)
    for i in range(warm):
        dr = draws[i]
        sg.train_step(torch.zeros_like(x), g2, d2, g2o, d2o, alpha, noise=dr["noise"], z_d=dr["z_d"], z_g=dr["z_g"], eps=dr["eps"])
    snap = {k: v.detach().clone() for k, v in list(g2.named_parameters()) + list(d2.named_parameters())}
    dr = draws[warm + 1]
    o = sg.train_step(x, g2, d2, g2o, d2o, alpha, noise=dr["noise"], z_d=dr["z_d"], z_g=dr["z_g"], eps=dr["eps"])
    torch.cuda.synchronize()
    return g2, d2, snap, o

g2, d2, snapB, oB = eager_arm()
g3, d3, snapC, oC = eager_arm()
print("losses graph", [float(oA[k]) for k in ("d_loss", "gp", "g_loss")])
print("losses eager", [float(oB[k]) for k in ("d_loss", "gp", "g_loss")])
names = [n for n, _ in list(g1.named_parameters())] + [n for n, _ in list(d1.named_parameters())]
PA = [p for _, p in list(g1.named_parameters()) + list(d1.named_parameters())]
PB = [p for _, p in list(g2.named_parameters()) + list(d2.named_parameters())]
PC = [p for _, p in list(g3.named_parameters()) + list(d3.named_parameters())]
print("%-34s %12s %12s | %12s %12s | %12s" % ("param", "A-B mean", "A-B max", "B-C mean", "B-C max", "warm A-B max"))
kA = list(snapA.keys()); 
for i, n in enumerate(names):
    a, bb, c = PA[i].detach(), PB[i].detach(), PC[i].detach()
    if float((a - bb).abs().max()) == 0 and float((bb - c).abs().max()) == 0: continue
    wa = list(snapA.values())[i]; wb = list(snapB.values())[i]
    print("%-34s %12.3e %12.3e | %12.3e %12.3e | %12.3e" % (("G." if i < len(list(g1.parameters())) else "D.") + n, float((a - bb).abs().mean()), float((a - bb).abs().max()),
          float((bb - c).abs().mean()), float((bb - c).abs().max()), float((wa - wb).abs().max())))
