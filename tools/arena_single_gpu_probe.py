"""Debug probe (one GPU): which host-side operation fails after a captured step that used comm.ArenaAllReduce."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import saragan_b200 as sg
from saragan_b200 import comm
from saragan_b200.graph import GraphedTrainStep, make_capturable_optimizers
CFG = dict(phase=3, num_phases=4, base_dim=64, latent_dim=64, base_shape=(1, 1, 4, 4))
VOL, B = (4, 16, 16), 4
mode = sys.argv[1]
torch.manual_seed(5)
g = sg.Generator(CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
d = sg.Discriminator(CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
g_opt, d_opt = make_capturable_optimizers(g, d)
dp = comm.ArenaAllReduce(g, d) if mode != "none" else None
graphed = GraphedTrainStep(g, d, g_opt, d_opt, B, VOL, 0.5, warmup=2, seed=1, grad_sync=dp)
def attempt(name, fn):
    try:
        fn(); torch.cuda.synchronize(); print(mode, name, "OK", flush=True)
    except Exception as e:
        print(mode, name, "FAILED", str(e)[:80], flush=True)
attempt("sync after capture", lambda: None)
attempt("pinned alloc", lambda: torch.empty(16).pin_memory())
attempt("cpu copy", lambda: torch.ones(4, device="cuda").cpu())
attempt("item", lambda: torch.ones(4, device="cuda").sum().item())
for i in range(2):
    o = graphed(torch.rand(B, 1, *VOL, device="cuda"))
attempt("sync after replays", lambda: None)
attempt("pinned alloc", lambda: torch.empty(16).pin_memory())
attempt("cpu copy", lambda: torch.ones(4, device="cuda").cpu())
attempt("item", lambda: torch.ones(4, device="cuda").sum().item())
attempt("equal", lambda: torch.equal(torch.ones(8, device="cuda"), torch.ones(8, device="cuda")))
