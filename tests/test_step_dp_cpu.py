"""The N > 1 path of the train step itself on CPU: 2 gloo ranks, emulated kernels, each rank its own half of the
global batch and its own draws; after `train_step(..., grad_sync=...)` every rank must hold the AVERAGE of the two
ranks' single-process gradients (minibatch-stddev is per-rank local, as in the reference: network.py:118-133 never
syncs it), for both gradient-exchange implementations (hook-driven buckets and the flat all-reduce the graph step
uses) and for both network variants.  Inactive levels (no gradient at this phase) must stay without gradient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(phase=2, num_phases=3, base_dim=32, latent_dim=32, base_shape=(1, 1, 4, 4), batch=4)


def _build(variant, seed):
    import saragan_b200 as sg
    torch.manual_seed(seed)
    args = (CFG["phase"], CFG["num_phases"], CFG["base_dim"], CFG["latent_dim"], CFG["base_shape"])
    if variant == "network_dict":
        from saragan_b200 import network_dict as nd
        return nd.Generator(*args, "leaky_relu", param=0.3), nd.Discriminator(*args, "leaky_relu", param=0.3)
    return sg.Generator(*args), sg.Discriminator(*args)


def _grads(g, d):
    return [None if p.grad is None else p.grad.clone() for p in list(g.parameters()) + list(d.parameters())]


def _worker(rank, world, port, variant, sync, grow, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import saragan_b200 as sg
    from saragan_b200 import comm, kernels
    from tests import cpu_emul
    from tests.util import draw_inputs
    for name in cpu_emul.ALL:
        setattr(kernels, name, getattr(cpu_emul, name))
    sg.set_precision("fp32")
    g, d = _build(variant, seed=100 + rank)               # different init per rank: the broadcast must fix it
    if sync == "flat":
        dp = comm.FlatAllReduce(g, d)
    elif sync == "arena":      # the gradient arena of the single-graph multi-GPU step
        dp = comm.ArenaAllReduce(g, d)
    else:       # "buckets_bf16": bf16 payload, the analogue of hvd.Compression.fp16 (main.py:149-150)
        dp = comm.DataParallel(g, d, bucket_bytes=1 << 16, comm_dtype=torch.bfloat16 if sync == "buckets_bf16" else None)
    opts = sg.make_optimizers(g, d)
    inputs = [draw_inputs(CFG, seed=500 + r) for r in range(world)]
    mine = inputs[rank]
    kw = lambda i: dict(noise=i["noise"], z_d=i["z_d"], z_g=i["z_g"], eps=i["eps"], apply=False)     # noqa: E731
    sg.train_step(mine["x_real"], g, d, *opts, 0.5, grad_sync=dp, **kw(mine))
    if grow:
        # the next growth phase on the same replicas: more parameters become active (network.py) / are created
        # (network_dict.py grow()); the exchange must follow
        if variant == "network_dict":
            torch.manual_seed(7)                          # same new weights on both ranks, as main.py's deepcopy + grow
            g.grow()
            d.grow()
            opts = sg.make_optimizers(g, d)
        else:
            g.phase = d.phase = CFG["phase"] + 1
        cfg3 = dict(CFG, phase=CFG["phase"] + 1)
        inputs = [draw_inputs(cfg3, seed=600 + r) for r in range(world)]
        mine = inputs[rank]
        for p in list(g.parameters()) + list(d.parameters()):
            p.grad = None
        sg.train_step(mine["x_real"], g, d, *opts, 0.5, grad_sync=dp, **kw(mine))
    got = _grads(g, d)
    if rank == 0:
        truth = None
        for i in inputs:                                  # single-process gradients of each rank's batch, averaged
            for p in list(g.parameters()) + list(d.parameters()):
                p.grad = None
            sg.train_step(i["x_real"], g, d, *opts, 0.5, **kw(i))
            gr = _grads(g, d)
            truth = gr if truth is None else [None if a is None else a + b for a, b in zip(truth, gr)]
        ok, n_active = True, 0
        for a, t in zip(got, truth):
            ok &= (a is None) == (t is None)
            if a is not None:
                n_active += 1
                if sync == "buckets_bf16":
                    ok &= float((a - t / world).norm()) <= 8e-3 * float((t / world).norm()) + 1e-9
                else:
                    ok &= bool(torch.allclose(a, t / world, rtol=1e-4, atol=1e-7))
        q.put((ok, n_active, len(got)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("variant,sync,grow", [("network", "buckets", False), ("network", "flat", False),
                                               ("network_dict", "buckets", False), ("network", "buckets_bf16", False),
                                               ("network", "buckets", True),
                                               ("network_dict", "buckets", True),
                                               ("network", "arena", False), ("network", "arena", True)])
def test_train_step_two_ranks_gloo(variant, sync, grow):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + hash((variant, sync, grow))) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, variant, sync, grow, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_active, n_total = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    all_active = variant == "network_dict" or grow        # phase 3 of 3: every level of network.py is active
    assert 0 < n_active and (n_active == n_total if all_active else n_active < n_total)
