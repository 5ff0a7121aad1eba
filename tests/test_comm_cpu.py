"""The N>1 gradient-exchange path on CPU: 2 gloo ranks, each with its own batch, averaged
gradients must equal the single-process gradient of the concatenated (mean) loss."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from saragan_b200.comm import GradBucketer, broadcast_parameters
    torch.manual_seed(100 + rank)                      # different init per rank -> broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3),
                              torch.nn.Linear(3, 2))
    unused = torch.nn.Linear(4, 4)                     # inactive level: never gets a gradient
    mod = torch.nn.ModuleList([net, unused])
    broadcast_parameters(mod)
    bucketer = GradBucketer(mod, bucket_bytes=64, overlap=True)   # tiny buckets -> several of them
    xs = torch.randn(4, 6, generator=torch.Generator().manual_seed(7))
    results = []
    for it in range(3):                                # pass 0 records the order, 1-2 overlap
        for p in mod.parameters():
            p.grad = None
        bucketer.arm()
        net(xs[rank * 2:(rank + 1) * 2] * (it + 1)).pow(2).mean().backward()
        bucketer.finish()
        results.append([None if p.grad is None else p.grad.clone() for p in mod.parameters()])
    if rank == 0:
        # single-process truth on the full batch with rank 0's (broadcast) weights
        truth = []
        for it in range(3):
            for p in mod.parameters():
                p.grad = None
            (0.5 * net(xs[0:2] * (it + 1)).pow(2).mean() + 0.5 * net(xs[2:4] * (it + 1)).pow(2).mean()).backward()
            truth.append([None if p.grad is None else p.grad.clone() for p in mod.parameters()])
        ok = True
        for r, t in zip(results, truth):
            for a, b in zip(r, t):
                ok &= (a is None) == (b is None)
                if a is not None:
                    ok &= bool(torch.allclose(a, b, atol=1e-6))
        q.put((ok, len(bucketer._plan)))
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, nbuckets = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert nbuckets >= 2
