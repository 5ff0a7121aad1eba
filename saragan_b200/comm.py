"""Data-parallel gradient exchange: bucketed all-reduce over NCCL (NVLink 5 / NVSwitch),
overlapped with backward.  Replaces the reference's Horovod wiring
(pgan_pytorch/main.py:147-160 ``hvd.DistributedOptimizer``, train.py:36-39
``hvd.broadcast_parameters``).

One process per GPU, ``torch.distributed`` for the plumbing.  The path is pure data
parallelism: the only exchange per optimiser step is the average of the gradients of the
parameters that are ACTIVE at the current phase (network.py builds every level up front; at
phase < num_phases many parameters never get a gradient -- buckets are built from the
parameters that actually produced one).

Buckets are filled in the order gradients become ready during backward (post-accumulate-grad
hooks): as soon as a bucket is complete it is packed into a flat buffer and all-reduced on the
communication stream while the rest of backward keeps running on the compute stream;
``finish()`` waits for the outstanding work and scatters the averages back into ``p.grad``.
The gradient penalty's ``autograd.grad`` does not touch ``.grad``, so every hook fires exactly
once per ``backward()``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Start-of-phase sync of the replicas (reference: hvd.broadcast_parameters each epoch,
    train.py:36-39; replicas stay bit-identical afterwards, so once per phase suffices)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for p in module.parameters():
            dist.broadcast(p.data, src)


class GradBucketer:
    """Bucketed, overlapped gradient averaging for one module (G or D)."""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20,
                 comm_dtype: Optional[torch.dtype] = None, overlap: bool = True):
        self.module = module
        self.bucket_bytes = int(bucket_bytes)
        self.comm_dtype = comm_dtype          # e.g. torch.bfloat16 == hvd.Compression.fp16 analogue
        self.overlap = overlap
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters()]
        self._plan: Optional[List[List[int]]] = None     # bucket -> param indices, in ready order
        self._where: Dict[int, int] = {}                 # param index -> bucket
        self._order: List[int] = []                      # ready order recorded on the first pass
        self._pending: List[int] = []
        self._count: List[int] = []
        self._inflight = []                              # (bucket id, flat buffer, work, event)
        self._armed = False
        self._hooks = []
        self._comm_stream = None
        self._phase = getattr(module, "phase", None)     # the active parameter set follows `module.phase`
        if self.world > 1:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))

    # -- hooks --------------------------------------------------------------------------
    def _make_hook(self, idx: int):
        def hook(param):
            if not self._armed:
                return
            if self._plan is None:
                self._order.append(idx)
                return
            b = self._where.get(idx)
            if b is None:
                return
            self._count[b] += 1
            if self.overlap and self._count[b] == len(self._plan[b]):
                self._launch(b)
        return hook

    def arm(self) -> None:
        """Call right before ``loss.backward()`` of the step whose gradients are to be reduced."""
        if self.world == 1:
            return
        phase = getattr(self.module, "phase", None)
        if phase != self._phase:
            # network.py / network_dict.py select the active levels by `.phase` (grow() adds modules): the bucket plan
            # of the previous phase would skip the newly active parameters.  Record the ready order again on this pass
            # (its gradients are reduced in finish(), un-overlapped for this one step).
            self._phase, self._plan, self._where, self._order = phase, None, {}, []
            known = {id(p) for p in self.params}
            for p in self.module.parameters():           # parameters created by grow()
                if id(p) not in known:
                    self.params.append(p)
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(len(self.params) - 1)))
        self._armed = True
        self._inflight = []
        if self._plan is not None:
            self._count = [0] * len(self._plan)

    # -- bucket plan --------------------------------------------------------------------
    def _build_plan(self) -> None:
        plan, cur, size = [], [], 0
        for idx in self._order:
            p = self.params[idx]
            nbytes = p.numel() * 4
            if cur and size + nbytes > self.bucket_bytes:
                plan.append(cur)
                cur, size = [], 0
            cur.append(idx)
            size += nbytes
        if cur:
            plan.append(cur)
        self._plan = plan
        self._where = {idx: b for b, idxs in enumerate(plan) for idx in idxs}
        self._count = [len(idxs) for idxs in plan]

    def _launch(self, b: int) -> None:
        grads = [self.params[i].grad for i in self._plan[b]]
        dev = grads[0].device
        if dev.type == "cuda":
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ready)
                flat = torch.cat([g.reshape(-1) for g in grads])
                for g in grads:
                    g.record_stream(self._comm_stream)
                if self.comm_dtype is not None:
                    flat = flat.to(self.comm_dtype)
                flat.mul_(1.0 / self.world)
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            self._inflight.append((b, flat, work))
        else:
            flat = torch.cat([g.reshape(-1) for g in grads])
            flat.mul_(1.0 / self.world)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            self._inflight.append((b, flat, work))

    # -- completion ---------------------------------------------------------------------
    def finish(self) -> None:
        """Call after ``backward()`` and before ``optim.step()``."""
        if self.world == 1:
            return
        self._armed = False
        if self._plan is None:
            # first pass: every rank saw the same graph, so the ready order is identical
            self._build_plan()
            self._order = []
        launched = {b for b, _, _ in self._inflight}
        for b in range(len(self._plan)):
            if b not in launched:
                self._launch(b)
        for b, flat, work in self._inflight:
            grads = [self.params[i].grad for i in self._plan[b]]
            if flat.is_cuda:
                with torch.cuda.stream(self._comm_stream):
                    work.wait()
                    self._scatter(flat, grads)
            else:
                work.wait()
                self._scatter(flat, grads)
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._inflight = []

    @staticmethod
    def _scatter(flat: torch.Tensor, grads: List[torch.Tensor]) -> None:
        off = 0
        views = []
        for g in grads:
            n = g.numel()
            views.append(flat[off:off + n].view_as(g))
            off += n
        torch._foreach_copy_(grads, views)



class DataParallel:
    """The pair of bucketers ``train_step`` drives: ``arm(module)`` before ``backward()``,
    ``finish(module)`` before ``optim.step()``."""

    def __init__(self, generator, discriminator, **kw):
        self._b = {id(generator): GradBucketer(generator, **kw),
                   id(discriminator): GradBucketer(discriminator, **kw)}
        broadcast_parameters(generator)
        broadcast_parameters(discriminator)

    def arm(self, module) -> None:
        self._b[id(module)].arm()

    def finish(self, module) -> None:
        self._b[id(module)].finish()


class FlatAllReduce:
    """Gradient averaging for the CUDA-graph step: one flat all-reduce per module on the current
    stream, run eagerly between the graph segments of graph.GraphedTrainStep (no hooks: hooks do
    not fire during a graph replay).  The <= 0.5 GB of gradients cost ~1 ms un-overlapped on
    NVLink 5 against a ~30 ms step -- less than the host launch overhead the graph removes."""

    def __init__(self, generator, discriminator):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        broadcast_parameters(generator)
        broadcast_parameters(discriminator)

    def arm(self, module) -> None:
        pass

    def finish(self, module) -> None:
        self.finish_tensors([p.grad for p in module.parameters() if p.grad is not None])

    def finish_tensors(self, grads) -> None:
        """Average the given gradient tensors in place (graph.GraphedTrainStep passes the static
        gradient buffers of its capture, which `p.grad` stops pointing at once an eager step runs)."""
        if self.world == 1 or not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        flat.mul_(1.0 / self.world)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        GradBucketer._scatter(flat, grads)
