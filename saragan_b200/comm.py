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

import os

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


# timing diagnosis only (tools): everything of the exchange but the collective itself -- the replicas then drift apart
_SKIP_NCCL = os.environ.get("SARAGAN_ARENA_SKIP_NCCL", "0") == "1"


class ArenaAllReduce:
    """Gradient averaging through ONE contiguous arena per network, capturable in a CUDA graph: a multi-tensor kernel
    gathers the (1/world-scaled) gradients of the active parameters into the arena (`sg_multi_copy_scale`, one launch),
    `ncclAllReduce` sums the arena in place, and the fused Adam reads its gradients straight from the arena
    (`FusedAdam.step(grads=...)`) -- no torch.cat, no scaling pass, no scatter back into `p.grad`.

        sync = ArenaAllReduce(generator, discriminator)
        ... backward ...
        grads = sync.reduce(discriminator)          # {param: arena view}; the all-reduce is enqueued on the current stream
        d_optim.step(grads=grads)

    The arena and its tables are (re)built when the set of parameters with gradients or their gradient buffers change
    (never inside a capture: build once in the eager warm-up steps)."""

    def __init__(self, generator, discriminator, comm_dtype: Optional[torch.dtype] = None, dedicated_group: bool = True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self._plans: Dict[int, dict] = {}
        # the all-reduces that end up inside a captured graph get their own communicator: eager collectives of the
        # application (barriers, metric reductions) keep using the default group
        self.group = dist.new_group() if (self.world > 1 and dedicated_group) else None
        if comm_dtype is not None:
            raise NotImplementedError("ArenaAllReduce exchanges fp32 gradients")
        broadcast_parameters(generator)
        broadcast_parameters(discriminator)

    def arm(self, module) -> None:
        pass

    def reserve_capture_tables(self) -> None:
        """Call right before capturing a step in a CUDA graph: inside the capture the gradient buffers are new, so the
        copy table is rebuilt there from pinned host memory that must exist beforehand and outlive the graph (the
        captured H2D copy re-reads it on every replay).  The arena itself -- and with it the Adam table -- is kept."""
        for plan in self._plans.values():
            if "rows" in plan:
                plan["spare"] = dict(rows=torch.empty(tuple(plan["rows"].shape), dtype=torch.int64).pin_memory())

    def _plan(self, module) -> dict:
        params = [p for p in module.parameters() if p.grad is not None]
        key = tuple((id(p), p.grad.data_ptr()) for p in params)
        plan = self._plans.get(id(module))
        if plan is not None and plan["key"] == key:
            return plan
        dev = params[0].device
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if plan is not None and [id(p) for p in plan["params"]] == [id(p) for p in params] and "rows" in plan:
            # same parameters, new gradient buffers: only the source pointers of the copy table change
            src = torch.tensor([p.grad.data_ptr() for p in params], dtype=torch.int64)
            if capturing:
                spare = plan.pop("spare", None)
                if spare is None:
                    raise RuntimeError("ArenaAllReduce: call reserve_capture_tables() before capturing the step")
                host = spare["rows"]
                host.copy_(plan["rows_host"])
                host[:, 0] = src
                # nothing pinned may be RELEASED while the stream captures (the host allocator would record its
                # reuse event inside the capture and later query it: cudaErrorInvalidValue at the next pinned
                # allocation): keep the table it replaces, and the device copy made from it, alive as well
                plan.setdefault("keepalive", []).extend([host, plan["rows_host"], plan["rows"]])
            else:
                host = plan["rows_host"].clone()
                host[:, 0] = src
                host = host.pin_memory()
            plan["rows"] = host.to(dev, non_blocking=True)
            plan["rows_host"], plan["key"] = host, key
            return plan
        if capturing:
            raise RuntimeError("ArenaAllReduce: the set of parameters with gradients changed inside a CUDA-graph "
                               "capture; run one eager step at this phase first")
        sizes = [p.numel() for p in params]
        # 16-byte aligned segments so that both the copy kernel and Adam use float4 accesses
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4
        arena = torch.zeros(total, dtype=torch.float32, device=dev)
        views = {p: arena[o:o + n].view_as(p) for p, o, n in zip(params, offs, sizes)}
        plan = dict(key=key, params=params, arena=arena, views=views)
        if dev.type == "cuda":
            rows, brow, boff = [], [], []
            for ri, (p, o, n) in enumerate(zip(params, offs, sizes)):
                rows.append([p.grad.data_ptr(), arena.data_ptr() + 4 * o, n])
                for b in range(0, n, 1024):
                    brow.append(ri), boff.append(b)
            rows_host = torch.tensor(rows, dtype=torch.int64)
            plan.update(rows_host=rows_host, rows=rows_host.to(dev), brow=torch.tensor(brow, dtype=torch.int32).to(dev),
                        boff=torch.tensor(boff, dtype=torch.int64).to(dev), n_blocks=len(brow))
        self._plans[id(module)] = plan
        return plan

    def reduce(self, module) -> Dict[torch.Tensor, torch.Tensor]:
        """Pack, all-reduce (on the current stream), return {param: averaged gradient view}."""
        plan = self._plan(module)
        scale = 1.0 / self.world
        if plan["arena"].is_cuda:
            from ._lib import call
            call("sg_multi_copy_scale", plan["rows"], plan["brow"], plan["boff"], plan["n_blocks"], float(scale))
        else:
            for p in plan["params"]:
                plan["views"][p].copy_(p.grad * scale)
        if self.world > 1 and not _SKIP_NCCL:
            dist.all_reduce(plan["arena"], op=dist.ReduceOp.SUM, group=self.group)
        return plan["views"]

    def finish(self, module) -> None:
        """comm.DataParallel-compatible form: the averages end up in `p.grad` (one extra copy; the graph path uses
        `reduce()` + `FusedAdam.step(grads=...)` instead)."""
        views = self.reduce(module)
        with torch.no_grad():
            torch._foreach_copy_([p.grad for p in views], list(views.values()))

    def finish_tensors(self, grads) -> None:
        raise NotImplementedError("ArenaAllReduce works per module: use reduce(module)")
