"""Fused multi-tensor Adam (+ optional weight EMA) on the CUDA path -- SURVEY 8f row 1.

Drop-in for ``torch.optim.Adam(params, lr, betas=(0, 0.99))`` as used by the reference
(pgan_pytorch/main.py:141-142): one kernel launch updates every parameter that has a gradient
(parameters of inactive growth levels have none and are skipped, like torch does), the step
counter lives on the device so ``step()`` is CUDA-graph capturable, and an exponential moving
average of the weights (the TF path's SURFGAN_3D/ExtendedEMA.py, beta 0.99) can ride along.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import call

_CHUNK = 1024   # elements per block (256 threads x 4)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.0, 0.99), eps: float = 1e-8,
                 ema_beta: Optional[float] = None, lr_on_device: bool = False):
        """lr_on_device: the kernel reads each group's learning rate from a device scalar that `step()` refreshes
        from `group["lr"]` OUTSIDE any capture (`sync_lr()`): a CUDA graph that captured `step()` then follows an
        LR scheduler (main.py:145 LambdaLR) without being re-captured."""
        super().__init__(params, dict(lr=lr, betas=tuple(float(b) for b in betas), eps=eps))
        self.ema_beta = ema_beta
        self.lr_on_device = lr_on_device
        self._lr_dev: Dict[int, torch.Tensor] = {}
        self._lr_host: Dict[int, float] = {}
        self._step_dev: Optional[torch.Tensor] = None
        self._tables: Dict[int, dict] = {}
        self._spare: Dict[int, dict] = {}
        self._graph_keepalive = []      # pinned tables that captured H2D copies re-read on every replay

    def reserve_capture_tables(self) -> None:
        """Call right before capturing ``step()`` in a CUDA graph: inside the capture the gradient
        buffers are new, so the pointer tables are rebuilt there -- but pinned host memory cannot be
        allocated while a stream is capturing, and the captured H2D copies re-read their pinned
        source on every replay.  This pre-allocates pinned buffers that the next table build takes
        over and nothing else ever writes to."""
        for gi, group in enumerate(self.param_groups):
            n_t = len(group["params"])
            n_b = sum((p.numel() + _CHUNK - 1) // _CHUNK for p in group["params"])
            self._spare[gi] = dict(t=torch.empty((n_t, 6), dtype=torch.int64).pin_memory(),
                                   bt=torch.empty(n_b, dtype=torch.int32).pin_memory(),
                                   bo=torch.empty(n_b, dtype=torch.int64).pin_memory())

    def _state_for(self, p: torch.Tensor, beta1: float) -> dict:
        st = self.state[p]
        if "exp_avg_sq" not in st:
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg"] = torch.zeros_like(p) if beta1 > 0 else None
            st["ema"] = p.detach().clone() if self.ema_beta is not None else None
        return st

    def _table(self, gi: int, params, beta1: float, grads=None) -> dict:
        gof = (lambda p: p.grad) if grads is None else (lambda p: grads[p])
        key = tuple((p.data_ptr(), gof(p).data_ptr()) for p in params)
        tab = self._tables.get(gi)
        if tab is not None and tab["key"] == key:
            return tab
        dev = params[0].device
        rows, block_tensor, block_offset = [], [], []
        for ti, p in enumerate(params):
            st = self._state_for(p, beta1)
            if not (p.is_contiguous() and gof(p).is_contiguous() and p.dtype == torch.float32):
                raise RuntimeError("FusedAdam needs contiguous fp32 parameters and gradients")
            rows.append([p.data_ptr(), gof(p).data_ptr(), st["exp_avg"].data_ptr() if st["exp_avg"] is not None else 0,
                         st["exp_avg_sq"].data_ptr(), st["ema"].data_ptr() if st["ema"] is not None else 0, p.numel()])
            for off in range(0, p.numel(), _CHUNK):
                block_tensor.append(ti)
                block_offset.append(off)
        t_rows = torch.tensor(rows, dtype=torch.int64)
        t_bt = torch.tensor(block_tensor, dtype=torch.int32)
        t_bo = torch.tensor(block_offset, dtype=torch.int64)
        if torch.cuda.is_current_stream_capturing():
            spare = self._spare.pop(gi, None)
            if spare is None:
                raise RuntimeError("FusedAdam: call reserve_capture_tables() before capturing step() in a CUDA graph")
            host = dict(t=spare["t"][:len(rows)], bt=spare["bt"][:len(block_tensor)], bo=spare["bo"][:len(block_offset)])
            host["t"].copy_(t_rows), host["bt"].copy_(t_bt), host["bo"].copy_(t_bo)
            self._graph_keepalive.append(spare)   # must outlive the graph even if an eager step rebuilds the table
            if gi in self._tables:                # ... and nothing pinned may be released while the stream captures
                self._graph_keepalive.append(self._tables[gi])
        else:
            host = dict(t=t_rows.pin_memory(), bt=t_bt.pin_memory(), bo=t_bo.pin_memory())
        tab = dict(key=key, host=host, n_blocks=len(block_tensor),
                   t=host["t"].to(dev, non_blocking=True), bt=host["bt"].to(dev, non_blocking=True),
                   bo=host["bo"].to(dev, non_blocking=True))
        self._tables[gi] = tab
        return tab

    def sync_lr(self) -> None:
        """Copy every group's current `lr` into its device scalar (lr_on_device); call between graph replays after
        a scheduler step.  `step()` does it itself when it is not being captured."""
        for gi, group in enumerate(self.param_groups):
            dev_t = self._lr_dev.get(gi)
            if dev_t is not None and self._lr_host.get(gi) != float(group["lr"]):
                dev_t.fill_(float(group["lr"]))
                self._lr_host[gi] = float(group["lr"])

    def init_state(self, params=None) -> None:
        """Create the Adam state (and the device step counter) without taking a step -- before a CUDA-graph capture,
        where a lazy creation would be recorded and replayed (state reset on every replay)."""
        for group in self.param_groups:
            for p in (group["params"] if params is None else params):
                self._state_for(p, group["betas"][0])
                if self._step_dev is None:
                    self._step_dev = torch.zeros(4, dtype=torch.int32, device=p.device)

    @torch.no_grad()
    def step(self, closure=None, grads=None):
        """grads (optional): {parameter: gradient tensor} to use instead of `p.grad` -- the views of the all-reduced
        gradient arena of comm.ArenaAllReduce (no copy back into `p.grad`)."""
        assert closure is None
        first = None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if (p.grad is not None if grads is None else p in grads)]
            if not params:
                continue
            first = params[0]
            if self._step_dev is None:
                self._step_dev = torch.full((4,), int(getattr(self, "_pending_step", 0)), dtype=torch.int32,
                                            device=first.device)
            beta1, beta2 = group["betas"]
            tab = self._table(gi, params, beta1, grads)
            lr_dev = None
            if self.lr_on_device:
                if gi not in self._lr_dev:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("FusedAdam(lr_on_device): take one eager step (or call init_state + sync_lr) "
                                           "before capturing")
                    self._lr_dev[gi] = torch.zeros(4, dtype=torch.float32, device=first.device)
                if not torch.cuda.is_current_stream_capturing():
                    self.sync_lr()
                lr_dev = self._lr_dev[gi]
            call("sg_adam_step", tab["t"], tab["bt"], tab["bo"], tab["n_blocks"], self._step_dev, float(group["lr"]),
                 lr_dev, float(beta1), float(beta2), float(group["eps"]), float(self.ema_beta or 0.0))
            # the kernel wrote the parameters through raw pointers: tell torch (the packed-weight caches of the conv
            # layers key on the version counter -- without this they kept serving the weights of the first step)
            torch.autograd.graph.increment_version(params)
        if first is not None:
            call("sg_adam_advance", self._step_dev)

    def state_dict(self):
        """torch's layout plus the device-side step counter (`fused_step`), so that a resumed run continues the bias
        correction where it stopped.  The counter is shared by all parameters: like main.py:141-145, create one
        optimiser per growth phase (a parameter that first gets a gradient later would inherit the global t)."""
        sd = super().state_dict()
        sd["fused_step"] = int(self._step_dev[0]) if self._step_dev is not None else 0
        return sd

    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        step = int(sd.pop("fused_step", 0))
        super().load_state_dict(sd)
        self._tables.clear()          # the state tensors were replaced
        for st in self.state.values():
            if "exp_avg_sq" in st and self._step_dev is None:
                self._step_dev = torch.zeros(4, dtype=torch.int32, device=st["exp_avg_sq"].device)
        if self._step_dev is not None:
            self._step_dev.fill_(step)
        self._pending_step = step     # applied when the counter is created by the first step()

    def ema_state(self) -> Dict[torch.Tensor, torch.Tensor]:
        """parameter -> its EMA shadow (for evaluation/checkpoints, cf. ExtendedEMA.py:27-58)."""
        return {p: st["ema"] for p, st in self.state.items() if st.get("ema") is not None}
