"""Whole-step CUDA graph: one D update (with the WGAN-GP double backward) + one G update +
both Adam steps captured once and replayed per batch.

The step launches ~700 of our kernels plus a few hundred tiny torch ops (autograd bookkeeping,
minibatch-stddev, Adam); issued one by one from Python they cost more host time than the GPU
needs to execute them once the convolutions run on tcgen05.  Capturing the step removes the
host from the loop (the north-star's "CUDA streams and graphs instead of a tracing compiler").

On several GPUs the step is still ONE graph: with comm.ArenaAllReduce the NCCL all-reduces are captured inside it
(comm.FlatAllReduce keeps round 1's form, four segments with an eager all-reduce between them) -- see __init__.

Inputs live in static buffers: the real batch, the instance-noise draw, the two latent draws and
the GP interpolation draw are refreshed outside the graph (`draw()`), then `replay()` runs the
captured work.  `alpha` (train.py:33,63: it changes once per epoch) and the learning rate (main.py:145
LambdaLR) live in DEVICE scalars the captured kernels read, so one capture serves a whole growth phase:
`set_alpha()` / the optimisers' `param_groups[...]["lr"]` take effect at the next replay.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import kernels
from .train import d_phase, g_phase, top_image, train_step


def make_capturable_optimizers(generator, discriminator, lr: float = 1e-3, world_size: int = 1,
                               fused: bool = True, ema_beta: Optional[float] = None):
    """main.py:138-145's Adam(lr*sqrt(world), betas=(0, .99)) in a form whose `step()` can live inside
    a CUDA graph: our fused multi-tensor kernel (optim.FusedAdam, device-side step counter; optional
    generator-weight EMA), or torch's capturable Adam with fused=False."""
    lr = lr * float(world_size) ** 0.5
    if fused:
        from .optim import FusedAdam
        d_optim = FusedAdam(discriminator.parameters(), lr=lr, betas=(0.0, 0.99), lr_on_device=True)
        g_optim = FusedAdam(generator.parameters(), lr=lr, betas=(0.0, 0.99), ema_beta=ema_beta, lr_on_device=True)
    else:
        d_optim = torch.optim.Adam(discriminator.parameters(), lr=lr, betas=(0.0, 0.99), capturable=True)
        g_optim = torch.optim.Adam(generator.parameters(), lr=lr, betas=(0.0, 0.99), capturable=True)
    return g_optim, d_optim


def _snapshot(nets, opts):
    """Copies of every parameter and of every tensor of the optimisers' state (plus FusedAdam's device counter)."""
    snap = {"p": [(p, p.detach().clone()) for net in nets for p in net.parameters()], "o": []}
    for opt in opts:
        st = {p: {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
              for p, d in opt.state.items()}
        snap["o"].append((opt, st, None if getattr(opt, "_step_dev", None) is None else opt._step_dev.clone()))
    return snap


@torch.no_grad()
def _restore(snap):
    """Undo the warm-up steps: parameters back to their values, optimiser state back to what it was -- state the
    warm-up CREATED is kept allocated (a capture must not create it) but reset to its initial value (zeros, step 0,
    EMA = the weights)."""
    for p, v in snap["p"]:
        p.copy_(v)
        p.grad = None
    for opt, st, step_dev in snap["o"]:
        for p, d in opt.state.items():
            old = st.get(p, {})
            for k, v in d.items():
                if not isinstance(v, torch.Tensor):
                    continue
                if isinstance(old.get(k), torch.Tensor):
                    v.copy_(old[k])
                elif k == "ema":
                    v.copy_(p)
                else:
                    v.zero_()
        if getattr(opt, "_step_dev", None) is not None:
            if step_dev is not None:
                opt._step_dev.copy_(step_dev)
            else:
                opt._step_dev.zero_()
    for p, _ in snap["p"]:
        torch.autograd.graph.increment_version(p)      # packed-weight caches key on the version counter


class GraphedTrainStep:
    def __init__(self, generator, discriminator, g_optim, d_optim, batch: int, volume, alpha: float,
                 warmup: int = 3, seed: Optional[int] = None, grad_sync=None):
        """The warm-up steps (lazy initialisation, kernel attributes, allocator state, Adam state creation) run on
        the all-zero static input; weights and optimiser state are restored afterwards, so building the graph does
        not train on garbage or advance Adam's step count."""
        if warmup < 1:
            # the first optim.step() creates the Adam state; inside a capture that initialisation
            # would be replayed (state reset to zero) on every step
            raise ValueError("GraphedTrainStep needs at least one eager warm-up step before the capture")
        self.grad_sync = grad_sync    # e.g. comm.CapturableAllReduce on several GPUs
        self.g, self.d, self.g_optim, self.d_optim = generator, discriminator, g_optim, d_optim
        dev = discriminator.device
        self.dev = dev
        # alpha as a device scalar: the blend kernels read it (ops.blend_coef), set_alpha() changes it between replays
        self.alpha = torch.full((), float(alpha), dtype=torch.float32, device=dev)
        latent = generator.latent_dim
        self.x = torch.zeros((batch, 1, *volume), device=dev)
        self.noise = torch.zeros_like(self.x)
        self.z_d = torch.zeros((batch, latent), device=dev)
        self.z_g = torch.zeros((batch, latent), device=dev)
        self.eps = torch.zeros((batch, 1, 1, 1, 1), device=dev)
        self.rng = torch.Generator(device=dev)
        if seed is not None:
            self.rng.manual_seed(seed)
        self.out: Dict[str, torch.Tensor] = {}
        self._raw = None
        self.launch_desc = "cuda-graph replay of the whole step"
        self.graph = torch.cuda.CUDAGraph()
        # warm-up on a side stream (lazy inits, kernel attributes, allocator state), then capture
        rng_state = self.rng.get_state()
        snap = _snapshot((generator, discriminator), (g_optim, d_optim))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.draw()
                self._step()
            _restore(snap)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        del snap
        self.rng.set_state(rng_state)
        # the packings were made OUTSIDE the graph: mark them stale, so that the first `prepack` inside the capture
        # re-packs every weight (into the same persistent buffers, one captured launch per network)
        for net in (generator, discriminator):
            for m in net.modules():
                if hasattr(m, "_packed"):
                    m._packed._ver = {}
        self.draw()
        for opt in (g_optim, d_optim):
            if hasattr(opt, "reserve_capture_tables"):
                opt.reserve_capture_tables()
        from . import _lib
        n0 = _lib.launch_count()
        f0 = int(_lib.load().sg_cuda_core_fallbacks(0))
        if hasattr(grad_sync, "reserve_capture_tables"):
            grad_sync.reserve_capture_tables()
        if grad_sync is None:
            with torch.cuda.graph(self.graph):
                self.out = self._step()
            self.segments = None
        elif hasattr(grad_sync, "reduce"):
            # several GPUs, gradient arena (comm.ArenaAllReduce): ONE graph with the NCCL all-reduces inside.  The D
            # gradients' all-reduce is a fork of the graph that runs beside the G update's generator forward; only
            # the G gradients' all-reduce (and two pack kernels) sit on the critical path.  NCCL's watchdog thread
            # polls CUDA events while we capture: capture_error_mode="thread_local" keeps its calls out of our capture.
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.out = self._step()
            self.segments = None
            self.launch_desc = ("one cuda-graph replay per step with the NCCL all-reduces captured inside (D-gradient "
                                "all-reduce forked beside the generator forward; gradient arena read by the fused Adam)")
        else:
            # several GPUs: the gradient all-reduce runs EAGERLY between graph segments
            #   g1: D forward/backward | g2a: the G update's generator forward | g2b: D update + D(fakes) +
            #   backward | g3: G update
            # that share one memory pool, so NCCL never has to be captured; the gradients are static buffers of
            # the pool.  g2a needs nothing the D update changes, so it replays WHILE the D gradients are being
            # averaged on a side stream; only the G gradients' all-reduce is exposed.
            g1, g2a, g2b, g3 = self.graph, torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                od = d_phase(self.x, self.g, self.d, self.d_optim, self.alpha, noise=self.noise, z_d=self.z_d,
                             eps=self.eps)
            # the gradient buffers the captured backward writes (p.grad is re-bound by any later eager step)
            self._d_grads = [p.grad for p in self.d.parameters() if p.grad is not None]
            with torch.cuda.graph(g2a, pool=g1.pool()):
                self.g.train()
                for p in self.g.parameters():
                    p.requires_grad = True
                x_fake_g = top_image(self.g(self.z_g, self.alpha))
            with torch.cuda.graph(g2b, pool=g1.pool()):
                self.d_optim.step()
                og = g_phase(self.x.shape[0], self.g, self.d, self.g_optim, self.alpha, z_g=self.z_g, x_fake=x_fake_g)
                dist_ = od["d_real_mean"] - og["d_fake_mean"]
            self._g_grads = [p.grad for p in self.g.parameters() if p.grad is not None]
            with torch.cuda.graph(g3, pool=g1.pool()):
                self.g_optim.step()
            self.segments = (g1, g2a, g2b, g3)
            self.launch_desc = ("4 cuda-graph segments + eager NCCL all-reduce (D-gradient all-reduce overlapped with "
                                "the generator forward)")
            self._comm_stream = torch.cuda.Stream(device=dev)
            self.out = {"d_loss": od["d_loss"], "gp": od["gp"], "g_loss": og["g_loss"], "distance": dist_,
                        "x_fake": og["x_fake"]}
        # the LeakyReLU slope is a constant of the kernel library read at run time: a replay must see the value the
        # capture's forward passes set (another model variant may have changed it in between)
        self._leaky_slope = kernels.get_leaky_slope()
        self.launches_per_step = _lib.launch_count() - n0   # our kernels inside one replay
        self.cuda_core_conv_fallbacks = int(_lib.load().sg_cuda_core_fallbacks(0)) - f0

    def _step(self):
        o = train_step(self.x, self.g, self.d, self.g_optim, self.d_optim, self.alpha, noise=self.noise,
                       z_d=self.z_d, z_g=self.z_g, eps=self.eps, grad_sync=self.grad_sync)
        return {k: o[k] for k in ("d_loss", "gp", "g_loss", "distance", "x_fake")}

    def set_alpha(self, alpha: float) -> None:
        """New fade-in coefficient for the following replays (train.py:33,63: once per epoch); no re-capture."""
        self.alpha.fill_(float(alpha))

    def close(self) -> None:
        """Release the captured graph(s).  With NCCL captured inside (gradient arena, N > 1) this MUST run before
        `torch.distributed.destroy_process_group()`: a live graph keeps the communicator's work alive and the
        communicator teardown then waits for ever (measured on 2 x B200)."""
        torch.cuda.synchronize(self.dev)
        for gr in {id(x): x for x in (self.segments or ()) + (self.graph,)}.values():
            gr.reset()
        self.segments, self.graph = None, None

    def draw(self) -> None:
        """Fresh random draws of train.py:144-145,178 and loss.py:11 into the static buffers."""
        self.noise.normal_(generator=self.rng)
        self.z_d.normal_(generator=self.rng)
        self.z_g.normal_(generator=self.rng)
        self.eps.uniform_(generator=self.rng)

    def __call__(self, x_real: torch.Tensor) -> Dict[str, torch.Tensor]:
        """x_real: host (pinned) or device batch.  Returns the step's scalars as device tensors
        (valid until the next call)."""
        kernels.ensure_leaky_slope(self._leaky_slope)
        for opt in (self.g_optim, self.d_optim):
            if hasattr(opt, "sync_lr"):
                opt.sync_lr()                  # a scheduler may have changed group["lr"] since the last replay
        if x_real.dtype == torch.uint16:
            # raw dataset voxels (data.VolumeLoader batches / pinned host memory): H2D of 2 bytes per voxel, then
            # cast and 1/1024 scale in one kernel straight into the graph's input buffer (main.py:85-87)
            from .data import prepare_real
            if self._raw is None or self._raw.shape != x_real.shape:
                self._raw = torch.empty(x_real.shape, dtype=torch.uint16, device=self.dev)
            self._raw.copy_(x_real, non_blocking=True)
            prepare_real(self._raw, None, out=self.x)
        else:
            self.x.copy_(x_real, non_blocking=True)
        self.draw()
        if self.segments is None:
            self.graph.replay()
        else:
            g1, g2a, g2b, g3 = self.segments
            main, comm = torch.cuda.current_stream(self.dev), self._comm_stream
            g1.replay()
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                self.grad_sync.finish_tensors(self._d_grads)     # overlaps the generator forward below
            g2a.replay()
            main.wait_stream(comm)
            g2b.replay()
            self.grad_sync.finish_tensors(self._g_grads)
            g3.replay()
        return self.out
