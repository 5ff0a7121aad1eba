"""The train step of the reference's ``pgan_pytorch/train.py:126-198`` on the CUDA path.

``train_epoch`` keeps the reference's signature and return value.  ``train_step`` is one
iteration of its loop body (train.py:133-190); it takes optional pre-drawn random tensors so a
parity test can replay the oracle's exact draws, and returns the scalars as 0-dim device
tensors (the reference's five ``.item()`` host syncs per step, train.py:163-164,187-188, are
left to the caller).

Differences from the reference file as written: ``G(z, alpha)[-1]`` replaces
``G(z, alpha)`` because network.py's generator returns the list of images at every resolution
and ``.detach()`` on a list raises (train.py:146).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import kernels as K
from . import ops
from .loss import compute_gradient_penalty, wasserstein_loss

import os

# one-GPU step: run the G update's generator forward beside the D phase (A/B switch: SARAGAN_EARLY_G=0)
EARLY_G_FORWARD = os.environ.get("SARAGAN_EARLY_G", "1") != "0"

_side_streams: Dict[tuple, "torch.cuda.Stream"] = {}


def _side_stream(dev, index: int = 0) -> "torch.cuda.Stream":
    key = (torch.device(dev), index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key[0])
    return _side_streams[key]


def top_image(images) -> torch.Tensor:
    """network.py's generator returns the list of images at every resolution (network.py:274-284), network_dict.py's
    the image at the current one (network_dict.py:385-390): the step trains on the latter either way."""
    return images[-1] if isinstance(images, (list, tuple)) else images


def _set_requires_grad(module, flag: bool) -> None:
    for p in module.parameters():
        p.requires_grad = flag


def d_phase(x_real: torch.Tensor, generator, discriminator, discriminator_optim, alpha, *,
            noise: Optional[torch.Tensor] = None, z_d: Optional[torch.Tensor] = None,
            eps: Optional[torch.Tensor] = None, grad_sync=None, overlap_gp: bool = True) -> Dict[str, torch.Tensor]:
    """train.py:134-159: D forward passes, gradient penalty, d_loss.backward() (no optimiser step)."""
    dev = discriminator.device
    batch = x_real.shape[0]
    generator.eval()
    discriminator.train()
    _set_requires_grad(generator, False)
    _set_requires_grad(discriminator, True)

    # every stale weight packing of both networks in one launch each (the optimiser steps of the previous iteration
    # bumped the parameters' versions)
    ops.prepack(generator)
    ops.prepack(discriminator)
    x_real = x_real.to(dev, non_blocking=True).float().contiguous()
    if noise is None:
        noise = torch.randn_like(x_real)
    x_real = K.lincomb(x_real, noise.to(dev).float().contiguous(), 1.0, 1e-2)
    if z_d is None:
        z_d = torch.randn(batch, generator.latent_dim)
    with torch.no_grad():
        x_fake = top_image(generator(z_d, alpha)).detach()

    # The gradient-penalty chain (D(interpolates), its input gradient and the double backward) shares
    # nothing with the D(real)/D(fake) chain but the weights: on a GPU the two run on two streams and meet at
    # the gradient sum.  The low-resolution levels of either chain occupy a fraction of the SMs, which the
    # other chain's kernels fill.  (Hook-based gradient sync fires per accumulated gradient, so it keeps
    # the single-stream order.)
    # (A layer's FIRST packing happens lazily at its first use and is cached: on a fresh network -- first step,
    # after grow() -- that would be on whichever stream gets there first while the other stream reads the same
    # buffer un-ordered, so that one pass stays on a single stream.)
    hooked = grad_sync is not None and not hasattr(grad_sync, "reduce")      # hook-driven bucketing needs one stream
    two_streams = overlap_gp and x_real.is_cuda and not hooked and ops.packs_settled(discriminator)
    if two_streams:
        d_params = [p for p in discriminator.parameters() if p.requires_grad]
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            gp_loss = compute_gradient_penalty(discriminator, x_real, x_fake, alpha, random_uniform=eps)
            gp_grads = torch.autograd.grad(gp_loss, d_params, allow_unused=True)
    # D(real) and D(fake) as ONE batch-2B pass (minibatch-stddev keeps the two minibatches apart, so
    # the values equal the reference's two calls, train.py:148-149): half the launches and twice the
    # tiles per launch at the low-resolution levels
    d_both = discriminator(torch.cat([x_real, x_fake]), alpha, sub_batches=2)
    d_real, d_fake = d_both[:batch], d_both[batch:]
    real_loss = wasserstein_loss(d_real)
    fake_loss = wasserstein_loss(d_fake)
    drift_loss = 1e-3 * (d_real ** 2).mean()
    discriminator_optim.zero_grad()
    if two_streams:
        d_loss = -real_loss + fake_loss + drift_loss
        d_loss.backward()
        main.wait_stream(side)
        pairs = [(p.grad, g) for p, g in zip(d_params, gp_grads) if g is not None and p.grad is not None]
        torch._foreach_add_([a for a, _ in pairs], [b for _, b in pairs])
        for p, g in zip(d_params, gp_grads):
            if g is not None and p.grad is None:
                p.grad = g
        d_loss = d_loss.detach() + gp_loss.detach()
    else:
        gp_loss = compute_gradient_penalty(discriminator, x_real, x_fake, alpha, random_uniform=eps)
        d_loss = -real_loss + fake_loss + gp_loss + drift_loss
        if grad_sync is not None:
            grad_sync.arm(discriminator)
        d_loss.backward()
    ops.mark_packs_settled(discriminator)
    return {"d_loss": d_loss.detach(), "gp": gp_loss.detach(), "d_real_mean": real_loss.detach(),
            "x_real": x_real}


def g_phase(batch: int, generator, discriminator, generator_optim, alpha, *,
            z_g: Optional[torch.Tensor] = None, grad_sync=None, x_fake: Optional[torch.Tensor] = None
            ) -> Dict[str, torch.Tensor]:
    """train.py:166-184: G forward, D forward on the fakes, g_loss.backward() (no optimiser step).
    x_fake: G(z_g, alpha)[-1] with its autograd graph, when the caller ran the generator forward ahead of
    time (graph.GraphedTrainStep overlaps it with the D gradients' all-reduce)."""
    generator.train()
    discriminator.eval()
    _set_requires_grad(generator, True)
    _set_requires_grad(discriminator, False)
    ops.prepack(discriminator)                # the D update just changed its weights
    if x_fake is None:
        if z_g is None:
            z_g = torch.randn(batch, generator.latent_dim)
        x_fake = top_image(generator(z_g, alpha))
    d_fake = discriminator(x_fake, alpha)
    g_loss = -wasserstein_loss(d_fake)
    generator_optim.zero_grad()
    if grad_sync is not None:
        grad_sync.arm(generator)
    g_loss.backward()
    _set_requires_grad(generator, True)
    _set_requires_grad(discriminator, True)
    ops.mark_packs_settled(generator)
    return {"g_loss": g_loss.detach(), "d_fake_mean": d_fake.detach().mean(), "x_fake": x_fake.detach()}


def train_step(x_real: torch.Tensor, generator, discriminator, generator_optim, discriminator_optim,
               alpha, *, noise: Optional[torch.Tensor] = None, z_d: Optional[torch.Tensor] = None,
               z_g: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None,
               apply: bool = True, grad_sync=None, overlap_gp: bool = True) -> Dict[str, torch.Tensor]:
    """One D update followed by one G update (train.py:133-190).

    grad_sync: optional ``comm.DataParallel``; ``arm(module)`` is called before ``backward()``
    and ``finish(module)`` before ``optim.step()`` (the data-parallel gradient all-reduce,
    reference: hvd.DistributedOptimizer, main.py:153-160)."""
    # The generator forward of the G update needs nothing the D update changes (the generator's weights move at the END of
    # the step), so on one GPU it runs on its own stream beside the whole D phase: its low-resolution levels -- a chain
    # of small kernels -- fill SMs and launch slots the D chains leave idle.  Same arithmetic, same gradients: autograd
    # runs each backward node on its forward's stream and joins the streams itself.  (Needs every weight packing of both
    # networks to exist already, see ops.packs_settled; with a gradient exchange the arena path below places the same
    # forward beside the D all-reduce instead.)
    dev = discriminator.device
    x_fake_g = None
    early = (grad_sync is None and overlap_gp and dev.type == "cuda" and EARLY_G_FORWARD
             and ops.packs_settled(discriminator) and ops.packs_settled(generator))
    if early:
        if z_d is None:                              # keep train.py's draw order on the host generator: z (D), then z (G)
            z_d = torch.randn(x_real.shape[0], generator.latent_dim)
        if z_g is None:
            z_g = torch.randn(x_real.shape[0], generator.latent_dim)
        ops.prepack(generator)                       # on the main stream, before the fork
        main, side_g = torch.cuda.current_stream(dev), _side_stream(dev, 1)
        side_g.wait_stream(main)
        generator.train()
        _set_requires_grad(generator, True)
        with torch.cuda.stream(side_g):
            x_fake_g = top_image(generator(z_g, alpha))
    out = d_phase(x_real, generator, discriminator, discriminator_optim, alpha, noise=noise, z_d=z_d, eps=eps,
                  grad_sync=grad_sync, overlap_gp=overlap_gp)
    if early:
        torch.cuda.current_stream(dev).wait_stream(side_g)
    if hasattr(grad_sync, "reduce") and hasattr(discriminator_optim, "ema_state"):
        # gradient arena + fused Adam (comm.ArenaAllReduce): the D gradients are averaged on a second stream WHILE the
        # generator forward of the G update runs (it needs nothing the D update changes); Adam reads the arena
        dev = discriminator.device
        if z_g is None:
            z_g = torch.randn(x_real.shape[0], generator.latent_dim)
        if dev.type == "cuda":
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                d_grads = grad_sync.reduce(discriminator)
        else:
            d_grads = grad_sync.reduce(discriminator)
        generator.train()
        _set_requires_grad(generator, True)
        x_fake_g = top_image(generator(z_g, alpha))
        if dev.type == "cuda":
            main.wait_stream(side)
        if apply:
            discriminator_optim.step(grads=d_grads)
        g = g_phase(x_real.shape[0], generator, discriminator, generator_optim, alpha, z_g=z_g, x_fake=x_fake_g)
        g_grads = grad_sync.reduce(generator)
        if apply:
            generator_optim.step(grads=g_grads)
    else:
        if grad_sync is not None:
            grad_sync.finish(discriminator)
        if apply:
            discriminator_optim.step()
        g = g_phase(x_real.shape[0], generator, discriminator, generator_optim, alpha, z_g=z_g, grad_sync=grad_sync,
                    x_fake=x_fake_g)
        if grad_sync is not None:
            grad_sync.finish(generator)
        if apply:
            generator_optim.step()
    out["g_loss"] = g["g_loss"]
    out["distance"] = out["d_real_mean"] - g["d_fake_mean"]
    out["x_fake"] = g["x_fake"]
    return out


def train_epoch(data_loader, generator, discriminator, generator_optim, discriminator_optim, alpha):
    """train.py:126-198, same signature and return tuple."""
    d_losses, g_losses, distances, gradient_penalties = [], [], [], []
    out = None
    for x_real in data_loader:
        out = train_step(x_real, generator, discriminator, generator_optim, discriminator_optim, alpha)
        d_losses.append(out["d_loss"])
        g_losses.append(out["g_loss"])
        distances.append(out["distance"])
        gradient_penalties.append(out["gp"])
    if out is None:
        raise ValueError("empty data_loader")
    # one host sync per epoch instead of five per step
    stats = torch.stack([torch.stack(v).mean() for v in (d_losses, g_losses, distances,
                                                         gradient_penalties)]).cpu().numpy()
    return (out["x_fake"].cpu(), out["x_real"].cpu(), np.float64(stats[0]), np.float64(stats[1]),
            np.float64(stats[2]), np.float64(stats[3]))


def make_optimizers(generator, discriminator, lr: float = 1e-3, world_size: int = 1):
    """main.py:138-145: Adam(lr * sqrt(world), betas=(0, 0.99)) for both nets."""
    lr = lr * float(np.sqrt(world_size))
    d_optim = torch.optim.Adam(discriminator.parameters(), lr=lr, betas=(0.0, 0.99))
    g_optim = torch.optim.Adam(generator.parameters(), lr=lr, betas=(0.0, 0.99))
    return g_optim, d_optim
