"""CPU oracle for the saraGAN 3D-PGAN train step.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 restatement of the reference's hot path:

    pgan_pytorch/network.py   (whole file)        -> generator_forward / discriminator_forward
    pgan_pytorch/loss.py:3-27                     -> wasserstein / gradient_penalty
    pgan_pytorch/train.py:133-190 (step body)     -> d_step_loss / g_step_loss / TrainState.step
    pgan_pytorch/main.py:141-142 (Adam b=(0,.99)) -> TrainState

It is the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; nothing under ``saragan_b200/`` does.

The arithmetic itself (conv3d, linear, pooling, autograd) is third-party in the
reference too: it calls ``torch`` (version unpinned by the reference; 2.11.0 here), so
the restatement calls the same torch CPU operators in the same order.  The reference
publishes no golden vectors; parity is pinned by ``oracle/pin_against_reference.py``
which imports the unmodified reference modules from /root/reference (in the build
container only), checks this file against them bit-for-bit, and mints
``tests/golden/*.npz``.

Parameters travel as a plain ``dict`` keyed by the reference's ``state_dict`` names
(e.g. ``blocks.0.conv1.weight``), so the same dict drives the reference modules, this
oracle and the CUDA modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- helpers
def num_filters(phase: int, num_phases: int, base_dim: int):
    """network.py:10-13.  Channel count at growth level ``phase`` (1 = 4x4 base)."""
    shift = phase - num_phases + int(np.log2(base_dim / 16))
    return min(base_dim // (2 ** shift), base_dim)


def _eq_std(weight: torch.Tensor):
    """network.py:16-23: gain('conv3d'|'linear') == 1, std = 1/sqrt(fan_in) (float64)."""
    fan_in = weight[0].numel()
    return 1.0 / np.sqrt(fan_in)


def eq_conv3d(x, weight, bias, padding):
    """network.py:54-56."""
    return F.conv3d(x, weight * _eq_std(weight), bias, 1, padding)


def eq_linear(x, weight, bias):
    """network.py:76-77."""
    return F.linear(x, weight * _eq_std(weight), bias)


def lrelu(x):
    return F.leaky_relu(x, 0.2)


def pixel_norm(x):
    """network.py:196-197."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def mbstd_group(batch: int, group_size: int = 4) -> int:
    """network.py:119-124: min(4, B), bumped to the next divisor of B."""
    g = min(group_size, batch)
    if g < batch:
        for i in range(g, batch + 1):
            if batch % i == 0:
                g = i
                break
    return g


def minibatch_stddev(x):
    """network.py:118-133.  The reference centres its *input in place* (`y -= mean` on a
    view), so the C feature channels that get concatenated are the group-centred ones;
    restated here out of place with the same arithmetic."""
    b, c, d, h, w = x.shape
    g = mbstd_group(b)
    y = x.view(g, -1, c, d, h, w)
    yc = y - torch.mean(y, dim=0, keepdim=True)
    s = torch.sqrt(torch.mean(yc ** 2, dim=0) + 1e-8)
    t = torch.mean(s, dim=[1, 2, 3, 4], keepdim=True)
    t = t.repeat([g, 1, d, h, w])
    return torch.cat([yc.reshape(b, c, d, h, w), t], dim=1)


def up2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def pool2(x):
    return F.avg_pool3d(x, 2)


# --------------------------------------------------------------------------- generator
def generator_forward(p: Params, z, alpha, phase: int, base_shape=(1, 4, 4),
                      taps: Optional[dict] = None) -> List[torch.Tensor]:
    """network.py:270-284.  Returns the list of `phase` images (one per resolution)."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    w0 = p["generator_in.0.weight"]
    base_dim = w0.shape[0] // int(np.prod(base_shape))
    x = lrelu(eq_linear(z, w0, p["generator_in.0.bias"]))
    x = torch.reshape(x, [-1, base_dim] + list(base_shape))
    x = eq_conv3d(x, p["generator_in.3.weight"], p["generator_in.3.bias"], 1)
    x = rec("g.in", pixel_norm(lrelu(x)))
    img = eq_conv3d(x, p["to_rgbs.0.conv.weight"], p["to_rgbs.0.conv.bias"], 0)
    outs = [img]
    for i in range(phase - 1):
        x = up2(x)
        x = eq_conv3d(x, p[f"blocks.{i}.conv1.weight"], p[f"blocks.{i}.conv1.bias"], 1)
        x = rec(f"g.b{i}.c1", pixel_norm(lrelu(x)))
        x = eq_conv3d(x, p[f"blocks.{i}.conv2.weight"], p[f"blocks.{i}.conv2.bias"], 1)
        x = rec(f"g.b{i}.c2", lrelu(pixel_norm(x)))
        gen = eq_conv3d(x, p[f"to_rgbs.{i + 1}.conv.weight"], p[f"to_rgbs.{i + 1}.conv.bias"], 0)
        img = alpha * up2(img) + (1 - alpha) * gen
        outs.append(img)
    return outs


# ----------------------------------------------------------------------- discriminator
def discriminator_forward(p: Params, x_in, alpha, phase: int, num_phases: int,
                          taps: Optional[dict] = None) -> torch.Tensor:
    """network.py:169-189 (without the five debug prints).  ModuleList indices:
    `fromrgbs` has num_phases entries, `blocks` has num_phases-1; the reference indexes
    both from the end (`[-phase]`, `[-i]`)."""
    def rec(name, t):
        if taps is not None:
            taps[name] = t
        return t

    n_rgb, n_blk = num_phases, num_phases - 1

    def fromrgb(j, img):
        return lrelu(eq_conv3d(img, p[f"fromrgbs.{j}.fromrgb.0.weight"],
                               p[f"fromrgbs.{j}.fromrgb.0.bias"], 0))

    img = x_in
    x = rec("d.rgb", fromrgb(n_rgb - phase, x_in))
    for i in reversed(range(1, phase)):
        j = n_blk - i
        x = lrelu(eq_conv3d(x, p[f"blocks.{j}.conv1.weight"], p[f"blocks.{j}.conv1.bias"], 1))
        rec(f"d.b{i}.c1", x)
        x = lrelu(eq_conv3d(x, p[f"blocks.{j}.conv2.weight"], p[f"blocks.{j}.conv2.bias"], 1))
        rec(f"d.b{i}.c2", x)
        x = pool2(x)
        img = pool2(img)
        prev = fromrgb(n_rgb - i, img)
        x = rec(f"d.b{i}.out", alpha * prev + (1 - alpha) * x)
    x = minibatch_stddev(x)
    x = lrelu(eq_conv3d(x, p["discriminator_out.1.weight"], p["discriminator_out.1.bias"], 1))
    rec("d.out.conv", x)
    x = torch.flatten(x, 1)
    x = lrelu(eq_linear(x, p["discriminator_out.4.weight"], p["discriminator_out.4.bias"]))
    x = eq_linear(x, p["discriminator_out.6.weight"], p["discriminator_out.6.bias"])
    return x


# ------------------------------------------------------------------------------ losses
def gradient_penalty(pd: Params, real, fake, alpha, phase, num_phases, eps, weight=10.0,
                     return_grad=False):
    """loss.py:7-27.  `eps` is the (B,1,1,1,1) uniform draw of loss.py:11, passed in so
    both arms of a parity test see the same numbers."""
    inter = (eps * real + ((1 - eps) * fake)).requires_grad_(True)
    d_inter = discriminator_forward(pd, inter, alpha, phase, num_phases)
    ones = torch.ones(real.shape[0], 1)
    grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=ones,
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    flat = grads.view(grads.size(0), -1)
    gp = ((flat.norm(2, dim=1) - 1) ** 2).mean() * weight
    return (gp, grads) if return_grad else gp


def d_step_loss(pg: Params, pd: Params, x_real_noisy, z, eps, alpha, phase, num_phases):
    """train.py:143-157 given the already-noised reals.  Returns dict of scalars/tensors."""
    with torch.no_grad():
        x_fake = generator_forward(pg, z, alpha, phase)[-1]
    x_fake = x_fake.detach()
    d_real = discriminator_forward(pd, x_real_noisy, alpha, phase, num_phases)
    d_fake = discriminator_forward(pd, x_fake, alpha, phase, num_phases)
    gp, gp_grad = gradient_penalty(pd, x_real_noisy, x_fake, alpha, phase, num_phases, eps,
                                   return_grad=True)
    drift = 1e-3 * (d_real ** 2).mean()
    d_loss = -d_real.mean() + d_fake.mean() + gp + drift
    return dict(d_loss=d_loss, gp=gp, d_real=d_real, d_fake=d_fake, x_fake=x_fake,
                gp_grad=gp_grad)


def g_step_loss(pg: Params, pd: Params, z, alpha, phase, num_phases):
    """train.py:178-181."""
    x_fake = generator_forward(pg, z, alpha, phase)[-1]
    d_fake = discriminator_forward(pd, x_fake, alpha, phase, num_phases)
    return dict(g_loss=-d_fake.mean(), d_fake=d_fake, x_fake=x_fake)


# ------------------------------------------------------------------------- param setup
def active_names(kind: str, phase: int, num_phases: int) -> List[str]:
    """Names of the parameters that receive a gradient at `phase` (network.py builds all
    num_phases levels up front; inactive ones keep grad None)."""
    names: List[str] = []
    if kind == "g":
        names += [f"generator_in.{i}.{s}" for i in (0, 3) for s in ("weight", "bias")]
        names += [f"to_rgbs.0.conv.{s}" for s in ("weight", "bias")] if phase == 1 else []
        for i in range(phase - 1):
            for c in ("conv1", "conv2"):
                names += [f"blocks.{i}.{c}.{s}" for s in ("weight", "bias")]
        # with fade-in every to_rgb up to the current level feeds the last image
        if phase > 1:
            for i in range(phase):
                names += [f"to_rgbs.{i}.conv.{s}" for s in ("weight", "bias")]
    else:
        n_rgb, n_blk = num_phases, num_phases - 1
        for i in range(1, phase + 1):
            names += [f"fromrgbs.{n_rgb - i}.fromrgb.0.{s}" for s in ("weight", "bias")]
        for i in range(1, phase):
            for c in ("conv1", "conv2"):
                names += [f"blocks.{n_blk - i}.{c}.{s}" for s in ("weight", "bias")]
        for i in (1, 4, 6):
            names += [f"discriminator_out.{i}.{s}" for s in ("weight", "bias")]
    return names


def init_params(kind: str, num_phases: int, base_dim: int, latent_dim: int,
                base_shape=(1, 1, 4, 4), generator: Optional[torch.Generator] = None) -> Params:
    """Fresh parameters with the reference's distributions (weight ~ N(0,1),
    bias ~ U(+-1/sqrt(fan_in)), network.py:16-23,46-52,67-74) and its state_dict names and
    shapes.  (Not the reference's RNG stream: fixtures that need the reference's exact
    draw are minted from the reference modules by pin_against_reference.py.)"""
    ch = base_shape[0]
    vol = int(np.prod(base_shape[1:]))
    shapes: Dict[str, tuple] = {}
    if kind == "g":
        shapes["generator_in.0"] = (vol * base_dim, latent_dim)
        shapes["generator_in.3"] = (base_dim, base_dim, 3, 3, 3)
        shapes["to_rgbs.0.conv"] = (ch, base_dim, 1, 1, 1)
        for k, i in enumerate(range(2, num_phases + 1)):
            fi, fo = num_filters(i, num_phases, base_dim), num_filters(i + 1, num_phases, base_dim)
            shapes[f"blocks.{k}.conv1"] = (fo, fi, 3, 3, 3)
            shapes[f"blocks.{k}.conv2"] = (fo, fo, 3, 3, 3)
            shapes[f"to_rgbs.{k + 1}.conv"] = (ch, fo, 1, 1, 1)
    else:
        fo = base_dim
        for k, i in enumerate(reversed(range(2, num_phases + 1))):
            fi, fo = num_filters(i, num_phases, base_dim), num_filters(i - 1, num_phases, base_dim)
            shapes[f"blocks.{k}.conv1"] = (fi, fi, 3, 3, 3)
            shapes[f"blocks.{k}.conv2"] = (fo, fi, 3, 3, 3)
            shapes[f"fromrgbs.{k}.fromrgb.0"] = (fi, ch, 1, 1, 1)
        shapes[f"fromrgbs.{num_phases - 1}.fromrgb.0"] = (base_dim, ch, 1, 1, 1)
        shapes["discriminator_out.1"] = (base_dim, fo + 1, 3, 3, 3)
        shapes["discriminator_out.4"] = (latent_dim, vol * base_dim)
        shapes["discriminator_out.6"] = (1, latent_dim)
    out: Params = {}
    for name, shp in shapes.items():
        shp = tuple(int(s) for s in shp)
        w = torch.empty(shp).normal_(0, 1, generator=generator)
        bound = 1.0 / math.sqrt(w[0].numel())
        b = torch.empty(shp[0]).uniform_(-bound, bound, generator=generator)
        out[name + ".weight"], out[name + ".bias"] = w, b
    return out


# ------------------------------------------------------------------------ training step
class TrainState:
    """Restatement of one iteration of train.py:133-190 with main.py:141-142's optimisers
    (Adam, lr 1e-3 * sqrt(world), betas (0, 0.99)); `G(z, alpha)[-1]` replaces the
    as-written `G(z, alpha)` because network.py's generator returns a list."""

    def __init__(self, pg: Params, pd: Params, phase: int, num_phases: int, lr=1e-3):
        self.pg = {k: v.detach().clone().requires_grad_(True) for k, v in pg.items()}
        self.pd = {k: v.detach().clone().requires_grad_(True) for k, v in pd.items()}
        self.phase, self.num_phases = phase, num_phases
        self.g_opt = torch.optim.Adam(list(self.pg.values()), lr=lr, betas=(0.0, 0.99))
        self.d_opt = torch.optim.Adam(list(self.pd.values()), lr=lr, betas=(0.0, 0.99))

    def step(self, x_real, noise, z_d, eps, z_g, alpha, apply=True):
        """x_real (B,1,D,H,W); noise = randn_like(x_real) (train.py:144 multiplies by 1e-2);
        z_d/z_g the two latent draws (train.py:145,178); eps the GP draw (loss.py:11)."""
        for v in self.pg.values():
            v.requires_grad_(False)
        for v in self.pd.values():
            v.requires_grad_(True)
        x = x_real + noise * 1e-2
        d = d_step_loss(self.pg, self.pd, x, z_d, eps, alpha, self.phase, self.num_phases)
        self.d_opt.zero_grad()
        d["d_loss"].backward()
        d_grads = {k: (None if v.grad is None else v.grad.detach().clone())
                   for k, v in self.pd.items()}
        if apply:
            self.d_opt.step()
        for v in self.pg.values():
            v.requires_grad_(True)
        for v in self.pd.values():
            v.requires_grad_(False)
        g = g_step_loss(self.pg, self.pd, z_g, alpha, self.phase, self.num_phases)
        self.g_opt.zero_grad()
        g["g_loss"].backward()
        g_grads = {k: (None if v.grad is None else v.grad.detach().clone())
                   for k, v in self.pg.items()}
        if apply:
            self.g_opt.step()
        for v in list(self.pg.values()) + list(self.pd.values()):
            v.requires_grad_(True)
        return dict(d_loss=float(d["d_loss"].detach()), gp=float(d["gp"].detach()), g_loss=float(g["g_loss"].detach()),
                    distance=float(d["d_real"].detach().mean()) - float(g["d_fake"].detach().mean()),
                    d_grads=d_grads, g_grads=g_grads, gp_grad=d["gp_grad"].detach(),
                    x_fake=g["x_fake"].detach(), d_real=d["d_real"].detach())


# --------------------------------------------------------------------------- cost model
def conv_flops_per_image(kind: str, phase: int, num_phases: int, base_dim: int, latent_dim: int,
                         base_shape=(1, 1, 4, 4)) -> float:
    """Forward conv+linear FLOPs per image (2*Cin*Cout*k^3*voxels), SURVEY.md 8(d)."""
    vol0 = int(np.prod(base_shape[1:]))
    f = lambda i: num_filters(i, num_phases, base_dim)
    total = 0.0
    if kind == "g":
        total += 2 * latent_dim * vol0 * base_dim
        total += 2 * base_dim * base_dim * 27 * vol0
        total += 2 * base_dim * vol0
        v = vol0
        for i in range(2, phase + 1):
            v *= 8
            fi, fo = f(i), f(i + 1)
            total += 2 * fi * fo * 27 * v + 2 * fo * fo * 27 * v + 2 * fo * v
    else:
        v = vol0 * 8 ** (phase - 1)
        total += 2 * f(phase) * v
        for i in range(phase, 1, -1):
            fi, fo = f(i), f(i - 1)
            total += 2 * fi * fi * 27 * v + 2 * fi * fo * 27 * v
            v //= 8
            total += 2 * fo * v
        total += 2 * (base_dim + 1) * base_dim * 27 * vol0
        total += 2 * vol0 * base_dim * latent_dim + 2 * latent_dim
    return float(total)


def step_flops_per_image(phase, num_phases, base_dim, latent_dim, base_shape=(1, 1, 4, 4)):
    """4*Gf + 14*Df (SURVEY.md 8(d) / BASELINE.md 3)."""
    gf = conv_flops_per_image("g", phase, num_phases, base_dim, latent_dim, base_shape)
    df = conv_flops_per_image("d", phase, num_phases, base_dim, latent_dim, base_shape)
    return 4 * gf + 14 * df
