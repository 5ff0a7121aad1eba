"""CPU oracle for the ``network_dict.py`` variant of the 3D-PGAN train step.  TEST INFRASTRUCTURE ONLY.

Functional fp32 restatement of

    pgan_pytorch/network_dict.py (whole file)   -> generator_forward / discriminator_forward
    pgan_pytorch/loss.py:3-27                   -> gradient_penalty (shared formula, this file's D)
    pgan_pytorch/train.py:133-190               -> DictTrainState.step   (as written: network_dict's G returns
                                                   a tensor, so ``G(z, alpha).detach()`` of train.py:146 works)

calling the same torch CPU operators in the same order as the reference (torch is the reference's own
third-party arithmetic).  Pinned bit-for-bit against the unmodified reference modules by
``oracle/pin_dict_against_reference.py``, which mints ``tests/golden/dict_*.npz``.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.

Differences from network.py that matter for the arithmetic: the He gain of the equalized learning rate is
``calculate_gain(nonlinearity, param)`` (network_dict.py:31-38; 'linear' for ToRGB and the last linear), the
activation is ``LeakyReLU(0.3)`` -- the module constant LEAKINESS, whatever ``param`` -- or ReLU
(network_dict.py:18-23), there is no minibatch-stddev (network_dict.py:209-211), and the fade-in blend happens at
the top level only (network_dict.py:254-256, 379-388).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn.init import calculate_gain

Params = Dict[str, torch.Tensor]
LEAKINESS = 0.3     # network_dict.py:18


def _std(weight, nonlinearity, param):
    """network_dict.py:31-38."""
    return calculate_gain(nonlinearity, param) / np.sqrt(weight[0].numel())


def act(x, nonlinearity):
    """network_dict.py:19-23."""
    if nonlinearity == "relu":
        return F.relu(x)
    if nonlinearity == "leaky_relu":
        return F.leaky_relu(x, LEAKINESS)
    raise ValueError(f"Unsupported nonlinearity {nonlinearity}")


def eq_conv3d(x, p, name, padding, nonlinearity, param):
    w = p[name + ".weight"]
    return F.conv3d(x, w * _std(w, nonlinearity, param), p[name + ".bias"], 1, padding)


def eq_linear(x, p, name, nonlinearity, param):
    w = p[name + ".weight"]
    return F.linear(x, w * _std(w, nonlinearity, param), p[name + ".bias"])


def pixel_norm(x):
    """network_dict.py:271-272."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def generator_forward(p: Params, z, alpha, phase, nonlinearity, param, base_shape=(1, 4, 4)):
    """network_dict.py:372-390."""
    w0 = p["generator_in.0.weight"]
    base_dim = w0.shape[0] // int(np.prod(base_shape))
    x = act(eq_linear(z, p, "generator_in.0", nonlinearity, param), nonlinearity)
    x = torch.reshape(x, [-1, base_dim] + list(base_shape))
    x = pixel_norm(act(eq_conv3d(x, p, "generator_in.3", 1, nonlinearity, param), nonlinearity))
    x_up = None
    for i in range(2, phase + 1):
        if i == phase:
            x_up = F.interpolate(eq_conv3d(x, p, "torgb_prev.conv", 0, "linear", None), scale_factor=2, mode="nearest")
        b = f"blocks.block_phase_{i}"
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = pixel_norm(act(eq_conv3d(x, p, b + ".conv1", 1, nonlinearity, param), nonlinearity))
        x = act(pixel_norm(eq_conv3d(x, p, b + ".conv2", 1, nonlinearity, param)), nonlinearity)
    img = eq_conv3d(x, p, "torgb_current.conv", 0, "linear", None)
    if x_up is not None:
        img = alpha * x_up + (1 - alpha) * img
    return img


def discriminator_forward(p: Params, x_in, alpha, phase, nonlinearity, param):
    """network_dict.py:246-259."""
    x = act(eq_conv3d(x_in, p, "fromrgb_current.fromrgb.0", 0, nonlinearity, param), nonlinearity)
    for i in reversed(range(2, phase + 1)):
        b = f"blocks.block_phase_{i}"
        x = act(eq_conv3d(x, p, b + ".conv1", 1, nonlinearity, param), nonlinearity)
        x = act(eq_conv3d(x, p, b + ".conv2", 1, nonlinearity, param), nonlinearity)
        x = F.avg_pool3d(x, 2)
        if i == phase:
            prev = act(eq_conv3d(F.avg_pool3d(x_in, 2), p, "fromrgb_prev.fromrgb.0", 0, nonlinearity, param), nonlinearity)
            x = alpha * prev + (1 - alpha) * x
    x = act(eq_conv3d(x, p, "discriminator_out.0", 1, nonlinearity, param), nonlinearity)
    x = torch.flatten(x, 1)
    x = act(eq_linear(x, p, "discriminator_out.3", nonlinearity, param), nonlinearity)
    return eq_linear(x, p, "discriminator_out.5", "linear", None)


def gradient_penalty(pd: Params, real, fake, alpha, phase, nonlinearity, param, eps, weight=10.0):
    """loss.py:7-27 with the (B,1,1,1,1) draw of loss.py:11 passed in."""
    inter = (eps * real + ((1 - eps) * fake)).requires_grad_(True)
    d_inter = discriminator_forward(pd, inter, alpha, phase, nonlinearity, param)
    grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=torch.ones(real.shape[0], 1),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    flat = grads.view(grads.size(0), -1)
    return ((flat.norm(2, dim=1) - 1) ** 2).mean() * weight


class DictTrainState:
    """One iteration of train.py:133-190 with main.py:141-142's Adam(betas=(0, .99)) on network_dict parameters."""

    def __init__(self, pg: Params, pd: Params, phase, nonlinearity, param, lr=1e-3):
        self.pg = {k: v.detach().clone().requires_grad_(True) for k, v in pg.items()}
        self.pd = {k: v.detach().clone().requires_grad_(True) for k, v in pd.items()}
        self.phase, self.nonlinearity, self.param = phase, nonlinearity, param
        self.g_opt = torch.optim.Adam(list(self.pg.values()), lr=lr, betas=(0.0, 0.99))
        self.d_opt = torch.optim.Adam(list(self.pd.values()), lr=lr, betas=(0.0, 0.99))

    def _g(self, z, alpha):
        return generator_forward(self.pg, z, alpha, self.phase, self.nonlinearity, self.param)

    def _d(self, x, alpha):
        return discriminator_forward(self.pd, x, alpha, self.phase, self.nonlinearity, self.param)

    def step(self, x_real, noise, z_d, eps, z_g, alpha, apply=True):
        for v in self.pg.values():
            v.requires_grad_(False)
        for v in self.pd.values():
            v.requires_grad_(True)
        x = x_real + noise * 1e-2
        with torch.no_grad():
            x_fake = self._g(z_d, alpha)
        d_real, d_fake = self._d(x, alpha), self._d(x_fake, alpha)
        gp = gradient_penalty(self.pd, x, x_fake, alpha, self.phase, self.nonlinearity, self.param, eps)
        d_loss = -d_real.mean() + d_fake.mean() + gp + 1e-3 * (d_real ** 2).mean()
        self.d_opt.zero_grad()
        d_loss.backward()
        d_grads = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in self.pd.items()}
        if apply:
            self.d_opt.step()
        for v in self.pg.values():
            v.requires_grad_(True)
        for v in self.pd.values():
            v.requires_grad_(False)
        img = self._g(z_g, alpha)
        g_loss = -self._d(img, alpha).mean()
        self.g_opt.zero_grad()
        g_loss.backward()
        g_grads = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in self.pg.items()}
        if apply:
            self.g_opt.step()
        for v in list(self.pg.values()) + list(self.pd.values()):
            v.requires_grad_(True)
        return dict(d_loss=float(d_loss.detach()), gp=float(gp.detach()), g_loss=float(g_loss.detach()),
                    d_grads=d_grads, g_grads=g_grads, img=img.detach(), d_real=d_real.detach(), d_fake=d_fake.detach())
