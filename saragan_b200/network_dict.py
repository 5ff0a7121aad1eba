"""Drop-in for the reference's ``pgan_pytorch/network_dict.py`` -- the variant ``main.py:11`` actually imports
(SURVEY.md 8f row 3): ``grow()`` API, ``nn.ModuleDict`` blocks keyed ``block_phase_{i}``, one
``fromrgb_current`` / ``fromrgb_prev`` (``torgb_current`` / ``torgb_prev``) pair instead of a list per level,
configurable nonlinearity, no minibatch-stddev (commented out at network_dict.py:209-211), fade-in at the top
level only (network_dict.py:254-256, 379-388), and a generator that returns ONE tensor (network_dict.py:385-390).

Same classes, constructor arguments, attribute / sub-module / parameter names, ``state_dict`` layout and RNG
consumption at construction as the reference file; every operator runs in libsaragan_b200.so through the same
autograd Functions as ``saragan_b200.network``.

Nonlinearities (network_dict.py:18-23): ``'leaky_relu'`` is ``nn.LeakyReLU(LEAKINESS)`` with the MODULE
constant 0.3 whatever ``param`` says -- ``param`` only enters the He gain ``calculate_gain('leaky_relu', param)``
of the equalized learning rate (network_dict.py:31-38); ``'relu'`` is slope 0.  ``'swish'`` cannot be constructed
in the reference either (``calculate_gain('swish')`` raises ValueError) and raises the same error here.
The slope is a process-wide constant of the kernel library (``sg_set_leaky_slope``), like ``LEAKINESS`` is a
module constant of the reference: building a network here sets it, and ``forward`` checks it.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn.init import calculate_gain

from . import config, kernels, ops
from . import network as _net
from .network import ChannelNormalization, Reshape, _voxels, num_filters  # noqa: F401

LEAKINESS = 0.3                                         # network_dict.py:18
_SLOPES = {"relu": 0.0, "leaky_relu": LEAKINESS}        # network_dict.py:19-23 ('swish': see module docstring)


def _gain(nonlinearity, param):
    """network_dict.py:31-34: gain of torch's calculate_gain; raises ValueError for 'swish' as the reference does."""
    return calculate_gain(nonlinearity, param)


def _slope(nonlinearity) -> float:
    if nonlinearity not in _SLOPES:
        raise ValueError(f"Unsupported nonlinearity {nonlinearity}")
    return _SLOPES[nonlinearity]


def _activate_slope(nonlinearity) -> None:
    kernels.ensure_leaky_slope(_slope(nonlinearity))


def activation(nonlinearity):
    """network_dict.py:124-125 (the module objects only document the architecture; the activation itself is
    fused into the producing kernel's epilogue)."""
    return nn.ReLU() if nonlinearity == "relu" else nn.LeakyReLU(_slope(nonlinearity))


class EqualizedConv3d(_net.EqualizedConv3d):
    """network_dict.py:41-72."""

    def __init__(self, in_channels, out_channels, kernel_size, nonlinearity, stride=1, padding=0, param=None):
        if nonlinearity == "leaky_relu":
            assert param is not None
        self.nonlinearity, self.param = nonlinearity, param
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding)

    def _gain(self):
        return _gain(self.nonlinearity, self.param)


class EqualizedLinear(_net.EqualizedLinear):
    """network_dict.py:75-99."""

    def __init__(self, in_features, out_features, nonlinearity, param=None):
        if nonlinearity == "leaky_relu":
            assert param is not None
        self.nonlinearity, self.param = nonlinearity, param
        super().__init__(in_features, out_features)

    def _gain(self):
        return _gain(self.nonlinearity, self.param)


class DiscriminatorBlock(_net.DiscriminatorBlock):
    """network_dict.py:102-121."""

    def __init__(self, filters_in, filters_out, nonlinearity, param=None):
        nn.Sequential.__init__(self)
        self.filters_in, self.filters_out = filters_in, filters_out
        self.conv1 = EqualizedConv3d(filters_in, filters_in, 3, padding=1, nonlinearity=nonlinearity, param=param)
        self.conv2 = EqualizedConv3d(filters_in, filters_out, 3, padding=1, nonlinearity=nonlinearity, param=param)
        self.act = activation(nonlinearity)
        self.downsampling = nn.AvgPool3d(2)


class FromRGB(_net.FromRGB):
    """network_dict.py:128-141."""

    def __init__(self, channels_in, filters, nonlinearity, param=None):
        nn.Sequential.__init__(self)
        self.fromrgb = nn.Sequential(
            EqualizedConv3d(in_channels=channels_in, out_channels=filters, kernel_size=1, nonlinearity=nonlinearity,
                            param=param),
            activation(nonlinearity),
        )


class GeneratorBlock(_net.GeneratorBlock):
    """network_dict.py:276-296."""

    def __init__(self, filters_in, filters_out, nonlinearity, param=None):
        nn.Sequential.__init__(self)
        self.upsampling = nn.Upsample(scale_factor=2)
        self.conv1 = EqualizedConv3d(filters_in, filters_out, 3, nonlinearity, padding=1, param=param)
        self.conv2 = EqualizedConv3d(filters_out, filters_out, 3, nonlinearity, padding=1, param=param)
        self.act = activation(nonlinearity)
        self.cn = ChannelNormalization()


class ToRGB(_net.ToRGB):
    """network_dict.py:299-305."""

    def __init__(self, filters_in, channels):
        nn.Sequential.__init__(self)
        self.conv = EqualizedConv3d(filters_in, channels, 1, nonlinearity="linear")


class Discriminator(nn.Module):
    """network_dict.py:176-259."""

    def __init__(self, phase, num_phases, base_dim, latent_dim, base_shape, nonlinearity, param=None):
        super().__init__()
        self.channels = base_shape[0]
        self.base_shape = base_shape[1:]
        self.phase = phase
        self.num_phases = num_phases
        self.base_dim = base_dim
        self.nonlinearity = nonlinearity
        if nonlinearity == "leaky_relu":
            assert param is not None
        self.param = param
        if self.channels != 1:
            raise NotImplementedError("saragan_b200 covers single-channel volumes (CT), as the reference's data does")
        _slope(nonlinearity)

        filters_in = num_filters(phase, num_phases, base_dim)
        filters_out = num_filters(phase - 1, num_phases, base_dim)
        self.fromrgb_current = FromRGB(self.channels, filters_in, nonlinearity, param=param)
        self.fromrgb_prev = FromRGB(self.channels, filters_out, nonlinearity, param=param) if self.phase > 1 else None
        self.blocks = nn.ModuleDict()
        for i in range(2, phase + 1):
            filters_in = num_filters(i, num_phases, base_dim)
            filters_out = num_filters(i - 1, num_phases, base_dim)
            self.blocks[f"block_phase_{i}"] = DiscriminatorBlock(filters_in, filters_out, nonlinearity, param)
        self.downscale = nn.AvgPool3d(2)
        self.discriminator_out = nn.Sequential(
            EqualizedConv3d(base_dim, base_dim, 3, padding=1, nonlinearity=nonlinearity, param=param),
            activation(self.nonlinearity),
            nn.Flatten(),
            EqualizedLinear(int(np.prod(self.base_shape)) * base_dim, latent_dim, nonlinearity=nonlinearity,
                            param=param),
            activation(self.nonlinearity),
            EqualizedLinear(latent_dim, 1, nonlinearity="linear"),
        )
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(self.device)
        _activate_slope(nonlinearity)

    def grow(self):
        """network_dict.py:229-243."""
        self.phase += 1
        filters_in = num_filters(self.phase, self.num_phases, self.base_dim)
        filters_out = num_filters(self.phase - 1, self.num_phases, self.base_dim)
        self.blocks[f"block_phase_{self.phase}"] = DiscriminatorBlock(filters_in, filters_out, self.nonlinearity,
                                                                      param=self.param)
        self.fromrgb_prev = self.fromrgb_current
        self.fromrgb_current = FromRGB(self.channels, filters_in, self.nonlinearity, param=self.param)
        self.to(self.device)

    def forward(self, input, alpha, sub_batches: int = 1):
        """network_dict.py:246-259.  ``sub_batches`` is accepted for the train step's stacked D(real, fake) pass;
        without minibatch-stddev the samples never interact, so it changes nothing."""
        _activate_slope(self.nonlinearity)
        img = input.to(self.device).float().contiguous()
        alpha, beta = ops.blend_coef(alpha, img.device)
        # at phase > 1 the top FromRGB feeds only the first block's conv1, whose dgrad epilogue applies its mask
        x = self.fromrgb_current(img, premasked=self.phase > 1)
        for i in reversed(range(2, self.phase + 1)):
            x = self.blocks[f"block_phase_{i}"](x, input_is_lrelu=(i == self.phase))
            if i == self.phase:
                prev = self.fromrgb_prev(ops.Down2.apply(img, 0.125))
                x = ops.Lincomb.apply(prev, x, alpha, beta)
        out = self.discriminator_out
        x = out[0](x, lrelu=True)
        x = torch.flatten(ops.ToPlain.apply(x, out[0].out_channels), 1)
        x = out[3](x, lrelu=True)
        return out[5](x)


class Generator(nn.Module):
    """network_dict.py:318-390.  ``forward`` returns the image at the current resolution (one tensor)."""

    def __init__(self, phase, num_phases, base_dim, latent_dim, base_shape, nonlinearity, param=None):
        super().__init__()
        self.channels = base_shape[0]
        self.base_shape = base_shape[1:]
        self.phase = phase
        self.latent_dim = latent_dim
        self.base_dim = base_dim
        self.num_phases = num_phases
        self.nonlinearity = nonlinearity
        self.param = param
        if nonlinearity == "leaky_relu":
            assert param is not None
        if self.channels != 1:
            raise NotImplementedError("saragan_b200 covers single-channel volumes (CT), as the reference's data does")
        _slope(nonlinearity)

        self.generator_in = nn.Sequential(
            EqualizedLinear(latent_dim, int(np.prod(self.base_shape)) * base_dim, nonlinearity=nonlinearity,
                            param=param),
            activation(nonlinearity),
            Reshape([-1, base_dim] + list(self.base_shape)),
            EqualizedConv3d(base_dim, base_dim, 3, padding=1, nonlinearity=nonlinearity, param=param),
            activation(nonlinearity),
            ChannelNormalization(),
        )
        filters_in = num_filters(phase - 1, num_phases, base_dim)
        filters_out = num_filters(phase, num_phases, base_dim)
        self.torgb_current = ToRGB(filters_out, self.channels)
        self.torgb_prev = ToRGB(filters_in, self.channels) if phase > 1 else None
        self.blocks = nn.ModuleDict()
        for i in range(2, phase + 1):
            filters_in = num_filters(i - 1, num_phases, base_dim)
            filters_out = num_filters(i, num_phases, base_dim)
            self.blocks[f"block_phase_{i}"] = GeneratorBlock(filters_in, filters_out, nonlinearity, param)
        self.upsample = nn.Upsample(scale_factor=2)
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(self.device)
        _activate_slope(nonlinearity)

    def grow(self):
        """network_dict.py:360-370."""
        self.phase += 1
        filters_in = num_filters(self.phase - 1, self.num_phases, self.base_dim)
        filters_out = num_filters(self.phase, self.num_phases, self.base_dim)
        self.blocks[f"block_phase_{self.phase}"] = GeneratorBlock(filters_in, filters_out, self.nonlinearity,
                                                                  param=self.param)
        self.torgb_prev = self.torgb_current
        self.torgb_current = ToRGB(filters_out, self.channels)
        self.to(self.device)

    def forward(self, input, alpha):
        _activate_slope(self.nonlinearity)
        gin = self.generator_in
        x = gin[0](input.to(self.device), lrelu=True)
        alpha, beta = ops.blend_coef(alpha, x.device)
        x = gin[2](x)
        x = ops.ToAct.apply(x, config.act_dtype(_voxels(x)))
        x = gin[3](x, lrelu=True, premasked=True)
        x = gin[5](x, channels=gin[3].out_channels, mask_input=True)
        x_upsample = None
        for i in range(2, self.phase + 1):
            if i == self.phase:
                x_upsample = ops.Up2.apply(self.torgb_prev(x), 1.0)
            x = self.blocks[f"block_phase_{i}"](x)
        images_out = self.torgb_current(x)
        if x_upsample is not None:
            images_out = ops.Lincomb.apply(x_upsample, images_out, alpha, beta)
        return images_out
