"""Drop-in for the reference's ``pgan_pytorch/network.py``: same classes, constructor
arguments, attribute / sub-module / parameter names, ``state_dict`` layout, RNG consumption at
construction, ``phase`` / ``alpha`` semantics and return types -- with every operator running in
libsaragan_b200.so (hand-written sm_100a CUDA) on the channel-blocked activation layout.

Deliberate differences from the reference file:
  * the five debug ``print(x.sum())`` host syncs of ``Discriminator.forward``
    (network.py:176-188) and the constructor banner (network.py:259) are not reproduced;
  * ``MinibatchStandardDeviation`` does not mutate its argument; it returns the same values
    the reference's in-place version produces (the concatenated features are the
    group-centred ones, network.py:127-133).

Building blocks accept either plain ``(N,C,D,H,W)`` fp32 tensors (converted on entry and
exit, for stand-alone use) or blocked activations (zero-copy between blocks inside
``Generator`` / ``Discriminator``).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn.modules.utils import _triple

from . import config, kernels, ops

LEAKINESS = 0.2     # nn.LeakyReLU(negative_slope=0.2) everywhere in network.py


def num_filters(phase, num_phases, base_dim):
    """network.py:10-13."""
    num_downscales = int(np.log2(base_dim / 16))
    return min(base_dim // (2 ** (phase - num_phases + num_downscales)), base_dim)


def _fan_in(weight: torch.Tensor) -> int:
    return int(weight[0].numel())


def _voxels(x: torch.Tensor) -> int:
    return int(x.shape[2] * x.shape[3] * x.shape[4])



class _Blocked:
    """Marker mixin: helpers to enter/leave the blocked layout at module boundaries."""

    @staticmethod
    def is_act(x: torch.Tensor) -> bool:
        return x.dim() == 6

    @staticmethod
    def enter(x: torch.Tensor) -> torch.Tensor:
        return ops.ToAct.apply(x.float(), config.act_dtype(_voxels(x)))

    @staticmethod
    def leave(x: torch.Tensor, c: int) -> torch.Tensor:
        return ops.ToPlain.apply(x, c)


class EqualizedConv3d(nn.Module, _Blocked):
    """network.py:26-56.  Weight ~ N(0,1) with the He constant applied at run time
    (``std = 1/sqrt(fan_in)``), bias ~ U(+-1/sqrt(fan_in)).  Supported kernels: 3x3x3/pad 1
    (tcgen05 implicit GEMM) and 1x1x1/pad 0 with a single image channel on one side
    (To/FromRGB)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.Tensor(out_channels))
        self.std = None
        self.reset_parameters()
        self._packed = ops.PackedWeight(self.weight)

    def _gain(self):
        """torch's calculate_gain('conv3d') == 1 (network.py:16-23); network_dict.py's layers override."""
        return 1.0

    def reset_parameters(self):
        fan_in = _fan_in(self.weight)
        self.std = self._gain() / np.sqrt(fan_in)
        with torch.no_grad():
            self.weight.normal_(0, 1)
            bound = 1 / np.sqrt(fan_in)
            self.bias.uniform_(-bound, bound)

    def _kind(self) -> str:
        if self.kernel_size == (3, 3, 3) and self.padding == (1, 1, 1) and self.stride == (1, 1, 1):
            return "3x3x3"
        if self.kernel_size == (1, 1, 1) and self.padding == (0, 0, 0) and self.stride == (1, 1, 1):
            if self.in_channels == 1:
                return "from_rgb"
            if self.out_channels == 1:
                return "to_rgb"
        raise NotImplementedError(
            "saragan_b200.EqualizedConv3d covers the PGAN hot path only: 3x3x3/stride 1/pad 1, "
            "and 1x1x1 with one image channel")

    def forward(self, input, lrelu: bool = False, premasked: bool = False, mask_input_grad: bool = False,
                dd_fuse: bool = False):
        """lrelu fuses the following LeakyReLU(0.2); premasked / mask_input_grad / dd_fuse are the
        LeakyReLU-backward fusion flags of ops.Conv3x3 (internal wiring of the blocks)."""
        kind = self._kind()
        if kind == "3x3x3":
            plain = not self.is_act(input)
            x = self.enter(input) if plain else input
            if self._packed.weight is not self.weight:  # parameter was re-bound (.to(), load)
                self._packed = ops.PackedWeight(self.weight, known=self._packed.known)
            y = ops.Conv3x3.apply(x, self.weight, self.bias, self._packed, float(self.std), lrelu,
                                  premasked and not plain, mask_input_grad and not plain, None,
                                  dd_fuse and not plain)
            return self.leave(y, self.out_channels) if plain else y
        if kind == "from_rgb":
            return ops.PwExpand.apply(input.float().contiguous(), self.weight.reshape(-1), self.bias,
                                      float(self.std), lrelu, self.out_channels,
                                      config.act_dtype(_voxels(input)), premasked, dd_fuse and premasked)
        plain = not self.is_act(input)
        x = self.enter(input) if plain else input
        img = ops.PwReduce.apply(x, self.weight.reshape(-1), self.bias, float(self.std),
                                 self.in_channels)
        return ops.LeakyRelu.apply(img) if lrelu else img


class EqualizedLinear(nn.Module):
    """network.py:59-77 (including its double RNG draw for the weight)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.randn(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features))
        self.std = None
        self.reset_parameters()

    def _gain(self):
        return 1.0

    def reset_parameters(self):
        fan_in = self.weight.shape[1]
        self.std = self._gain() / np.sqrt(fan_in)
        with torch.no_grad():
            self.weight.normal_(0, 1)
            bound = 1 / np.sqrt(fan_in)
            self.bias.uniform_(-bound, bound)

    def forward(self, input, lrelu: bool = False):
        return ops.Linear.apply(input.float().contiguous(), self.weight, self.bias, float(self.std),
                                lrelu)


class DiscriminatorBlock(nn.Sequential, _Blocked):
    """network.py:80-98: conv1 -> lrelu -> conv2 -> lrelu -> AvgPool3d(2) (lrelu fused into
    the conv epilogues)."""

    def __init__(self, filters_in, filters_out):
        super().__init__()
        self.filters_in = filters_in
        self.filters_out = filters_out
        self.conv1 = EqualizedConv3d(filters_in, filters_in, 3, padding=1)
        self.conv2 = EqualizedConv3d(filters_in, filters_out, 3, padding=1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self.downsampling = nn.AvgPool3d(2)

    def forward(self, input, input_is_lrelu: bool = False):
        """input_is_lrelu: `input` is the LeakyReLU output of a premasked, dd_fuse-wired producer (the
        top-level FromRGB) whose only consumer is this block."""
        plain = not self.is_act(input)
        x = self.enter(input) if plain else input
        # LeakyReLU backward masks ride in the consumers' kernels: conv2's dgrad epilogue masks
        # for conv1, the avg-pool backward masks for conv2; in the gradient penalty's double backward
        # they ride in the producers' epilogues (dd_fuse, ops.Conv3x3)
        x = self.conv1(x, lrelu=True, premasked=True, mask_input_grad=input_is_lrelu and not plain, dd_fuse=True)
        conv2 = self.conv2
        if (kernels.conv_pool_supported(x, conv2.in_channels, conv2.out_channels)
                and config.act_dtype(_voxels(x) // 8) == x.dtype):
            # conv2 -> lrelu -> avg-pool in one kernel (the pooling rides in the convolution's epilogue)
            if conv2._packed.weight is not conv2.weight:
                conv2._packed = ops.PackedWeight(conv2.weight, known=conv2._packed.known)
            x = ops.ConvPool.apply(x, conv2.weight, conv2.bias, conv2._packed, float(conv2.std), True, True)
        else:
            x = conv2(x, lrelu=True, premasked=True, mask_input_grad=True, dd_fuse=True)
            x = ops.Down2.apply(x, 0.125, config.act_dtype(_voxels(x) // 8), True, True)
        return self.leave(x, self.filters_out) if plain else x


class FromRGB(nn.Sequential):
    """network.py:101-110: 1x1x1 conv (1 -> filters) + lrelu; takes the fp32 image, returns a
    blocked activation."""

    def __init__(self, channels_in, filters):
        super().__init__()
        self.fromrgb = nn.Sequential(
            EqualizedConv3d(channels_in, filters, 1),
            nn.LeakyReLU(negative_slope=0.2),
        )

    def forward(self, input, premasked: bool = False):
        return self.fromrgb[0](input, lrelu=True, premasked=premasked, dd_fuse=premasked)


class MinibatchStandardDeviation(nn.Module):
    """network.py:113-133 on the plain fp32 (B,C,D,H,W) base-level tensor: hand-written forward,
    backward and double-backward kernels (ops.Mbstd).  The group size follows the reference's rule
    (min(4, B), bumped to the next divisor of B); the returned features are the group-centred ones,
    as the reference's in-place `y -= mean` produces."""

    def __init__(self, group_size=4):
        super().__init__()
        self.group_size = group_size

    def forward(self, input, sub_batches: int = 1):
        """sub_batches > 1: `input` stacks that many independent minibatches along the batch axis
        (each gets the statistics it would get alone)."""
        n = input.shape[0] // sub_batches
        group_size = min(self.group_size, n)
        if group_size < n:
            for i in range(group_size, n + 1):
                if n % i == 0:
                    group_size = i
                    break
        return ops.Mbstd.apply(input.float(), group_size, sub_batches)


class Discriminator(nn.Module):
    """network.py:136-189."""

    def __init__(self, phase, num_phases, base_dim, latent_dim, base_shape):
        super().__init__()
        self.channels = base_shape[0]
        self.base_shape = base_shape[1:]
        self.phase = phase
        if self.channels != 1:
            raise NotImplementedError("saragan_b200 covers single-channel volumes (CT), as the reference's data does")

        self.fromrgbs = nn.ModuleList()
        self.blocks = nn.ModuleList()
        filters_out = base_dim
        for i in reversed(range(2, num_phases + 1)):
            filters_in = num_filters(i, num_phases, base_dim)
            filters_out = num_filters(i - 1, num_phases, base_dim)
            self.blocks.append(DiscriminatorBlock(filters_in, filters_out))
            self.fromrgbs.append(FromRGB(self.channels, filters_in))
        self.fromrgbs.append(FromRGB(self.channels, base_dim))
        self.downscale = nn.AvgPool3d(2)
        self.discriminator_out = nn.Sequential(
            MinibatchStandardDeviation(),
            EqualizedConv3d(filters_out + 1, base_dim, 3, padding=1),
            nn.LeakyReLU(negative_slope=0.2),
            nn.Flatten(),
            EqualizedLinear(int(np.prod(self.base_shape)) * base_dim, latent_dim),
            nn.LeakyReLU(negative_slope=0.2),
            EqualizedLinear(latent_dim, 1),
        )
        self._trunk_channels = filters_out
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.to(self.device)

    def forward(self, input, alpha, sub_batches: int = 1):
        """sub_batches (extension): `input` stacks that many independent minibatches along the batch
        axis -- e.g. D(cat(real, fake)) in one pass; minibatch-stddev treats them separately, so the
        result equals the concatenation of the separate calls."""
        kernels.ensure_leaky_slope(LEAKINESS)
        img = input.to(self.device).float().contiguous()
        alpha, beta = ops.blend_coef(alpha, img.device)
        # at phase > 1 the top FromRGB feeds only the first block's conv1, whose dgrad epilogue
        # then applies FromRGB's LeakyReLU mask
        x = self.fromrgbs[-self.phase](img, premasked=self.phase > 1)
        for i in reversed(range(1, self.phase)):
            x = self.blocks[-i](x, input_is_lrelu=(i == self.phase - 1))
            img = ops.Down2.apply(img, 0.125)
            prev = self.fromrgbs[-i](img)
            x = ops.Lincomb.apply(prev, x, alpha, beta)
        out = self.discriminator_out
        c = out[1].in_channels - 1
        x = out[0](ops.ToPlain.apply(x, c), sub_batches)
        x = out[1](ops.ToAct.apply(x, config.act_dtype(_voxels(x))), lrelu=True)
        x = torch.flatten(ops.ToPlain.apply(x, out[1].out_channels), 1)
        x = out[4](x, lrelu=True)
        return out[6](x)


class ChannelNormalization(nn.Module, _Blocked):
    """network.py:192-197."""

    def __init__(self):
        super().__init__()

    def forward(self, input, lrelu_after: bool = False, channels: int = None, mask_input: bool = False):
        if self.is_act(input):
            assert channels is not None
            return ops.PixelNorm.apply(input, channels, lrelu_after, mask_input)
        c = input.shape[1]
        return self.leave(ops.PixelNorm.apply(self.enter(input), c, lrelu_after), c)


class GeneratorBlock(nn.Sequential, _Blocked):
    """network.py:200-217: up x2 -> conv1 -> lrelu -> pixel-norm -> conv2 -> pixel-norm ->
    lrelu (note the swapped order after the second conv)."""

    def __init__(self, filters_in, filters_out):
        super().__init__()
        self.upsampling = nn.Upsample(scale_factor=2)
        self.conv1 = EqualizedConv3d(filters_in, filters_out, 3, padding=1)
        self.conv2 = EqualizedConv3d(filters_out, filters_out, 3, padding=1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self.cn = ChannelNormalization()

    def forward(self, input):
        plain = not self.is_act(input)
        x = self.enter(input) if plain else input
        c = self.conv1.out_channels
        x = ops.Up2.apply(x, 1.0, config.act_dtype(_voxels(x) * 8, "g"))
        x = self._conv_norm(self.conv1, x, lrelu=True, lrelu_after=False)     # conv1 -> lrelu -> pixel-norm
        x = self._conv_norm(self.conv2, x, lrelu=False, lrelu_after=True)     # conv2 -> pixel-norm -> lrelu
        return self.leave(x, c) if plain else x

    def _conv_norm(self, conv, x, lrelu: bool, lrelu_after: bool):
        """conv + ChannelNormalization (+ the LeakyReLU on either side): one fused tcgen05 kernel where the layer's
        channels fit one N tile of the weight-resident kernel (the two top levels of the benchmarked configurations),
        the convolution and the pixel-norm kernel otherwise."""
        c = conv.out_channels
        if kernels.conv_pixelnorm_supported(x, conv.in_channels, c):
            if conv._packed.weight is not conv.weight:
                conv._packed = ops.PackedWeight(conv.weight, known=conv._packed.known)
            return ops.ConvPixelNorm.apply(x, conv.weight, conv.bias, conv._packed, float(conv.std), lrelu, lrelu_after)
        if lrelu:
            x = conv(x, lrelu=True, premasked=True)            # pixel-norm's backward applies the mask
            return self.cn(x, channels=c, mask_input=True)
        return self.cn(conv(x), lrelu_after=lrelu_after, channels=c)


class ToRGB(nn.Sequential):
    """network.py:219-225: 1x1x1 conv to the single image channel; returns the fp32 image."""

    def __init__(self, filters_in, channels):
        super().__init__()
        self.conv = EqualizedConv3d(filters_in, channels, 1)

    def forward(self, input):
        return self.conv(input)


class Reshape(nn.Module):
    def __init__(self, shape):
        super().__init__()
        self.shape = shape

    def forward(self, input):
        return torch.reshape(input, self.shape)


class Generator(nn.Module):
    """network.py:237-284.  ``forward`` returns the LIST of images at every resolution up to
    ``phase`` (network.py:274-284), each fp32 (N,1,D,H,W)."""

    def __init__(self, phase, num_phases, base_dim, latent_dim, base_shape):
        super().__init__()
        self.channels = base_shape[0]
        self.base_shape = base_shape[1:]
        self.phase = phase
        self.latent_dim = latent_dim
        if self.channels != 1:
            raise NotImplementedError("saragan_b200 covers single-channel volumes (CT), as the reference's data does")
        filters_out = base_dim
        self.generator_in = nn.Sequential(
            EqualizedLinear(latent_dim, int(np.prod(self.base_shape)) * filters_out),
            nn.LeakyReLU(negative_slope=0.2),
            Reshape([-1, filters_out] + list(self.base_shape)),
            EqualizedConv3d(filters_out, filters_out, 3, padding=1),
            nn.LeakyReLU(negative_slope=0.2),
            ChannelNormalization(),
        )
        self.blocks = nn.ModuleList()
        self.to_rgbs = nn.ModuleList([ToRGB(filters_out, self.channels)])
        for i in range(2, num_phases + 1):
            filters_in = num_filters(i, num_phases, base_dim)
            filters_out = num_filters(i + 1, num_phases, base_dim)
            self.blocks.append(GeneratorBlock(filters_in, filters_out))
            self.to_rgbs.append(ToRGB(filters_out, self.channels))
        self.upsample = nn.Upsample(scale_factor=2)
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.to(self.device)

    def forward(self, input, alpha):
        kernels.ensure_leaky_slope(LEAKINESS)
        gin = self.generator_in
        x = gin[0](input.to(self.device), lrelu=True)
        alpha, beta = ops.blend_coef(alpha, x.device)
        x = gin[2](x)
        x = ops.ToAct.apply(x, config.act_dtype(_voxels(x), "g"))
        x = gin[3](x, lrelu=True, premasked=True)
        x = gin[5](x, channels=gin[3].out_channels, mask_input=True)

        all_out = []
        images_out = self.to_rgbs[0](x)
        all_out.append(images_out)
        for i in range(0, self.phase - 1):
            x = self.blocks[i](x)
            img_gen = self.to_rgbs[i + 1](x)
            images_out = ops.Lincomb.apply(ops.Up2.apply(images_out, 1.0), img_gen, alpha, beta)
            all_out.append(images_out)
        return all_out
