"""GPU drop-in for the reference's evaluation metrics: ``pgan_pytorch/metrics/swd.py``
(3-D sliced Wasserstein distance over a Laplacian pyramid), ``metrics/kms.py`` (Kolmogorov-Smirnov histogram
distance) and ``train.py:12-27`` ``get_metrics`` -- same names, arguments and return values, computed by
libsaragan_b200.so on the device the volumes already live on (the reference copies both batches to the host and
runs numpy/scipy there after every stabilising epoch, train.py:76,99,120).

Inputs are (N, 1, D, H, W) fp32 volumes, torch tensors (any device; moved to the current CUDA device) or numpy
arrays.  Random numbers: the reference draws neighbourhood positions and projection directions from numpy's GLOBAL
generator.  ``rng=None`` draws them on the device (torch's CUDA generator; same estimator, different numbers, no
host work); ``rng=numpy.random`` (the module, i.e. the global state the reference uses) or a
``numpy.random.RandomState`` draws the reference's exact numbers in the reference's exact order on the host, which
is how the parity tests reproduce its outputs.

What swd.py computes is restated, quirks included, in the header of ``csrc/metrics.cu``; the sliced Wasserstein
estimate needs ``batch`` rows per arm instead of the reference's ``128 * batch`` (its rows repeat 128 times).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import kernels as K

NHOOD_SIZE = (1, 2, 8, 8)               # swd.py:99
NHOODS_PER_IMAGE = 128                  # swd.py:99
DIR_REPEATS, DIRS_PER_REPEAT = 4, 128   # swd.py:108
PATCH = 3 * 9 * 9


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("saragan_b200.metrics runs on a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _volume(x, dev=None) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.dim() != 5 or x.shape[1] != 1:
        raise NotImplementedError("saragan_b200.metrics covers (N, 1, D, H, W) single-channel volumes")
    return x.to(_device() if dev is None else dev, torch.float32).contiguous()


def pyr_down(minibatch: torch.Tensor) -> torch.Tensor:
    """swd.py:61-63."""
    return K.pyr_down(minibatch)


def generate_laplacian_pyramid(minibatch: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    """swd.py:73-78."""
    pyramid = [minibatch.float().contiguous()]
    for _ in range(1, num_levels):
        pyramid.append(K.pyr_down(pyramid[-1]))
        pyramid[-2] = K.pyr_up_sub(pyramid[-2], pyramid[-1])
    return pyramid


def swd_resolutions(width: int) -> List[int]:
    """swd.py:89-93."""
    res, out = width, []
    while res >= 32:
        out.append(res)
        res //= 2
    return out


def _draw_positions(rng, shape, dev):
    """swd.py:14-16: three draws of N = 128 * batch offsets, in the order z, last-axis offset, second-to-last-axis
    offset.  Returns int32 device tensors (z, y, x) with x indexing the LAST axis."""
    n = NHOODS_PER_IMAGE * shape[0]
    d, h, w = NHOOD_SIZE[1] // 2, NHOOD_SIZE[2] // 2, NHOOD_SIZE[3] // 2
    if shape[2] - d <= d or shape[4] - w <= w or shape[3] - h <= h:
        raise ValueError(f"volume {tuple(shape[2:])} is smaller than the 3x9x9 neighbourhood")   # numpy: low >= high
    if rng is None:
        z = torch.randint(d, shape[2] - d, (n,), device=dev, dtype=torch.int32)
        x = torch.randint(w, shape[4] - w, (n,), device=dev, dtype=torch.int32)
        y = torch.randint(h, shape[3] - h, (n,), device=dev, dtype=torch.int32)
        return z, y, x
    z = rng.randint(d, shape[2] - d, size=(n, 1, 1, 1))
    x = rng.randint(w, shape[4] - w, size=(n, 1, 1, 1))
    y = rng.randint(h, shape[3] - h, size=(n, 1, 1, 1))
    to = lambda a: torch.from_numpy(a.reshape(n).astype(np.int32)).to(dev)    # noqa: E731
    return to(z), to(y), to(x)


def _draw_directions(rng, k: int, dev):
    """swd.py:40-42.  Host draw: normalised in fp64 and cast like the reference; device draw: raw normals, the
    column norms are accumulated by the projection kernel and applied in the finishing kernel."""
    if rng is None:
        return torch.randn((k, DIRS_PER_REPEAT), device=dev, dtype=torch.float32), True
    dirs = rng.randn(k, DIRS_PER_REPEAT)
    dirs /= np.sqrt(np.sum(np.square(dirs), axis=0, keepdims=True))
    return torch.from_numpy(dirs.astype(np.float32)).to(dev), False


def sliced_wasserstein_distance(minibatch_real, minibatch_fake, rng=None) -> List[float]:
    """swd.py:87-111: the per-level distances (x 1e3), highest resolution first (the reference also builds
    ``dist + [mean]`` and discards it, swd.py:109)."""
    real = _volume(minibatch_real)
    fake = _volume(minibatch_fake, real.device)
    assert real.shape == fake.shape, (real.shape, fake.shape)
    dev, b = real.device, real.shape[0]
    levels = len(swd_resolutions(real.shape[-1]))
    if levels == 0:
        return []
    pyr_real = generate_laplacian_pyramid(real, levels)
    pyr_fake = generate_laplacian_pyramid(fake, levels)
    # the reference's draw order: positions for every level of the reals, then of the fakes, then the directions
    pos_real = [_draw_positions(rng, lv.shape, dev) for lv in pyr_real]
    pos_fake = [_draw_positions(rng, lv.shape, dev) for lv in pyr_fake]
    k = NHOODS_PER_IMAGE * b * PATCH
    a = torch.empty((2 * b, k), dtype=torch.float32, device=dev)        # real rows, then fake rows
    out = torch.zeros((levels * DIR_REPEATS, 4), dtype=torch.float32, device=dev)
    for i in range(levels):
        K.swd_descriptors(pyr_real[i], *pos_real[i], a[:b])
        K.swd_descriptors(pyr_fake[i], *pos_fake[i], a[b:])
        for rep in range(DIR_REPEATS):
            dirs, raw = _draw_directions(rng, k, dev)
            p, colsq = K.swd_project(a, dirs, raw)
            K.swd_finish(p, colsq, out[i * DIR_REPEATS + rep])
    vals = out[:, 0].cpu().numpy().reshape(levels, DIR_REPEATS)          # the one host sync
    return [np.mean([v for v in row]) * 1e3 for row in vals]              # fp32 like the reference's np.mean(list)


def kolmogorov_smirnov_distance(real_images, fake_images, intercept, clip_range) -> float:
    """kms.py:4-24.  The device counts how often every integer value of ``clip(int(x*i + i))`` occurs per image
    (one pass over the volumes); numpy's own histogram code then bins those <= 3073 distinct values with their
    counts as weights, which reproduces ``np.histogram(image, bins, density=True)`` of kms.py:19-20 exactly."""
    real = _volume(real_images)
    fake = _volume(fake_images, real.device)
    lo, hi = int(clip_range[0]), int(clip_range[1])
    bins = hi - lo
    values = np.arange(lo, hi + 1, dtype=np.float64)

    def density(x):
        counts = K.value_hist(x.reshape(x.shape[0], -1), float(intercept), lo, hi).cpu().numpy()
        hists = []
        for c in counts:
            nz = c > 0
            hists.append(np.histogram(values[nz], bins=bins, weights=c[nz].astype(np.int64), density=True)[0])
        return np.stack(hists)

    return abs(density(real).mean(0) - density(fake).mean(0)).max()


def get_metrics(x_real, x_fake, rng=None) -> Dict[str, float]:
    """train.py:12-27 (labels as the reference's: 'mean_swd' is the LOWEST-resolution level, then swd_64, ...)."""
    kms = kolmogorov_smirnov_distance(x_real, x_fake, -1024, (-1024, 2048))
    swds = sliced_wasserstein_distance(x_real, x_fake, rng) if x_real.shape[-1] >= 32 else []
    d_dict = {}
    for i, swd in enumerate(reversed(swds)):
        if i == 0:
            d_dict['mean_swd'] = swd
        else:
            d_dict[f'swd_{32 * 2 ** i}'] = swd
    d_dict['kms'] = kms
    return d_dict
