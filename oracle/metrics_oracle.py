"""CPU restatement (numpy) of the reference's evaluation metrics -- TEST INFRASTRUCTURE ONLY.

Follows ``pgan_pytorch/metrics/swd.py`` and ``pgan_pytorch/metrics/kms.py`` as written, quirks included
(they are what ``train.py:12-27`` ``get_metrics`` logs every epoch):

* swd.py:8-20  ``get_descriptors_for_minibatch``: the random offsets are drawn with ``size=(N,1,1,1)`` and added
  to 5-D ``ogrid`` axes, so numpy broadcasting puts them on axis 1, not axis 0.  The result has shape
  ``(N, N, 3, 9, 9)`` (N = 128 * batch): axis 0 only selects the IMAGE (``nhood // 128``), axis 1 runs over the N
  random positions, and ``finalize_descriptors`` (swd.py:25-32) standardises per position.  Every image
  therefore contributes 128 identical rows of N*243 components.
* swd.py:12-18: the 4th ogrid axis is called ``x`` but is added to a draw in ``[W, S[4]-W)`` and indexes the LAST
  array axis; the 5th (``y``) indexes the one before.  The draws are made in the order z, x, y.
* swd.py:60-66 ``pyr_up`` scales the 5x5x5 Gaussian by 4 (the 2-D constant) although the zero-insertion is 3-D.
* swd.py:109 builds ``dist + [mean]`` and discards it: the function returns the per-level list only.
* kms.py:19-20: ``np.histogram`` without ``range`` bins every image over its own [min, max].

All random numbers come from a ``numpy.random.RandomState``; ``RandomState(s)`` reproduces the global stream the
reference uses after ``np.random.seed(s)``, which is how ``oracle/pin_metrics_against_reference.py`` checks this
file bit for bit against the unmodified reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import numpy as np
import scipy.ndimage

NHOOD_SIZE = (1, 2, 8, 8)       # swd.py:99
NHOODS_PER_IMAGE = 128          # swd.py:99
DIR_REPEATS, DIRS_PER_REPEAT = 4, 128   # swd.py:108

_f = np.array([1, 4, 6, 4, 1], dtype=np.float32)                                # swd.py:55-58
_f = _f[:, None, None] * _f[None, None, :] * _f[None, :, None]
GAUSSIAN = _f / _f.sum()


def descriptor_positions(rng, shape, nhood_size=NHOOD_SIZE, nhoods_per_image=NHOODS_PER_IMAGE):
    """The three ``randint`` draws of swd.py:14-16 in their order (z, then the last-axis offset, then the
    second-to-last-axis offset), each of length N = nhoods_per_image * batch."""
    S = shape
    N = nhoods_per_image * S[0]
    D, H, W = nhood_size[1] // 2, nhood_size[2] // 2, nhood_size[3] // 2
    z = rng.randint(D, S[2] - D, size=(N, 1, 1, 1))
    x = rng.randint(W, S[4] - W, size=(N, 1, 1, 1))
    y = rng.randint(H, S[3] - H, size=(N, 1, 1, 1))
    return z.reshape(N), x.reshape(N), y.reshape(N)


def get_descriptors_for_minibatch(minibatch, rng, nhood_size=NHOOD_SIZE, nhoods_per_image=NHOODS_PER_IMAGE):
    """swd.py:8-20 (full (N, N, 3, 9, 9) array: small inputs only)."""
    S = minibatch.shape
    N = nhoods_per_image * S[0]
    D, H, W = nhood_size[1] // 2, nhood_size[2] // 2, nhood_size[3] // 2
    nhood, chan, z, x, y = np.ogrid[0:N, 0:nhood_size[0], -D:D + 1, -H:H + 1, -W:W + 1]
    img = nhood // nhoods_per_image
    z = z + rng.randint(D, S[2] - D, size=(N, 1, 1, 1))
    x = x + rng.randint(W, S[4] - W, size=(N, 1, 1, 1))
    y = y + rng.randint(H, S[3] - H, size=(N, 1, 1, 1))
    idx = (((img * S[1] + chan) * S[2] + z) * S[3] + y) * S[4] + x
    return minibatch.flat[idx]


def finalize_descriptors(desc):
    """swd.py:25-32."""
    if isinstance(desc, list):
        desc = np.concatenate(desc, axis=0)
    assert desc.ndim == 5
    desc -= np.mean(desc, axis=(0, 2, 3, 4), keepdims=True)
    desc /= np.std(desc, axis=(0, 2, 3, 4), keepdims=True)
    return desc.reshape(desc.shape[0], -1)


def draw_directions(rng, n_components, dirs_per_repeat=DIRS_PER_REPEAT):
    """swd.py:40-42."""
    dirs = rng.randn(n_components, dirs_per_repeat)
    dirs /= np.sqrt(np.sum(np.square(dirs), axis=0, keepdims=True))
    return dirs.astype(np.float32)


def sliced_wasserstein(A, B, rng, dir_repeats=DIR_REPEATS, dirs_per_repeat=DIRS_PER_REPEAT):
    """swd.py:36-51."""
    assert A.ndim == 2 and A.shape == B.shape
    results = []
    for _ in range(dir_repeats):
        dirs = draw_directions(rng, A.shape[1], dirs_per_repeat)
        projA = np.sort(np.matmul(A, dirs), axis=0)
        projB = np.sort(np.matmul(B, dirs), axis=0)
        results.append(np.mean(np.abs(projA - projB)))
    return np.mean(results)


def pyr_down(minibatch):
    """swd.py:61-63."""
    assert minibatch.ndim == 5
    return scipy.ndimage.convolve(minibatch, GAUSSIAN[None, None], mode='mirror')[:, :, ::2, ::2, ::2]


def pyr_up(minibatch):
    """swd.py:65-70."""
    assert minibatch.ndim == 5
    S = minibatch.shape
    res = np.zeros((S[0], S[1], S[2] * 2, S[3] * 2, S[4] * 2), minibatch.dtype)
    res[:, :, ::2, ::2, ::2] = minibatch
    return scipy.ndimage.convolve(res, GAUSSIAN[None, None] * 4.0, mode='mirror')


def generate_laplacian_pyramid(minibatch, num_levels):
    """swd.py:73-78."""
    pyramid = [np.float32(minibatch)]
    for _ in range(1, num_levels):
        pyramid.append(pyr_down(pyramid[-1]))
        pyramid[-2] -= pyr_up(pyramid[-1])
    return pyramid


def swd_resolutions(width):
    """swd.py:89-93."""
    res, out = width, []
    while res >= 32:
        out.append(res)
        res //= 2
    return out


def sliced_wasserstein_distance(minibatch_real, minibatch_fake, rng):
    """swd.py:87-111; returns the per-level list (x 1e3), highest resolution first.  Order of the random draws:
    descriptor positions for every level of the real pyramid, then of the fake pyramid, then per level the
    directions of the four repeats."""
    n = len(swd_resolutions(minibatch_real.shape[-1]))
    desc_real = [get_descriptors_for_minibatch(level, rng) for level in generate_laplacian_pyramid(minibatch_real, n)]
    desc_fake = [get_descriptors_for_minibatch(level, rng) for level in generate_laplacian_pyramid(minibatch_fake, n)]
    desc_real = [finalize_descriptors(d) for d in desc_real]
    desc_fake = [finalize_descriptors(d) for d in desc_fake]
    dist = [sliced_wasserstein(a, b, rng) for a, b in zip(desc_real, desc_fake)]
    return [d * 1e3 for d in dist]


def kolmogorov_smirnov_distance(real_images, fake_images, intercept, clip_range):
    """kms.py:4-24."""
    real_images = ((real_images * intercept) + intercept).astype(int)
    fake_images = ((fake_images * intercept) + intercept).astype(int)
    real_images = real_images.clip(*clip_range)
    fake_images = fake_images.clip(*clip_range)
    fake_images = fake_images.mean(1)
    real_images = real_images.mean(1)
    real_images = real_images.reshape(real_images.shape[0], -1)
    fake_images = fake_images.reshape(real_images.shape[0], -1)
    bins = clip_range[1] - clip_range[0]
    real_hists = np.stack([np.histogram(real_images[i], bins=bins, density=True)[0] for i in range(real_images.shape[0])])
    fake_hists = np.stack([np.histogram(fake_images[i], bins=bins, density=True)[0] for i in range(fake_images.shape[0])])
    return abs(real_hists.mean(0) - fake_hists.mean(0)).max()


def get_metrics(x_real, x_fake, rng):
    """train.py:12-27 (the labels are the reference's: 'mean_swd' is the LOWEST-resolution level)."""
    kms = kolmogorov_smirnov_distance(x_real, x_fake, -1024, (-1024, 2048))
    swds = sliced_wasserstein_distance(x_real, x_fake, rng) if x_real.shape[-1] >= 32 else []
    d = {}
    for i, swd in enumerate(reversed(swds)):
        if i == 0:
            d['mean_swd'] = swd
        else:
            d[f'swd_{32 * 2 ** i}'] = swd
    d['kms'] = kms
    return d
