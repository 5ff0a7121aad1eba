#!/usr/bin/env python
"""Pin oracle/metrics_oracle.py against the UNMODIFIED reference metrics (pgan_pytorch/metrics/{swd,kms}.py,
imported from /root/reference -- available in the build container only) and write the golden fixtures
tests/golden/metrics_*.npz (inputs, seed, the reference's outputs).  The reference draws from numpy's global
RNG; np.random.seed(s) and RandomState(s) produce the same stream, so the comparison is exact.

  python oracle/pin_metrics_against_reference.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/pgan_pytorch")
from metrics import kolmogorov_smirnov_distance as ref_kms  # noqa: E402
from metrics import sliced_wasserstein_distance as ref_swd  # noqa: E402
from metrics import swd as ref_swd_mod                       # noqa: E402
from oracle import metrics_oracle as M                       # noqa: E402


def ct_like(rng, shape, shift=0.0, scale=1.0):
    """volumes in the value range the reference feeds get_metrics: (HU + 1024) / 1024 - 1 style floats"""
    x = rng.randn(*shape).astype(np.float32)
    x = scipy_smooth(x)
    # rounded to fp16-representable values so that the committed fixture can store the inputs in 2 bytes each
    return (shift + scale * 0.35 * x / x.std()).astype(np.float16).astype(np.float32)


def scipy_smooth(x):
    import scipy.ndimage
    return scipy.ndimage.uniform_filter(x, size=(1, 1, 3, 3, 3), mode="nearest")


CASES = {
    # name: (batch, depth, height, width, seed)
    "metrics_w32": (2, 8, 32, 32, 11),       # one pyramid level
    "metrics_w64": (2, 8, 64, 64, 12),       # two levels
    "metrics_w128_b3": (3, 12, 48, 128, 13),  # three levels, non-cubic, odd batch
}

os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
for name, (b, d, h, w, seed) in CASES.items():
    rng = np.random.RandomState(1000 + seed)
    real = ct_like(rng, (b, 1, d, h, w))
    fake = ct_like(rng, (b, 1, d, h, w), shift=0.05, scale=1.15)
    np.random.seed(seed)
    want_swd = [float(v) for v in ref_swd(real.copy(), fake.copy())]
    got_swd = [float(v) for v in M.sliced_wasserstein_distance(real.copy(), fake.copy(), np.random.RandomState(seed))]
    want_kms = float(ref_kms(real.copy(), fake.copy(), -1024, (-1024, 2048)))
    got_kms = float(M.kolmogorov_smirnov_distance(real.copy(), fake.copy(), -1024, (-1024, 2048)))
    # intermediate: the Laplacian pyramid of the reals (for the pyr_down / pyr_up kernels)
    n_levels = len(M.swd_resolutions(w))
    pyr_ref = ref_swd_mod.generate_laplacian_pyramid(real.copy(), n_levels)
    pyr_got = M.generate_laplacian_pyramid(real.copy(), n_levels)
    assert all(np.array_equal(a, c) for a, c in zip(pyr_ref, pyr_got)), name
    assert want_swd == got_swd, (name, want_swd, got_swd)
    assert want_kms == got_kms, (name, want_kms, got_kms)
    out = {"real": real.astype(np.float16), "fake": fake.astype(np.float16), "seed": np.int64(seed), "ref.swd": np.array(want_swd, dtype=np.float64),
           "ref.kms": np.float64(want_kms)}
    if real.size <= 70000:
        for i, p in enumerate(pyr_ref):
            out[f"ref.pyr{i}"] = p.astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(f"{name}: swd {want_swd}  kms {want_kms:.6g}  -- oracle == reference (bit-exact)")
