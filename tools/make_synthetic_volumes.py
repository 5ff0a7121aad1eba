#!/usr/bin/env python
"""Synthetic stand-in for the reference's LIDC-IDRI pyramid: for each growth phase p a directory
``{res}x{res}/`` (res = 4 * 2**(p-1)) of ``NNNN.npy`` files, shape ``(res/4, res, res)`` (D, H, W), dtype uint16,
values ``clip(1024 + 350 * smooth(N(0,1)), 0, 3072)`` -- the range and offset data_scripts/create_lidc_idri_dataset.py:
185-212 produces (HU clipped to [-1024, 2048], + 1024) -- drawn from ``numpy.random.default_rng(1234 + p)``.
``saragan_b200.data.VolumeLoader(root/{res}x{res}, ...)`` reads them the way main.py:69-91 reads the real ones.

    python tools/make_synthetic_volumes.py OUT_DIR [--phases 1 2 3 4 5] [--count 64]
"""
import argparse
import os

import numpy as np
import scipy.ndimage


def synthetic_volume(rng, res):
    x = rng.standard_normal((max(res // 4, 1), res, res)).astype(np.float32)
    x = scipy.ndimage.uniform_filter(x, size=3, mode="nearest")
    x = x / max(float(x.std()), 1e-6)
    return np.clip(1024 + 350 * x, 0, 3072).astype(np.uint16)


def write_phase(root, phase, count):
    res = 4 * 2 ** (phase - 1)
    d = os.path.join(root, f"{res}x{res}")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(1234 + phase)
    for i in range(count):
        np.save(os.path.join(d, f"{i:04d}.npy"), synthetic_volume(rng, res))
    return d


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--phases", type=int, nargs="+", default=[1, 2, 3, 4, 5])
    ap.add_argument("--count", type=int, default=64)
    a = ap.parse_args()
    for p in a.phases:
        print(write_phase(a.out, p, a.count))
