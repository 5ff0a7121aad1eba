"""The tcgen05 tiling planner is host code (sg_tc_plan_debug): without a GPU, every 3x3x3 convolution of every BASELINE
configuration -- in its fprop and its dgrad role, at the per-GPU batch and at the stacked D(real, fake) batch, for both
network variants -- must get a tcgen05 plan within the SM's resources (no silent CUDA-core fallback in bf16 mode;
bench.py reports `cuda_core_conv_fallbacks_per_step` = 0 on the GPU)."""
import ctypes

import numpy as np
import pytest

from saragan_b200 import _lib
from saragan_b200 import costmodel as C
from saragan_b200.network import num_filters

SMEM_MAX = 227 * 1024      # bytes of shared memory a CTA can opt in to on sm_100
TMEM_COLS = 512


def _layers(cfg, variant):
    out = [(ci, co, v) for kind in ("g", "d")
           for _, ci, co, v in C.conv_layers(kind, cfg["phase"], cfg["num_phases"], cfg["base_dim"])]
    if variant == "network_dict":          # generator blocks are f(i-1) -> f(i) -> f(i) there (network_dict.py:352-356)
        f = lambda i: int(num_filters(i, cfg["num_phases"], cfg["base_dim"]))      # noqa: E731
        out = [(ci, co, v) for _, ci, co, v in C.conv_layers("d", cfg["phase"], cfg["num_phases"], cfg["base_dim"])]
        for i in range(2, cfg["phase"] + 1):
            out += [(f(i - 1), f(i), C.volume(i)), (f(i), f(i), C.volume(i))]
    return [(ci, co, v) for ci, co, v in out if int(np.prod(v)) > 16]      # the 1x4x4 base level is fp32 (config.py)


@pytest.mark.parametrize("variant", ["network", "network_dict"])
@pytest.mark.parametrize("name", sorted(C.CONFIGS))
def test_every_convolution_has_a_tcgen05_plan(name, variant):
    lib = _lib.load()
    cfg = C.CONFIGS[name]
    out = (ctypes.c_int * 16)()
    checked = 0
    for ci, co, v in _layers(cfg, variant):
        for a, b in ((ci, co), (co, ci)):                  # fprop, and dgrad = fprop with the channel roles swapped
            for n in (cfg["batch"], 2 * cfg["batch"]):
                rc = lib.sg_tc_plan_debug(n, a, b, v[0], v[1], v[2], out)
                ok, nt, smem, tmem = out[0], out[1], out[13], out[14]
                assert rc == 0 and ok == 1, (name, variant, n, a, b, v, lib.sg_last_error())
                assert nt in (16, 32, 64, 128) and 0 < smem <= SMEM_MAX and 0 < tmem <= TMEM_COLS, (n, a, b, v, list(out))
                assert out[10] >= 1 and out[11] >= 1 and out[12] >= 1
                checked += 1
    assert checked >= 16
