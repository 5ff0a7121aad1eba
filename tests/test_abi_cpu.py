"""The C-ABI library loads without a GPU and exports exactly what include/saragan_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "saragan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    from saragan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    declared = _declared()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sg_version() >= 100
    assert _lib.packed_weight_elems(32, 16, 0) == 27 * 2 * 32 * 8


def test_product_path_has_no_cpu_fallback():
    """CPU tensors must be refused, not silently computed."""
    import torch
    from saragan_b200 import kernels
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        kernels.lincomb(torch.zeros(8), None, 1.0, 0.0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "saragan_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "").replace("CPU oracle", ""), f


def test_error_convention_without_a_gpu():
    """Invalid arguments are refused before any CUDA call: negative status + message from sg_last_error()
    (include/saragan_b200.h "Conventions"); checked on entry points whose validation comes first."""
    import ctypes
    from saragan_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(None)
    cases = [
        ("sg_set_leaky_slope", (ctypes.c_float(2.0),), "slope"),
        ("sg_swd_finish", (null, null, null, 0, null), "batch"),
        ("sg_swd_finish", (null, null, null, 65, null), "batch"),
        ("sg_value_hist", (null, null, 1, 10, ctypes.c_float(-1024.0), 5, 1, null), "value bins"),
        ("sg_swd_descriptors", (null, null, null, null, null, 0, 8, 32, 32, 256, 256 * 243, null), "empty batch"),
        ("sg_swd_descriptors", (null, null, null, null, null, 2, 2, 32, 32, 256, 256 * 243, null), "smaller than"),
        ("sg_swd_project", (null, null, null, null, 4, 100, 50, null), "bad shape"),
        ("sg_pyr_down", (null, null, 1, 0, 4, 4, null), "bad shape"),
        ("sg_pyr_up_sub", (null, null, null, 1, 2, 0, 2, null), "bad shape"),
    ]
    for name, args, needle in cases:
        rc = getattr(lib, name)(*args)
        msg = lib.sg_last_error().decode()
        assert rc < 0, (name, rc)
        assert needle in msg, (name, msg)
    assert abs(lib.sg_get_leaky_slope() - 0.2) < 1e-7          # the refused value was not stored


def test_metrics_refuse_to_run_without_cuda_and_validate_shapes():
    import numpy as np
    import torch
    from saragan_b200 import kernels, metrics
    x = np.zeros((2, 1, 8, 32, 32), dtype=np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        metrics.kolmogorov_smirnov_distance(x, x, -1024, (-1024, 2048))
    with pytest.raises(NotImplementedError, match="single-channel"):
        metrics._volume(np.zeros((2, 3, 8, 32, 32), dtype=np.float32), torch.device("cpu"))
    with pytest.raises(ValueError, match="not twice"):
        kernels.pyr_up_sub(torch.zeros(1, 1, 5, 8, 8), torch.zeros(1, 1, 2, 4, 4))
    with pytest.raises(ValueError, match="smaller than"):
        metrics._draw_positions(np.random.RandomState(0), (2, 1, 2, 32, 32), torch.device("cpu"))
    assert metrics.swd_resolutions(128) == [128, 64, 32] and metrics.swd_resolutions(16) == []
