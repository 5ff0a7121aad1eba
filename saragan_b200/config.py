"""Run-time configuration of the CUDA path."""
import contextlib

import torch

_precision = "bf16"


def set_precision(name: str) -> None:
    """'bf16': bf16 activations/gradients, fp32 accumulation (tcgen05 kind::f16 convolutions).
    'fp32': fp32 storage and CUDA-core convolutions -- the tight-tolerance verification mode."""
    global _precision
    if name not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _precision = name


def precision() -> str:
    return _precision


# Levels with at most this many voxels (the 1x4x4 base level) keep fp32 activations and
# gradients even in bf16 mode: MinibatchStandardDeviation subtracts the group mean there, in
# the forward AND (through autograd) in the backward pass, and early in training the samples
# of a group are nearly identical, so the centred values are small differences of large
# numbers -- bf16 rounding before the subtraction costs 10-30 % relative gradient error
# (measured); fp32 there brings it to ~1 %.  The base level is < 0.2 % of the FLOPs.
FP32_MAX_VOXELS = 16


def act_dtype(voxels: int = 1 << 30) -> torch.dtype:
    """Storage type of an activation / gradient tensor at a level with `voxels` = D*H*W."""
    if _precision == "fp32" or voxels <= FP32_MAX_VOXELS:
        return torch.float32
    return torch.bfloat16


def policy_key():
    """Everything that decides which storage types (hence which weight packings) a pass uses."""
    return (_precision, FP32_MAX_VOXELS)


def describe() -> str:
    """One line for logs / bench.py: which storage and arithmetic each level uses under the current policy."""
    if _precision == "fp32":
        return "fp32 storage, CUDA-core fp32 convolutions at every level"
    return (f"bf16 storage + tcgen05 kind::f16 (fp32 accumulate) above {FP32_MAX_VOXELS} voxels; "
            f"fp32 storage + fp32 CUDA-core kernels at <= {FP32_MAX_VOXELS} voxels (1x4x4 base level, minibatch-stddev)")


@contextlib.contextmanager
def use_precision(name: str):
    old = _precision
    set_precision(name)
    try:
        yield
    finally:
        set_precision(old)
