"""Run-time configuration of the CUDA path."""
import contextlib
import os

import torch

_precision = "bf16"
PRECISIONS = ("bf16", "tf32", "fp32")


def set_precision(name: str) -> None:
    """'bf16': bf16 activations/gradients + tcgen05 kind::f16 convolutions (fp32 accumulation) at the levels above
            TF32_MAX_VOXELS voxels; fp32 storage + tcgen05 kind::tf32 at the levels up to TF32_MAX_VOXELS; plain fp32
            (CUDA cores) at the base level (<= FP32_MAX_VOXELS).  The benchmarked policy.
    'tf32': fp32 storage everywhere, tcgen05 kind::tf32 convolutions (operands rounded to nearest to 10 mantissa bits,
            fp32 accumulation) above the base level -- the <= 1e-3 tier of BASELINE.json on tensor cores.
    'fp32': fp32 storage and CUDA-core fp32 convolutions everywhere -- the exact verification mode."""
    global _precision
    if name not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}")
    _precision = name


def precision() -> str:
    return _precision


# Levels with at most this many voxels (the 1x4x4 base level) keep fp32 activations and
# gradients AND exact fp32 arithmetic in every mode: MinibatchStandardDeviation subtracts the group mean there, in
# the forward AND (through autograd) in the backward pass, and early in training the samples
# of a group are nearly identical, so the centred values are small differences of large
# numbers -- bf16 rounding before the subtraction costs 10-30 % relative gradient error
# (measured); fp32 there brings it to ~1 %.  The base level is < 0.2 % of the FLOPs.
FP32_MAX_VOXELS = 16

# bf16 mode: levels with at most this many voxels (up to 4x16x16: < 9 % of the step's FLOPs, all of them latency- and
# weight-stream-bound) keep fp32 storage and run their convolutions as TF32 on the tensor cores.  Their pre-activations
# feed the deepest part of the gradient chain (and the minibatch-stddev statistics); at TF32 the whole-step parameter
# gradients of the golden configurations sit at 1.4 % median against the fp32 reference (bf16 there: 5 %).
# (SARAGAN_TF32_MAX_VOXELS overrides it for the per-policy table of DESIGN.md / profiles/: 16 = bf16 everywhere above the
# base level, 8192 = TF32 up to 8x32x32.)
TF32_MAX_VOXELS = int(os.environ.get("SARAGAN_TF32_MAX_VOXELS", "1024"))


# The generator's own threshold (experiment knob; default = the discriminator's): its low-resolution activations do not
# feed the minibatch-stddev statistics, only D's do.
G_TF32_MAX_VOXELS = int(os.environ.get("SARAGAN_G_TF32_MAX_VOXELS", str(TF32_MAX_VOXELS)))


def act_dtype(voxels: int = 1 << 30, net: str = "d") -> torch.dtype:
    """Storage type of an activation / gradient tensor at a level with `voxels` = D*H*W (`net`: "g" inside the
    generator, "d" elsewhere)."""
    limit = G_TF32_MAX_VOXELS if net == "g" else TF32_MAX_VOXELS
    if _precision != "bf16" or voxels <= max(FP32_MAX_VOXELS, limit):
        return torch.float32
    return torch.bfloat16


def conv_route(x: torch.Tensor, cin: int, cout: int):
    """(impl code, packed-weight kind) of a 3x3x3 convolution on the blocked activation `x` under the current policy:
    bf16 tensors -> tcgen05 kind::f16; fp32 tensors above the base level -> tcgen05 kind::tf32 (modes 'bf16', 'tf32')
    when the kernel covers the shape; everything else -> the fp32 CUDA-core kernels."""
    from . import _lib
    if x.dtype == torch.float32 and _precision != "fp32":
        n, _, d, h, w, _ = x.shape
        if d * h * w > FP32_MAX_VOXELS and tf32_supported(n, cin, cout, d, h, w):
            return _lib.IMPL_TF32, "tf32"
    return _lib.IMPL_AUTO, x.dtype


def wgrad_impl(x: torch.Tensor, cin: int, cout: int) -> int:
    """impl code of the weight gradient: like conv_route, except that under the 'bf16' policy the fp32-storage levels
    round the wgrad operands to bf16 (one MMA per product instead of the split's three): 3e-3 of noise on a weight
    gradient, no mask involved -- the accuracy the bf16 levels' weight gradients have anyway."""
    from . import _lib
    impl = conv_route(x, cin, cout)[0]
    if impl == _lib.IMPL_TF32 and _precision == "bf16":
        return _lib.IMPL_F32_AS_BF16
    return impl


_tf32_ok = {}


def tf32_supported(n, cin, cout, d, h, w) -> bool:
    key = (n, cin, cout, d, h, w)
    if key not in _tf32_ok:
        from . import kernels
        _tf32_ok[key] = kernels.conv_tf32_supported(*key)
    return _tf32_ok[key]


def policy_key():
    """Everything that decides which storage types (hence which weight packings) a pass uses."""
    return (_precision, FP32_MAX_VOXELS, TF32_MAX_VOXELS, G_TF32_MAX_VOXELS)


def describe() -> str:
    """One line for logs / bench.py: which storage and arithmetic each level uses under the current policy."""
    base = f"fp32 storage + fp32 CUDA-core kernels at <= {FP32_MAX_VOXELS} voxels (1x4x4 base level, minibatch-stddev)"
    if _precision == "fp32":
        return "fp32 storage, CUDA-core fp32 convolutions at every level"
    if _precision == "tf32":
        return f"fp32 storage + tcgen05 kind::tf32 (rounded to nearest, fp32 accumulate) above {FP32_MAX_VOXELS} voxels; " + base
    return (f"bf16 storage + tcgen05 kind::f16 (fp32 accumulate) above {TF32_MAX_VOXELS} voxels; fp32 storage + tcgen05 "
            f"kind::tf32 from {FP32_MAX_VOXELS + 1} to {TF32_MAX_VOXELS} voxels; " + base)


@contextlib.contextmanager
def use_precision(name: str):
    old = _precision
    set_precision(name)
    try:
        yield
    finally:
        set_precision(old)
