"""ctypes binding of libsaragan_b200.so (include/saragan_b200.h).

Thin, typed wrappers: torch tensors in, raw device pointers + sizes out.  There is no
fallback of any kind: if the shared library is missing or a tensor is not on a CUDA
device the call raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SARAGAN_B200_LIB: load another build of the same library (A/B runs of build-time switches, e.g. SG_TC_WATCHDOG)
LIB_PATH = os.environ.get("SARAGAN_B200_LIB") or os.path.join(_HERE, "libsaragan_b200.so")

BF16, F32, TF32_PACK = 0, 1, 2
IMPL_AUTO, IMPL_DIRECT, IMPL_TCGEN05, IMPL_TF32, IMPL_F32_AS_BF16 = 0, 1, 2, 3, 4

_c_int, _c_i64, _c_f, _c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (all return int unless listed in _RESTYPES); mirrors include/saragan_b200.h
SIGNATURES = {
    "sg_version": [],
    "sg_last_error": [],
    "sg_launch_count": [_c_int],
    "sg_cuda_core_fallbacks": [_c_int],
    "sg_plain_to_act": [_c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_p],
    "sg_act_to_plain": [_c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_p],
    "sg_packed_weight_elems": [_c_int, _c_int, _c_int],
    "sg_pack_conv_weight": [_c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_p],
    "sg_pack_conv_weights_multi": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_p],
    "sg_conv3d_fprop": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                        _c_int, _c_f, _c_int, _c_int, _c_p, _c_i64, _c_p],
    "sg_set_pdl": [_c_int],
    "sg_set_leaky_slope": [_c_f],
    "sg_get_leaky_slope": [],
    "sg_tc_force_streaming": [_c_int],
    "sg_tc_res_zs_mode": [_c_int],
    "sg_tc_res_force": [_c_int, _c_int],
    "sg_tc_force_plan": [_c_int, _c_int, _c_int, _c_int],
    "sg_tc_plan_debug": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.POINTER(ctypes.c_int)],
    "sg_conv3d_pixelnorm_supported": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "sg_conv3d_fprop_pixelnorm": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f,
                                  _c_int, _c_int, _c_f, _c_p],
    "sg_conv3d_pool_supported": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "sg_conv3d_fprop_pool": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f,
                             _c_int, _c_f, _c_p],
    "sg_conv3d_tf32_supported": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "sg_conv3d_workspace_bytes": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "sg_conv3d_wgrad": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                        _c_f, _c_int, _c_p, _c_i64, _c_p],
    "sg_pw_expand": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_int, _c_p],
    "sg_pw_expand_masked": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_int, _c_p],
    "sg_pw_reduce": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_p],
    "sg_pw_wgrad": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_p],
    "sg_down2": [_c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_f, _c_p],
    "sg_up2": [_c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_f, _c_p],
    "sg_lincomb": [_c_p, _c_p, _c_p, _c_int, _c_i64, _c_f, _c_f, _c_p],
    "sg_lincomb_dev": [_c_p, _c_p, _c_p, _c_int, _c_i64, _c_p, _c_p],
    "sg_lrelu_fwd": [_c_p, _c_p, _c_int, _c_i64, _c_p],
    "sg_mask_mul": [_c_p, _c_p, _c_p, _c_int, _c_i64, _c_p],
    "sg_pixelnorm_fwd": [_c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_int, _c_p],
    "sg_pixelnorm_bwd": [_c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_i64, _c_f, _c_int, _c_int, _c_p],
    "sg_interp": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_i64, _c_p],
    "sg_sumsq_rows": [_c_p, _c_p, _c_int, _c_i64, _c_p],
    "sg_rowscale": [_c_p, _c_p, _c_p, _c_int, _c_i64, _c_p],
    "sg_linear_fwd": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_f, _c_int, _c_p],
    "sg_linear_dgrad": [_c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_f, _c_p],
    "sg_linear_wgrad": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_f, _c_p],
    "sg_mbstd_fwd": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_p],
    "sg_mbstd_bwd": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_p],
    "sg_mbstd_bwdbwd": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_p],
    "sg_adam_step": [_c_p, _c_p, _c_p, _c_int, _c_p, _c_f, _c_p, _c_f, _c_f, _c_f, _c_f, _c_p],
    "sg_adam_advance": [_c_p, _c_p],
    "sg_multi_copy_scale": [_c_p, _c_p, _c_p, _c_int, _c_f, _c_p],
    "sg_prepare_real": [_c_p, _c_p, _c_p, _c_i64, _c_f, _c_f, _c_p],
    "sg_pyr_down": [_c_p, _c_p, _c_i64, _c_int, _c_int, _c_int, _c_p],
    "sg_pyr_up_sub": [_c_p, _c_p, _c_p, _c_i64, _c_int, _c_int, _c_int, _c_p],
    "sg_swd_descriptors": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_i64, _c_p],
    "sg_swd_project": [_c_p, _c_p, _c_p, _c_p, _c_int, _c_i64, _c_i64, _c_p],
    "sg_swd_finish": [_c_p, _c_p, _c_p, _c_int, _c_p],
    "sg_value_hist": [_c_p, _c_p, _c_int, _c_i64, _c_f, _c_int, _c_int, _c_p],
}
_RESTYPES = {"sg_last_error": ctypes.c_char_p, "sg_packed_weight_elems": ctypes.c_int64,
             "sg_conv3d_workspace_bytes": ctypes.c_int64, "sg_launch_count": ctypes.c_int64, "sg_cuda_core_fallbacks": ctypes.c_int64,
             "sg_set_pdl": None, "sg_get_leaky_slope": ctypes.c_float, "sg_tc_force_streaming": None, "sg_tc_res_zs_mode": None, "sg_tc_res_force": None, "sg_tc_force_plan": None}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C saragan_b200/csrc`). saragan_b200 has no CPU or "
                "library fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if os.environ.get("SARAGAN_PDL", "0") == "1":   # A/B switch for programmatic dependent launch
            lib.sg_set_pdl(1)
        _lib = lib
    return _lib


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"saragan_b200: unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("saragan_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("saragan_b200 kernels need contiguous tensors")
    p = t.data_ptr()
    if p % 16 and t.numel():
        raise RuntimeError("saragan_b200 kernels need 16-byte aligned tensors")
    return p


def _stream():
    return torch.cuda.current_stream().cuda_stream


# tools/profile_step.py / bench.py set this to a list to collect (name, int-args, start event, end event, pointer
# flags) for every ABI call (CUDA events on the launching stream; no effect when None); pointer flags = which of the
# tensor arguments were non-null (e.g. whether an up-sampling call carried a mask).
PROFILE = None


def call(name: str, *args):
    """Invoke an ABI function; tensors are turned into device pointers, the current torch
    stream is appended, a non-zero status raises RuntimeError(sg_last_error())."""
    lib = load()
    conv = [(_ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.device.index != torch.cuda.current_device():
                # launch on the tensors' device (and its current stream), not on whatever device is current
                with torch.cuda.device(a.device):
                    return call(name, *args)
            break
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        rc = getattr(lib, name)(*conv, _stream())
        e1.record(torch.cuda.current_stream())
        PROFILE.append((name, tuple(a for a in args if isinstance(a, (int, bool))), e0, e1,
                        tuple(a is not None for a in args if a is None or isinstance(a, torch.Tensor))))
    else:
        rc = getattr(lib, name)(*conv, _stream())
    if rc != 0:
        msg = lib.sg_last_error()
        raise RuntimeError(f"{name} failed (status {rc}): {msg.decode() if msg else ''}")


def launch_count(reset: bool = False) -> int:
    """Kernels launched by libsaragan_b200 in this process so far."""
    return int(load().sg_launch_count(int(reset)))


def packed_weight_elems(cout: int, cin: int, flip: int) -> int:
    return int(load().sg_packed_weight_elems(cout, cin, flip))


def conv_workspace_bytes(kind: int, dtype: int, n: int, cin: int, cout: int, d: int, h: int, w: int) -> int:
    return int(load().sg_conv3d_workspace_bytes(kind, dtype, n, cin, cout, d, h, w))


def chunks(c: int) -> int:
    """Number of 8-channel chunks of the blocked layout (channels padded to 16)."""
    return 2 * ((c + 15) // 16)
