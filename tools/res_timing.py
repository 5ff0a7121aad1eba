#!/usr/bin/env python
"""Where the weight-resident conv kernel waits (diagnostic build only):
    make -C saragan_b200/csrc BUILD=build_timing OUT=../libsaragan_b200_timing.so EXTRA=-DSG_RES_TIMING
    SARAGAN_B200_LIB=saragan_b200/libsaragan_b200_timing.so python tools/res_timing.py
Per shape: launch time, and (mean over the 148 CTAs) the cycles the MMA issuer spent waiting for halo stages (A_FULL) and
for a drained accumulator set (ACC_EMPTY), the producer waiting for a free stage, the epilogue waiting for ACC_FULL."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import _lib, kernels as K  # noqa: E402

lib = _lib.load()
lib.sg_tc_res_timing.argtypes = [ctypes.c_void_p]
BF = torch.bfloat16
SHAPES = [(4, 32, 32), (4, 64, 32), (4, 32, 64), (4, 16, 16)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n, cin, cout in SHAPES:
    d, h, w = 32, 128, 128
    x = K.plain_to_act(torch.randn(n, cin, d, h, w, device="cuda"), BF)
    wp = K.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, device="cuda"), BF, False)
    bias = torch.randn(cout, device="cuda")
    for force in ((0, 0), (4, 2), (2, 2)):
        lib.sg_tc_res_force(*force)
        ts = []
        for i in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.conv3d_fprop(x, wp, bias, None, cin, cout, 0.05, True, _lib.IMPL_TCGEN05)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out = (ctypes.c_longlong * (148 * 8))()
        lib.sg_tc_res_timing(ctypes.cast(out, ctypes.c_void_p))
        t = torch.tensor(list(out), dtype=torch.float64).view(148, 8)
        m = t.mean(0)
        print(f"{n}x{cin}->{cout} force(td,kb)={force}: {min(ts[1:]):.1f} us | issuer total {m[0]:.0f} cyc ({m[6]:.1f} tiles, {m[0] / m[6]:.0f}/tile), "
              f"wait A_FULL {m[1]:.0f} ({100 * m[1] / m[0]:.0f} %), wait ACC_EMPTY {m[2]:.0f} ({100 * m[2] / m[0]:.0f} %) | producer total {m[7]:.0f}, wait A_EMPTY {m[3]:.0f} "
              f"| epilogue total {m[4]:.0f}, wait ACC_FULL {m[5]:.0f} ({100 * m[5] / m[4]:.0f} %)", flush=True)
    lib.sg_tc_res_force(0, 0)
