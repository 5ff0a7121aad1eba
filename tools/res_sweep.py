#!/usr/bin/env python
"""Time the weight-resident conv kernel over its tile depth / K-block choices (sg_tc_res_force) on the top-level
shapes, with and without the LeakyReLU-mask epilogue.  python tools/res_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import _lib, kernels as K  # noqa: E402

BF = torch.bfloat16
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


for (n, cin, cout, vol) in [(4, 32, 32, (32, 128, 128)), (4, 64, 32, (32, 128, 128)), (4, 32, 64, (32, 128, 128)),
                            (4, 16, 16, (32, 128, 128)), (4, 32, 16, (32, 128, 128)), (2, 16, 32, (64, 256, 256)),
                            (4, 64, 64, (16, 64, 64))]:
    d, h, w = vol
    x = K.plain_to_act(torch.randn(n, cin, d, h, w, device="cuda"), BF)
    m = K.plain_to_act(torch.randn(n, cout, d, h, w, device="cuda"), BF)
    wp = K.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, device="cuda"), BF, False)
    bias = torch.randn(cout, device="cuda")
    flops = 2.0 * n * d * h * w * cin * cout * 27
    for zs in (0, 1):
        for td in (0, 1, 2, 4, 8):
            for kb in ((0,) if td == 0 else (2, 4)):
                lib.sg_tc_res_zs_mode(zs)
                lib.sg_tc_res_force(td, kb)
                try:
                    t = timeit(lambda: K.conv3d_fprop(x, wp, bias, None, cin, cout, 0.05, True, _lib.IMPL_TCGEN05))
                    tm = timeit(lambda: K.conv3d_fprop(x, wp, None, m, cin, cout, 0.05, False, _lib.IMPL_TCGEN05))
                    print(f"{n}x{cin}->{cout}@{d}x{h}x{w} zs_off={zs} td={td} kb={kb}: {t * 1e3:7.1f} us {flops / t / 1e9:7.1f} TF/s | "
                          f"with mask {tm * 1e3:7.1f} us {flops / tm / 1e9:7.1f} TF/s", flush=True)
                except RuntimeError as e:
                    print(f"{n}x{cin}->{cout} zs_off={zs} td={td} kb={kb}: {str(e)[:80]}")
lib.sg_tc_res_force(0, 0)
lib.sg_tc_res_zs_mode(0)
