#!/usr/bin/env python
"""Time the streaming tcgen05 fprop kernel under every candidate tiling (sg_tc_force_plan) for the layer
shapes whose weights are not shared-memory resident, next to the automatic choice: the data the cost model
in conv_tc.cu (make_plan_cfg) is fitted to.  python tools/plan_sweep.py [--shapes "4,128,128,8,32,32;..."]"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import _lib, kernels as K  # noqa: E402

DEFAULT = ("4,512,512,2,8,8;8,512,512,2,8,8;4,256,256,4,16,16;8,256,256,4,16,16;4,512,256,4,16,16;4,256,512,4,16,16;"
           "4,128,128,8,32,32;8,128,128,8,32,32;4,256,128,8,32,32;4,128,256,8,32,32;4,64,64,16,64,64;4,128,64,16,64,64;"
           "4,64,128,16,64,64;8,64,128,16,64,64")


def timed(fn, flush, reps=4):
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default=DEFAULT)
    ap.add_argument("--top", type=int, default=6)
    args = ap.parse_args()
    lib = _lib.load()
    lib.sg_tc_force_streaming(1)
    BF = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = (ctypes.c_int * 16)()
    for spec in args.shapes.split(";"):
        n, cin, cout, d, h, w = (int(v) for v in spec.split(","))
        x = K.plain_to_act(torch.randn(n, cin, d, h, w, device="cuda"), BF)
        wp = K.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, device="cuda"), BF, False)
        bias = torch.randn(cout, device="cuda")
        run = lambda: K.conv3d_fprop(x, wp, bias, None, cin, cout, 0.05, True, _lib.IMPL_TCGEN05)  # noqa: E731
        flops = 2.0 * n * d * h * w * cin * cout * 27
        lib.sg_tc_force_plan(0, 0, 0, 0)
        lib.sg_tc_plan_debug(n, cin, cout, d, h, w, out)
        auto = list(out)
        t_auto = timed(run, flush)
        rows = []
        seen = set()
        for nt in (128, 64, 32):
            for big in (0, 1):
                for td in (8, 4, 2, 1):
                    for splits in (1, 2, 4, 8, 16, 32):
                        lib.sg_tc_force_plan(nt, big, td, splits)
                        lib.sg_tc_plan_debug(n, cin, cout, d, h, w, out)
                        v = list(out)
                        if not v[0]:
                            continue
                        key = tuple(v[1:16])
                        if key in seen or v[10] * v[11] * v[12] > 8 * 148:
                            continue
                        seen.add(key)
                        rows.append((timed(run, flush, 3), nt, big, v))
        lib.sg_tc_force_plan(0, 0, 0, 0)
        rows.sort(key=lambda r: r[0])
        print(f"== n={n} {cin}->{cout} @{d}x{h}x{w}: auto {t_auto:.1f} us ({flops / t_auto / 1e6:.0f} TF/s)  plan NT={auto[1]} tn={auto[2]} "
              f"td={auto[3]} n_sub={auto[5]} kb={auto[6] % 100} tps={auto[6] // 100 % 100} spec={auto[6] // 10000} sw={auto[7]} splits={auto[8]} grid={auto[10]}x{auto[11]}x{auto[12]} smem={auto[13]}")
        for t, nt, big, v in rows[:args.top]:
            print(f"   {t:7.1f} us  NT={nt} big={big} tn={v[2]} td={v[3]} n_sub={v[5]} kb={v[6] % 100} tps={v[6] // 100 % 100} spec={v[6] // 10000} sw={v[7]} splits={v[8]} "
                  f"grid={v[10]}x{v[11]}x{v[12]} smem={v[13]}")
    lib.sg_tc_force_streaming(0)


if __name__ == "__main__":
    main()
