#!/usr/bin/env python
"""Launch the tcgen05 conv kernels alone on one layer shape (for ncu captures and CUDA-event
timing).  python tools/run_conv.py --n 4 --cin 32 --cout 64 --vol 32 128 128 --reps 5"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saragan_b200 import _lib, kernels as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--cin", type=int, default=32)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--vol", type=int, nargs=3, default=[32, 128, 128])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--what", default="fprop,wgrad")
    args = ap.parse_args()
    d, h, w = args.vol
    BF = torch.bfloat16
    x = K.plain_to_act(torch.randn(args.n, args.cin, d, h, w, device="cuda"), BF)
    gy = K.plain_to_act(torch.randn(args.n, args.cout, d, h, w, device="cuda"), BF)
    wt = torch.randn(args.cout, args.cin, 3, 3, 3, device="cuda")
    wp = K.pack_conv_weight(wt, BF, False)
    bias = torch.randn(args.cout, device="cuda")
    flops = 2.0 * args.n * d * h * w * args.cin * args.cout * 27
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for what in args.what.split(","):
        times = []
        for i in range(args.reps + 2):
            flush.zero_()                                    # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if what == "fprop":
                K.conv3d_fprop(x, wp, bias, None, args.cin, args.cout, 0.05, True, _lib.IMPL_TCGEN05)
            else:
                K.conv3d_wgrad(x, gy, args.cin, args.cout, 0.05, False, _lib.IMPL_TCGEN05)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
        t = sum(times) / len(times)
        print(f"{what} n={args.n} {args.cin}->{args.cout} @{d}x{h}x{w}: {t * 1e3:.1f} us  {flops / t / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
