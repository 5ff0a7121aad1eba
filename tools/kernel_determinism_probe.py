"""Kernel-level run-to-run determinism (one GPU): each conv entry point is called REPS times on the same operands and
every result is compared bit-for-bit with the first."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from saragan_b200 import _lib, kernels as K
from tests import cpu_emul as E
REPS = 30
F32, BF16 = torch.float32, torch.bfloat16
SHAPES = [(4, 64, 64, 1, 4, 4), (4, 64, 64, 2, 8, 8), (4, 64, 32, 4, 16, 16), (4, 32, 32, 4, 16, 16), (4, 32, 16, 4, 16, 16),
          (4, 16, 16, 4, 16, 16), (4, 512, 512, 2, 8, 8), (4, 256, 256, 4, 16, 16), (4, 128, 128, 8, 32, 32), (2, 32, 32, 32, 128, 128)]
for shape in SHAPES:
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    for tag, dt, kind, impl in (("tf32", F32, "tf32", _lib.IMPL_TF32), ("bf16", BF16, BF16, _lib.IMPL_AUTO)):
        if tag == "tf32" and not K.conv_tf32_supported(n, cin, cout, d, h, w):
            continue
        for flip in (False, True):
            kin, kout = (cout, cin) if flip else (cin, cout)
            x = E.plain_to_act(torch.randn(n, kin, d, h, w, generator=g), dt).cuda()
            wp = K.pack_conv_weight(wt, kind, flip)
            first, bad, worst = None, 0, 0.0
            for r in range(REPS):
                y = K.conv3d_fprop(x, wp, None if flip else bias, None, kin, kout, 0.05, not flip, impl)
                torch.cuda.synchronize()
                if first is None:
                    first = y.clone()
                elif not torch.equal(first, y):
                    bad += 1
                    worst = max(worst, float((first.float() - y.float()).abs().max() / first.float().abs().max()))
            print(f"{tag} {'dgrad' if flip else 'fprop'} {shape}: {bad}/{REPS - 1} repeats differ, worst {worst:.2e}", flush=True)
