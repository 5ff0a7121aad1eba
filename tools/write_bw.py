import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_());            print("memset 1 GiB: %.1f us  %.0f GB/s written" % (ms * 1e3, (1 << 30) / ms / 1e6))
ms = t(lambda: y.copy_(x));           print("copy 1 GiB:   %.1f us  %.0f GB/s read+written" % (ms * 1e3, 2 * (1 << 30) / ms / 1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.sum());             print("read 1 GiB:   %.1f us  %.0f GB/s read" % (ms * 1e3, (1 << 30) / ms / 1e6))
xs = x[: 128 << 20]
ms = t(lambda: xs.zero_());           print("memset 128 MiB: %.1f us  %.0f GB/s written" % (ms * 1e3, (128 << 20) / ms / 1e6))
