"""tcgen05 kind::tf32 convolution kernels (SG_IMPL_TF32: fp32 activations read through half-chunk tensor maps, weights in
the SG_TF32 packing, operands rounded to nearest to 10 mantissa bits in the kernel, fp32 accumulation) through the C ABI.

Two kinds of check per shape:
 * operands already representable in tf32: the kernel must agree with the exact fp32 CUDA-core kernel on the same values
   to accumulation-order noise (1e-5) -- pins the data path (tensor maps, descriptors, epilogue, tap geometry);
 * arbitrary fp32 operands: against the CPU restatement that rounds to nearest (tests/cpu_emul.py, impl 3) to 1e-5 --
   pins the in-kernel rounding (a kernel that let the tensor core truncate would sit at ~5e-4)."""
import pytest
import torch

from saragan_b200 import _lib
from saragan_b200 import kernels as K
from tests import cpu_emul as E
from tests.util import rel_err

pytestmark = pytest.mark.gpu
F32 = torch.float32

SHAPES = [
    # n, cin, cout, d, h, w
    (1, 16, 16, 4, 16, 8),
    (1, 32, 64, 4, 16, 16),
    (2, 32, 32, 8, 32, 32),
    (2, 64, 128, 2, 16, 16),
    (1, 128, 256, 4, 16, 16),
    (4, 64, 64, 2, 8, 8),      # H = 8: MMA tiles straddle halo lines (no sample-spanning tiles in the tf32 kernel)
    (3, 256, 128, 2, 8, 8),    # deep K, split-K workspace
    (2, 16, 8, 4, 16, 16),     # Cout = 8 padded to 16
    (1, 24, 48, 2, 16, 8),     # Cin padded (24 -> 32), CoutP = 48
    (4, 512, 512, 2, 8, 8),    # the cfg3 low-resolution shape class
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("flip", [False, True])
def test_tf32_fprop_and_dgrad(shape, flip):
    n, cin, cout, d, h, w = shape
    assert K.conv_tf32_supported(n, cin, cout, d, h, w)
    g = torch.Generator().manual_seed(sum(shape))
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g)
    kin, kout = (cout, cin) if flip else (cin, cout)
    x = torch.randn(n, kin, d, h, w, generator=g)
    ma = E.plain_to_act(torch.randn(n, kout, d, h, w, generator=g), F32)
    bias = torch.randn(kout, generator=g)
    # (a) tf32-representable operands against the exact fp32 kernel
    xr, wr = E._tf32(x), E._tf32(wt)
    xg, mg = E.plain_to_act(xr, F32).cuda(), ma.cuda()
    wp_tf32, wp_f32 = K.pack_conv_weight(wr.cuda(), "tf32", flip), K.pack_conv_weight(wr.cuda(), F32, flip)
    for (b, lrelu, mask) in [(None, False, False), (bias, True, False), (bias, False, True)]:
        tail = (None if b is None else b.cuda(), mg if mask else None, kin, kout, 0.05, lrelu)
        ref = K.conv3d_fprop(xg, wp_f32, *tail, _lib.IMPL_DIRECT)
        got = K.conv3d_fprop(xg, wp_tf32, *tail, _lib.IMPL_TF32)
        torch.cuda.synchronize()
        e = rel_err(got, ref)
        assert e < 1e-5, f"{shape} flip={flip} lrelu={lrelu} mask={mask}: tf32 kernel vs fp32 kernel on tf32 operands {e:.3e}"
    # (b) arbitrary operands against the round-to-nearest restatement
    xa = E.plain_to_act(x, F32)
    want = E.conv3d_fprop(xa, E.pack_conv_weight(wt, "tf32", flip), bias, None, kin, kout, 0.05, True, 3)
    got = K.conv3d_fprop(xa.cuda(), K.pack_conv_weight(wt.cuda(), "tf32", flip), bias.cuda(), None, kin, kout, 0.05, True,
                         _lib.IMPL_TF32)
    e = rel_err(got.cpu(), want)
    assert e < 1e-5, f"{shape} flip={flip}: in-kernel rounding {e:.3e}"
    exact = E.conv3d_fprop(xa, E.pack_conv_weight(wt, F32, flip), bias, None, kin, kout, 0.05, True)
    assert 2e-5 < rel_err(got.cpu(), exact) < 1e-3      # it IS tf32 arithmetic, and within the tier's tolerance


FORCED = [
    ((2, 64, 128, 4, 16, 16), (128, 0, 1, 1)),    # k_conv_tc<128,1,16,1,tf32>
    ((2, 64, 128, 4, 16, 16), (128, 0, 2, 2)),    # <128,2,16,1>, split-K 2
    ((1, 64, 128, 8, 16, 16), (128, 1, 4, 1)),    # <128,4,16,1>
    ((2, 128, 64, 4, 16, 16), (64, 0, 1, 4)),     # <64,1,16,1>, split-K 4
    ((2, 128, 64, 4, 32, 16), (64, 0, 2, 1)),     # <64,2,16,1>
    ((1, 64, 64, 8, 16, 16), (64, 0, 4, 1)),      # <64,4,16,1>
    ((4, 64, 128, 2, 8, 8), (128, 0, 2, 2)),      # <128,2,8,1>
    ((3, 64, 64, 2, 8, 8), (64, 0, 2, 2)),        # <64,2,8,1>
    ((2, 32, 32, 4, 16, 16), (32, 0, 2, 1)),      # generic kernel (NT = 32)
    ((1, 64, 64, 8, 16, 16), (64, 1, 8, 1)),      # generic kernel: td = 8
]


@pytest.mark.parametrize("shape,plan", FORCED)
def test_tf32_forced_tilings(shape, plan):
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 11)
    wr = E._tf32(torch.randn(cout, cin, 3, 3, 3, generator=g)).cuda()
    xg = E.plain_to_act(E._tf32(torch.randn(n, cin, d, h, w, generator=g)), F32).cuda()
    mg = E.plain_to_act(torch.randn(n, cout, d, h, w, generator=g), F32).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    lib = _lib.load()
    try:
        lib.sg_tc_force_plan(*plan)
        for (b, lrelu, mask) in [(None, False, False), (bias, True, True)]:
            tail = (b, mg if mask else None, cin, cout, 0.05, lrelu)
            got = K.conv3d_fprop(xg, K.pack_conv_weight(wr, "tf32", False), *tail, _lib.IMPL_TF32)
            ref = K.conv3d_fprop(xg, K.pack_conv_weight(wr, F32, False), *tail, _lib.IMPL_DIRECT)
            torch.cuda.synchronize()
            assert rel_err(got, ref) < 1e-5, (shape, plan, lrelu, rel_err(got, ref))
    finally:
        lib.sg_tc_force_plan(0, 0, 0, 0)


WGRAD_SHAPES = [
    (1, 16, 16, 4, 16, 8),      # Cout <= 32: all three kd stacked along M, NT = 16
    (2, 32, 32, 8, 16, 16),     # NT = 32
    (1, 32, 64, 4, 16, 16),     # Cout <= 64: two kd groups
    (2, 64, 128, 2, 16, 16),    # one kd per CTA group, two ci tiles
    (1, 128, 256, 4, 16, 16),   # two co tiles
    (4, 64, 64, 2, 8, 8),       # th = 8
    (2, 24, 48, 2, 16, 8),      # padded channels
    (4, 512, 512, 2, 8, 8),     # the cfg3 low-resolution shape class
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_fp32_level_wgrad_split_bf16(shape):
    """wgrad of the fp32 / TF32 levels (SG_IMPL_TF32 on fp32 tensors): tcgen05 kind::tf32 has no MN-major operands
    (tools/tf32_mn_probe.cu), so the kernel splits the fp32 tiles into bf16 hi + lo halves in shared memory and forms
    the product from three bf16 MMAs -- 16 mantissa bits per operand.  Against the exact fp32 CUDA-core kernel on the
    same operands: 5e-5 (TF32 would sit at 3e-4); against the CPU restatement of the same split: 1e-5."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 3)
    x, gy = torch.randn(n, cin, d, h, w, generator=g), torch.randn(n, cout, d, h, w, generator=g)
    xa, ga = E.plain_to_act(x, F32), E.plain_to_act(gy, F32)
    xg, gg = xa.cuda(), ga.cuda()
    ref_w, ref_b = K.conv3d_wgrad(xg, gg, cin, cout, 0.05, True, _lib.IMPL_DIRECT)
    f0 = int(_lib.load().sg_cuda_core_fallbacks(0))
    got_w, got_b = K.conv3d_wgrad(xg, gg, cin, cout, 0.05, True, _lib.IMPL_TF32)
    torch.cuda.synchronize()
    assert int(_lib.load().sg_cuda_core_fallbacks(0)) == f0, "the tensor-core wgrad declined this shape"
    assert rel_err(got_w, ref_w) < 5e-5, (shape, rel_err(got_w, ref_w))
    assert rel_err(got_b, ref_b) < 5e-5, (shape, rel_err(got_b, ref_b))
    want_w, want_b = E.conv3d_wgrad(xa, ga, cin, cout, 0.05, True, 3)
    assert rel_err(got_w.cpu(), want_w) < 1e-5, (shape, rel_err(got_w.cpu(), want_w))
    assert rel_err(got_b.cpu(), want_b) < 1e-5, (shape, rel_err(got_b.cpu(), want_b))


@pytest.mark.parametrize("shape", WGRAD_SHAPES[1:6])
def test_fp32_level_wgrad_bf16_operands(shape):
    """SG_IMPL_F32_AS_BF16: the same kernel forming only the hi halves (what the bf16 policy uses for the weight
    gradients of its fp32-storage levels) against the restatement that rounds both operands to bf16."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 5)
    xa = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), F32)
    ga = E.plain_to_act(torch.randn(n, cout, d, h, w, generator=g), F32)
    f0 = int(_lib.load().sg_cuda_core_fallbacks(0))
    got_w, got_b = K.conv3d_wgrad(xa.cuda(), ga.cuda(), cin, cout, 0.05, True, _lib.IMPL_F32_AS_BF16)
    torch.cuda.synchronize()
    assert int(_lib.load().sg_cuda_core_fallbacks(0)) == f0
    want_w, want_b = E.conv3d_wgrad(xa, ga, cin, cout, 0.05, True, 4)
    assert rel_err(got_w.cpu(), want_w) < 1e-5 and rel_err(got_b.cpu(), want_b) < 1e-5, shape
    exact_w, _ = E.conv3d_wgrad(xa, ga, cin, cout, 0.05, True)
    assert rel_err(got_w.cpu(), exact_w) < 6e-3


def test_tf32_packing_layout():
    """SG_TF32 packing = [tap][K/4][rows][4] fp32, values rounded to nearest tf32, pad rows / channels zero."""
    cout, cin = 24, 20
    wt = torch.randn(cout, cin, 3, 3, 3, generator=torch.Generator().manual_seed(1))
    for flip in (False, True):
        got = K.pack_conv_weight(wt.cuda(), "tf32", flip).cpu()
        k, r = (cout, cin) if flip else (cin, cout)
        kc4, rp = 4 * ((k + 15) // 16), 16 * ((r + 15) // 16)
        want = torch.zeros(27, kc4, rp, 4)
        src = E._tf32(wt).reshape(cout, cin, 27)
        for tap in range(27):
            m = src[:, :, 26 - tap].t() if flip else src[:, :, tap]        # [row][k]
            pad = torch.zeros(rp, kc4 * 4)
            pad[:r, :k] = m if not flip else m
            want[tap] = pad.reshape(rp, kc4, 4).permute(1, 0, 2)
        assert torch.equal(got.reshape(27, kc4, rp, 4), want), flip


@pytest.mark.parametrize("shape", [(4, 512, 512, 2, 8, 8), (4, 256, 256, 4, 16, 16), (4, 64, 64, 2, 8, 8)])
@pytest.mark.parametrize("kind", ["tf32", "bf16"])
def test_split_k_is_run_to_run_deterministic(shape, kind):
    """The split-K shapes of the low-resolution levels: every K slice stores to its own workspace slab and the finishing
    kernel adds the slabs in a fixed order, so repeated launches are bit-identical (with fp32 atomics they differed in
    the last bit, which the reduced-precision step amplified to 8e-2 on D's gradients: tools/determinism_probe.py)."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    dt, pk, impl = (F32, "tf32", _lib.IMPL_TF32) if kind == "tf32" else (torch.bfloat16, torch.bfloat16, _lib.IMPL_AUTO)
    x = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), dt).cuda()
    wp = K.pack_conv_weight(torch.randn(cout, cin, 3, 3, 3, generator=g).cuda(), pk, False)
    bias = torch.randn(cout, generator=g).cuda()
    first = None
    for _ in range(6):
        y = K.conv3d_fprop(x, wp, bias, None, cin, cout, 0.05, True, impl)
        torch.cuda.synchronize()
        if first is None:
            first = y.clone()
        assert torch.equal(first, y)
