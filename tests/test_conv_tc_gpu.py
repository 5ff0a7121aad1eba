"""tcgen05 implicit-GEMM convolution (SG_IMPL_TCGEN05) against the CUDA-core direct kernel
(SG_IMPL_DIRECT) and the torch-CPU restatement, through the C ABI, bf16 in / fp32 accumulate.
Both GPU arms read identical bf16 inputs and packed weights, so they may differ only by fp32
summation order and the final bf16 rounding (tolerance 3e-3 norm-wise)."""
import pytest
import torch

from saragan_b200 import _lib
from saragan_b200 import kernels as K
from tests import cpu_emul as E
from tests.util import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16

SHAPES = [
    # n, cin, cout, d, h, w                     what it exercises
    (1, 16, 16, 4, 16, 8),     # one tile, NT=16, 4 MMA tiles, single K block of 2 chunks
    (1, 32, 64, 4, 16, 16),    # the D.b6 shape class: kb_chunks=4, NT=64, two tiles along w
    (2, 32, 32, 8, 32, 32),    # several tiles in every direction
    (2, 64, 128, 2, 16, 16),   # NT=128, two K blocks (A stage reuse / EMPTY_A handshake)
    (1, 128, 256, 4, 16, 16),  # two N tiles, four K blocks
    (4, 64, 64, 2, 8, 8),      # H=8: MMA tiles straddle halo lines, tiles span samples, split-K
    (3, 256, 128, 2, 8, 8),    # ragged batch vs tn, deep K, split-K workspace
    (2, 16, 8, 4, 16, 16),     # Cout=8 padded to 16
    (1, 24, 48, 2, 16, 8),     # Cin padded (24 -> 32), CoutP=48 -> NT=16 x 3
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("flip", [False, True])
def test_tcgen05_fprop_matches_direct(shape, flip):
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g)
    kin, kout = (cout, cin) if flip else (cin, cout)
    xa = E.plain_to_act(torch.randn(n, kin, d, h, w, generator=g), BF)
    ma = E.plain_to_act(torch.randn(n, kout, d, h, w, generator=g), BF)
    bias = torch.randn(kout, generator=g)
    xg, mg, wp = xa.cuda(), ma.cuda(), K.pack_conv_weight(wt.cuda(), BF, flip)
    for (b, lrelu, mask) in [(None, False, False), (bias, True, False), (bias, False, True)]:
        args = (xg, wp, None if b is None else b.cuda(), mg if mask else None, kin, kout, 0.05, lrelu)
        ref = K.conv3d_fprop(*args, _lib.IMPL_DIRECT)
        got = K.conv3d_fprop(*args, _lib.IMPL_TCGEN05)
        torch.cuda.synchronize()
        e = rel_err(got.float(), ref.float())
        assert e < 3e-3, f"{shape} flip={flip} lrelu={lrelu} mask={mask}: tcgen05 vs direct {e:.3e}"
    want = E.conv3d_fprop(xa, E.pack_conv_weight(wt, BF, flip), bias, None, kin, kout, 0.05, True)
    got = K.conv3d_fprop(xg, wp, bias.cuda(), None, kin, kout, 0.05, True, _lib.IMPL_TCGEN05)
    assert rel_err(got.float().cpu(), want.float()) < 3e-3


# (shape, forced tiling nt/big/td_max/splits): every specialised (compile-time-geometry) streaming kernel, a
# generic one, split-K with vector reductions, tiles spanning two samples
FORCED = [
    ((2, 64, 128, 4, 16, 16), (128, 0, 1, 1)),    # k_conv_tc<128,1,16,1>
    ((2, 64, 128, 4, 16, 16), (128, 0, 2, 2)),    # <128,2,16,1>, split-K 2
    ((1, 64, 128, 8, 16, 16), (128, 1, 4, 1)),    # <128,4,16,1>: 512 TMEM columns
    ((2, 128, 64, 4, 16, 16), (64, 0, 1, 4)),     # <64,1,16,1>, split-K 4
    ((2, 128, 64, 4, 32, 16), (64, 0, 2, 1)),     # <64,2,16,1>
    ((1, 64, 64, 8, 16, 16), (64, 0, 4, 1)),      # <64,4,16,1>
    ((4, 64, 128, 2, 8, 8), (128, 0, 2, 2)),      # <128,2,8,1>: th = 8, MMA tiles straddle halo lines
    ((4, 64, 128, 2, 8, 8), (128, 1, 4, 1)),      # <128,2,8,2>: a tile spans two samples
    ((3, 64, 64, 2, 8, 8), (64, 0, 4, 2)),        # <64,2,8,2>, ragged batch
    ((2, 32, 32, 4, 16, 16), (32, 0, 2, 1)),      # generic kernel (NT = 32)
    ((1, 64, 64, 8, 16, 16), (64, 1, 8, 1)),      # generic kernel: td = 8, 8 MMA tiles per CTA
]


@pytest.mark.parametrize("shape,plan", FORCED)
def test_tcgen05_streaming_forced_tilings(shape, plan):
    """The cost model may pick any of these tilings: each one is forced through the tuning hook and held to the
    CUDA-core kernel (also checks that the tiling the test names is the one that ran)."""
    import ctypes
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 11)
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g).cuda()
    xg = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), BF).cuda()
    mg = E.plain_to_act(torch.randn(n, cout, d, h, w, generator=g), BF).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    wp = K.pack_conv_weight(wt, BF, False)
    lib = _lib.load()
    try:
        lib.sg_tc_force_streaming(1)
        lib.sg_tc_force_plan(*plan)
        out = (ctypes.c_int * 16)()
        lib.sg_tc_plan_debug(n, cin, cout, d, h, w, out)
        assert out[0] == 1 and out[1] == plan[0] and out[8] == plan[3], list(out)
        for (b, lrelu, mask) in [(None, False, False), (bias, True, True)]:
            args = (xg, wp, b, mg if mask else None, cin, cout, 0.05, lrelu)
            got = K.conv3d_fprop(*args, _lib.IMPL_TCGEN05)
            ref = K.conv3d_fprop(*args, _lib.IMPL_DIRECT)
            torch.cuda.synchronize()
            assert rel_err(got.float(), ref.float()) < 3e-3, (shape, plan, lrelu)
    finally:
        lib.sg_tc_force_plan(0, 0, 0, 0)
        lib.sg_tc_force_streaming(0)


RES_SHAPES = [
    (1, 32, 64, 8, 32, 16),    # D.b6.conv2 class: NT=64, td=2, one K block, 16 tiles
    (2, 64, 32, 4, 16, 32),    # its dgrad class: two K blocks per tile through the halo ring
    (2, 32, 32, 8, 16, 16),    # td=4 accumulator sets
    (1, 16, 16, 8, 32, 32),    # kb_chunks=2, 4 halo stages
    (2, 64, 64, 4, 16, 16),    # two N tiles (blockIdx.y)
    (3, 16, 8, 4, 16, 8),      # Cout padded, ragged tile count per CTA
]


@pytest.mark.parametrize("shape", RES_SHAPES)
@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("zs", [0, 1])
def test_tcgen05_resident_kernel_matches_direct(shape, flip, zs):
    """The persistent weight-resident kernel (forced via the test hook so that small shapes take it
    and every CTA walks several tiles) against the direct kernel and the streaming tcgen05 kernel."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 7)
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g)
    kin, kout = (cout, cin) if flip else (cin, cout)
    xg = E.plain_to_act(torch.randn(n, kin, d, h, w, generator=g), BF).cuda()
    mg = E.plain_to_act(torch.randn(n, kout, d, h, w, generator=g), BF).cuda()
    bias = torch.randn(kout, generator=g).cuda()
    wp = K.pack_conv_weight(wt.cuda(), BF, flip)
    lib = _lib.load()
    lib.sg_tc_res_zs_mode(zs)   # 0: z-stacked accumulators where NT <= 32 allows; 1: one accumulator per output plane
    try:
        for (b, lrelu, mask) in [(None, False, False), (bias, True, True)]:
            args = (xg, wp, b, mg if mask else None, kin, kout, 0.05, lrelu)
            ref = K.conv3d_fprop(*args, _lib.IMPL_DIRECT)
            lib.sg_tc_force_streaming(2)
            res = K.conv3d_fprop(*args, _lib.IMPL_TCGEN05)
            lib.sg_tc_force_streaming(1)
            stream = K.conv3d_fprop(*args, _lib.IMPL_TCGEN05)
            torch.cuda.synchronize()
            assert rel_err(res.float(), ref.float()) < 3e-3, (shape, flip, "resident vs direct")
            assert rel_err(res.float(), stream.float()) < 3e-3, (shape, flip, "resident vs streaming")
    finally:
        lib.sg_tc_force_streaming(0)
        lib.sg_tc_res_zs_mode(0)


WGRAD_SHAPES = [
    (1, 16, 16, 2, 16, 8),     # one tile, NT=16, 2 gy chunks (14 garbage row-chunks)
    (1, 32, 64, 4, 16, 16),    # D.b6 class: NT=32, td=2, several tiles, 3-stage ring wraps
    (2, 64, 128, 2, 16, 16),   # two ci tiles, full M=128
    (1, 128, 256, 4, 16, 16),  # two co tiles x four ci tiles
    (4, 64, 64, 2, 8, 8),      # H=8 planes
    (3, 256, 128, 2, 8, 8),    # ragged batch, many channel tiles
    (2, 16, 8, 4, 16, 16),     # Cout=8 padded
    (1, 24, 48, 2, 16, 8),     # padded Cin/Cout
    (2, 32, 32, 8, 32, 32),    # persistent CTAs loop over many tiles; all three kd stacked along M
    (2, 32, 32, 1, 16, 16),    # D=1: the kd = 0, 2 row blocks only ever see out-of-volume gy planes
    (1, 48, 80, 2, 16, 16),    # 10 gy chunks: no room for a second kd block
    (8, 32, 16, 4, 32, 32),    # Cout=16: eight row blocks fit, three are live
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_tcgen05_wgrad_matches_direct(shape):
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    xa = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), BF)
    ga = E.plain_to_act(torch.randn(n, cout, d, h, w, generator=g), BF)
    xg, gg = xa.cuda(), ga.cuda()
    ref_w, ref_b = K.conv3d_wgrad(xg, gg, cin, cout, 0.07, True, _lib.IMPL_DIRECT)
    got_w, got_b = K.conv3d_wgrad(xg, gg, cin, cout, 0.07, True, _lib.IMPL_TCGEN05)
    torch.cuda.synchronize()
    e = rel_err(got_w, ref_w)
    assert e < 1e-3, f"{shape}: tcgen05 wgrad vs direct {e:.3e}"
    assert rel_err(got_b, ref_b) < 1e-4
    want_w, _ = E.conv3d_wgrad(xa, ga, cin, cout, 0.07, False)
    assert rel_err(got_w.cpu(), want_w) < 1e-3


def test_tcgen05_wgrad_linearity_full_size():
    """Size-independent property at the cfg3 top-level shape (32->64 @32x128x128):
    <wgrad(x, g), w> == <conv(x, w), g> ties the tcgen05 wgrad to the tcgen05 fprop."""
    n, cin, cout, d, h, w = 1, 32, 64, 32, 128, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g).bfloat16().float()
    x = K.plain_to_act(torch.randn(n, cin, d, h, w, device="cuda", generator=g), BF)
    gy = K.plain_to_act(torch.randn(n, cout, d, h, w, device="cuda", generator=g), BF)
    gw, _ = K.conv3d_wgrad(x, gy, cin, cout, 1.0, False, _lib.IMPL_TCGEN05)
    y = K.conv3d_fprop(x, K.pack_conv_weight(wt, BF, False), None, None, cin, cout, 1.0, False,
                       _lib.IMPL_TCGEN05)
    lhs = float((gw.double() * wt.double()).sum())
    rhs = float((y.double() * gy.double()).sum())
    assert abs(lhs - rhs) < 5e-3 * max(abs(lhs), abs(rhs)), (lhs, rhs)


def test_tcgen05_pad_channels_are_zero():
    """pad channels of the output chunk must be written as zeros (the next layer contracts over them)"""
    n, cin, cout, d, h, w = 1, 16, 8, 4, 16, 8
    wt = torch.randn(cout, cin, 3, 3, 3)
    x = E.plain_to_act(torch.randn(n, cin, d, h, w), BF).cuda()
    y = K.conv3d_fprop(x, K.pack_conv_weight(wt.cuda(), BF, False), torch.ones(cout).cuda(), None, cin, cout,
                       1.0, False, _lib.IMPL_TCGEN05)
    assert float(y[:, 1].abs().max()) == 0.0


def test_tcgen05_adjoint_property_full_size():
    """Size-independent property at the BASELINE cfg3 top-level shape (32->64 @32x128x128):
    <conv(x), g> == <x, dgrad(g)> for the tcgen05 fprop and its flipped-packing dgrad."""
    n, cin, cout, d, h, w = 1, 32, 64, 32, 128, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda", generator=g)
    x = K.plain_to_act(torch.randn(n, cin, d, h, w, device="cuda", generator=g), BF)
    gy = K.plain_to_act(torch.randn(n, cout, d, h, w, device="cuda", generator=g), BF)
    y = K.conv3d_fprop(x, K.pack_conv_weight(wt, BF, False), None, None, cin, cout, 0.03, False, _lib.IMPL_TCGEN05)
    gx = K.conv3d_fprop(gy, K.pack_conv_weight(wt, BF, True), None, None, cout, cin, 0.03, False, _lib.IMPL_TCGEN05)
    lhs = float((y.double() * gy.double()).sum())
    rhs = float((x.double() * gx.double()).sum())
    assert abs(lhs - rhs) < 5e-3 * max(abs(lhs), abs(rhs)), (lhs, rhs)


PN_SHAPES = [(2, 32, 16, 8, 32, 16), (1, 64, 32, 4, 16, 32), (2, 16, 16, 8, 16, 16), (1, 32, 64, 4, 16, 16), (3, 16, 8, 4, 16, 8)]


@pytest.mark.parametrize("shape", PN_SHAPES)
@pytest.mark.parametrize("lrelu,lrelu_after", [(True, False), (False, True)])
def test_fused_conv_pixelnorm(shape, lrelu, lrelu_after):
    """sg_conv3d_fprop_pixelnorm (conv + ChannelNormalization in one tcgen05 kernel: the generator block's conv1 ->
    lrelu -> pixel-norm and conv2 -> pixel-norm -> lrelu) against the CPU restatement and against the two separate
    kernels (which round y to bf16 before normalising it: 1e-2)."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 21)
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g)
    xa = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), BF)
    bias = torch.randn(cout, generator=g)
    xg, wp = xa.cuda(), K.pack_conv_weight(wt.cuda(), BF, False)
    lib = _lib.load()
    try:
        lib.sg_tc_force_streaming(2)      # take the weight-resident kernel although the test shape has few tiles
        assert K.conv_pixelnorm_supported(xg, cin, cout)
        y, yn = K.conv3d_fprop_pixelnorm(xg, wp, bias.cuda(), cin, cout, 0.05, lrelu, lrelu_after)
        y_sep = K.conv3d_fprop(xg, wp, bias.cuda(), None, cin, cout, 0.05, lrelu, _lib.IMPL_TCGEN05)
        yn_sep = K.pixelnorm_fwd(y_sep, cout, lrelu_after)
        torch.cuda.synchronize()
    finally:
        lib.sg_tc_force_streaming(0)
    want_y, want_yn = E.conv3d_fprop_pixelnorm(xa, E.pack_conv_weight(wt, BF, False), bias, cin, cout, 0.05, lrelu, lrelu_after)
    assert rel_err(y.float().cpu(), want_y.float()) < 3e-3 and rel_err(yn.float().cpu(), want_yn.float()) < 3e-3
    assert torch.equal(y, y_sep)                                     # the conv output itself is the same kernel's
    assert rel_err(yn.float(), yn_sep.float()) < 1e-2
    if cout % 16:                                                    # pad channels stay exactly zero
        cc = y.shape[1]
        flat = yn.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cc * 8, d, h, w)
        assert float(flat[:, cout:].abs().max()) == 0.0


def test_generator_block_fused_equals_unfused(monkeypatch):
    """GeneratorBlock with the fused conv + pixel-norm kernels against the same block on the separate kernels: outputs
    and every gradient (bf16 rounding of the intermediate is the only difference)."""
    import saragan_b200 as sg
    from saragan_b200 import kernels
    torch.manual_seed(3)
    blk = sg.GeneratorBlock(64, 32).cuda()
    x = torch.randn(2, 64, 4, 16, 8, device="cuda")
    gy = torch.randn(2, 32, 8, 32, 16, device="cuda")
    lib = _lib.load()
    res = {}
    for fused in (True, False):
        if not fused:
            monkeypatch.setattr(kernels, "conv_pixelnorm_supported", lambda *a: False)
        try:
            lib.sg_tc_force_streaming(2)
            for rep in range(2):           # the first pass also packs the weights: count the second
                n0 = _lib.launch_count()
                xi = x.clone().requires_grad_(True)
                y = blk(xi)
                grads = torch.autograd.grad(y, [xi] + list(blk.parameters()), gy)
                torch.cuda.synchronize()
            res[fused] = (y, grads, _lib.launch_count() - n0)
        finally:
            lib.sg_tc_force_streaming(0)
    assert res[True][2] < res[False][2]                               # two kernels fewer in the forward
    assert rel_err(res[True][0], res[False][0]) < 1e-2
    for a, b in zip(res[True][1], res[False][1]):
        assert rel_err(a, b) < 5e-2, rel_err(a, b)       # (the unfused arm normalises the bf16-rounded conv output)


POOL_SHAPES = [(2, 32, 64, 8, 32, 16), (1, 16, 32, 4, 16, 32), (2, 16, 16, 8, 16, 16), (3, 32, 32, 4, 16, 8)]


@pytest.mark.parametrize("shape", POOL_SHAPES)
def test_fused_conv_avgpool(shape):
    """sg_conv3d_fprop_pool (the discriminator block's conv2 -> lrelu -> AvgPool3d(2) in one tcgen05 kernel) against the
    CPU restatement and the two separate kernels (which pool the bf16-rounded y)."""
    n, cin, cout, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 31)
    wt = torch.randn(cout, cin, 3, 3, 3, generator=g)
    xa = E.plain_to_act(torch.randn(n, cin, d, h, w, generator=g), BF)
    bias = torch.randn(cout, generator=g)
    xg, wp = xa.cuda(), K.pack_conv_weight(wt.cuda(), BF, False)
    lib = _lib.load()
    try:
        lib.sg_tc_force_streaming(2)
        assert K.conv_pool_supported(xg, cin, cout)
        y, yp = K.conv3d_fprop_pool(xg, wp, bias.cuda(), cin, cout, 0.05, True, 0.125)
        y_sep = K.conv3d_fprop(xg, wp, bias.cuda(), None, cin, cout, 0.05, True, _lib.IMPL_TCGEN05)
        yp_sep = K.down2(y_sep, 0.125)
        torch.cuda.synchronize()
    finally:
        lib.sg_tc_force_streaming(0)
    want_y, want_yp = E.conv3d_fprop_pool(xa, E.pack_conv_weight(wt, BF, False), bias, cin, cout, 0.05, True, 0.125)
    assert torch.equal(y, y_sep)
    assert rel_err(y.float().cpu(), want_y.float()) < 3e-3 and rel_err(yp.float().cpu(), want_yp.float()) < 3e-3
    assert rel_err(yp.float(), yp_sep.float()) < 6e-3


def test_discriminator_block_fused_pool_equals_unfused(monkeypatch):
    """DiscriminatorBlock with the fused conv2 + avg-pool kernel against the same block on the separate kernels: output,
    first-order gradients and the double backward of a gradient-penalty-like scalar."""
    import saragan_b200 as sg
    from saragan_b200 import kernels
    torch.manual_seed(4)
    blk = sg.DiscriminatorBlock(16, 32).cuda()
    x = torch.randn(2, 16, 16, 32, 32, device="cuda")     # pooled level 8x16x16 > config.TF32_MAX_VOXELS: bf16 on both sides
    lib = _lib.load()
    res = {}
    for fused in (True, False):
        if not fused:
            monkeypatch.setattr(kernels, "conv_pool_supported", lambda *a: False)
        try:
            lib.sg_tc_force_streaming(2)
            for rep in range(2):           # the first pass also packs the weights: count the second
                n0 = _lib.launch_count()
                xi = x.clone().requires_grad_(True)
                y = blk(xi)
                (gx,) = torch.autograd.grad(y.float().pow(2).sum(), xi, create_graph=True)
                gp = (gx.flatten(1).norm(dim=1) - 1).pow(2).mean()
                gw = torch.autograd.grad(gp, list(blk.parameters()))
                torch.cuda.synchronize()
            res[fused] = (y, gx, gw, _lib.launch_count() - n0)
        finally:
            lib.sg_tc_force_streaming(0)
    assert res[True][3] < res[False][3]
    assert rel_err(res[True][0], res[False][0]) < 6e-3
    assert rel_err(res[True][1], res[False][1]) < 2e-2
    for a, b in zip(res[True][2], res[False][2]):
        assert rel_err(a, b) < 5e-2, rel_err(a, b)
