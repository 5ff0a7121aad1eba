"""Per-entry-point parity of libsaragan_b200.so on the GPU (through the C ABI) against the
torch-CPU restatement of each entry point in tests/cpu_emul.py, on identical inputs.

bf16 inputs are rounded to bf16 before both arms see them, so the only differences are the
accumulation order (fp32 in both) and the final bf16 rounding of the output: tolerance
3e-3 norm-wise for bf16 outputs, 2e-5 for fp32 outputs.
"""
import pytest
import torch

from saragan_b200 import _lib
from saragan_b200 import kernels as K
from tests import cpu_emul as E
from tests.util import rel_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32]


def tol(dtype):
    return 3e-3 if dtype == torch.bfloat16 else 2e-5


def rnd(*shape, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype).float()


def act(n, c, d, h, w, dtype, seed=0):
    """random blocked activation (CPU copy, CUDA copy) with zero pad channels"""
    a = E.plain_to_act(rnd(n, c, d, h, w, seed=seed), dtype)
    return a, a.cuda()


def close(got, want, dtype, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    e = rel_err(got.float().cpu(), want.float())
    assert e < tol(dtype), f"{what}: rel err {e:.3e}"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [8, 16, 40])
def test_layout_roundtrip(dtype, c):
    plain = rnd(3, c, 2, 4, 8)
    a = K.plain_to_act(plain.cuda(), dtype)
    close(a, E.plain_to_act(plain, dtype), dtype, "plain_to_act")
    back = K.act_to_plain(a, c)
    close(back, plain.to(dtype).float(), dtype, "act_to_plain")


CONV_SHAPES = [
    # n, cin, cout, d, h, w
    (2, 16, 32, 4, 8, 8),
    (1, 32, 16, 2, 16, 16),
    (2, 24, 8, 3, 5, 7),      # padded channel counts, odd extents
    (4, 33, 32, 1, 4, 4),     # base level with the mbstd channel (small-volume fp32 path)
    (3, 64, 48, 2, 8, 8),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("flip", [False, True])
def test_conv_fprop_direct(dtype, shape, flip):
    n, cin, cout, d, h, w = shape
    wt = rnd(cout, cin, 3, 3, 3, seed=1)
    kin, kout = (cout, cin) if flip else (cin, cout)      # dgrad contracts over Cout
    xa, xg = act(n, kin, d, h, w, dtype, seed=2)
    bias = rnd(kout, seed=3)
    ma, mg = act(n, kout, d, h, w, dtype, seed=4)
    for (b, lrelu, mask, impl) in [(None, False, False, _lib.IMPL_DIRECT), (bias, True, False, _lib.IMPL_DIRECT),
                                   (bias, False, True, _lib.IMPL_DIRECT), (bias, True, False, _lib.IMPL_AUTO)]:
        want = E.conv3d_fprop(xa, E.pack_conv_weight(wt, dtype, flip), b, ma if mask else None, kin, kout,
                              0.37, lrelu)
        wp = K.pack_conv_weight(wt.cuda(), dtype, flip)
        got = K.conv3d_fprop(xg, wp, None if b is None else b.cuda(), mg if mask else None, kin, kout,
                             0.37, lrelu, impl)
        close(got, want, dtype, f"conv fprop flip={flip} lrelu={lrelu} mask={mask} impl={impl}")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_wgrad_direct(dtype, shape):
    n, cin, cout, d, h, w = shape
    xa, xg = act(n, cin, d, h, w, dtype, seed=5)
    ga, gg = act(n, cout, d, h, w, dtype, seed=6)
    want_w, want_b = E.conv3d_wgrad(xa, ga, cin, cout, 0.21, True)
    got_w, got_b = K.conv3d_wgrad(xg, gg, cin, cout, 0.21, True, _lib.IMPL_DIRECT)
    close(got_w, want_w, torch.float32, "wgrad")
    close(got_b, want_b, torch.float32, "bias grad")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [8, 16, 33])
def test_pointwise_rgb(dtype, c):
    n, d, h, w = 3, 2, 8, 8
    img = rnd(n, 1, d, h, w, seed=7)
    wv, bv = rnd(c, seed=8), rnd(c, seed=9)
    got = K.pw_expand(img.cuda(), wv.cuda(), bv.cuda(), dtype, c, 0.5, True)
    close(got, E.pw_expand(img, wv, bv, dtype, c, 0.5, True), dtype, "pw_expand")
    xa, xg = act(n, c, d, h, w, dtype, seed=10)
    b1 = rnd(1, seed=11)
    close(K.pw_reduce(xg, wv.cuda(), b1.cuda(), c, 0.5), E.pw_reduce(xa, wv, b1, c, 0.5), torch.float32,
          "pw_reduce")
    gw, gb = K.pw_wgrad(xg, img.cuda(), c, 0.5, True, True)
    ew, eb = E.pw_wgrad(xa, img, c, 0.5, True, True)
    close(gw, ew, torch.float32, "pw_wgrad w")
    close(gb, eb, torch.float32, "pw_wgrad b")
    _, gb2 = K.pw_wgrad(xg, None, c, 1.0, False, True)
    close(gb2, eb, torch.float32, "chan sum")


@pytest.mark.parametrize("din", DTYPES)
@pytest.mark.parametrize("dout", DTYPES)
def test_resample_act(din, dout):
    xa, xg = act(2, 24, 2, 4, 8, din, seed=12)
    close(K.down2(xg, 0.125, dout), E.down2(xa, 0.125, dout), dout, "down2")
    close(K.up2(xg, 1.0, dout), E.up2(xa, 1.0, dout), dout, "up2")


def test_resample_img():
    img = rnd(3, 1, 4, 8, 16, seed=13)
    close(K.down2(img.cuda(), 0.125), E.down2(img, 0.125), torch.float32, "down2 img")
    close(K.up2(img.cuda(), 0.5), E.up2(img, 0.5), torch.float32, "up2 img")


@pytest.mark.parametrize("dtype", DTYPES)
def test_elementwise(dtype):
    aa, ag = act(2, 16, 2, 4, 4, dtype, seed=14)
    ba, bg = act(2, 16, 2, 4, 4, dtype, seed=15)
    close(K.lincomb(ag, bg, 0.3, 0.7), E.lincomb(aa, ba, 0.3, 0.7), dtype, "lincomb")
    close(K.lincomb(ag, None, -2.0, 0.0), E.lincomb(aa, None, -2.0, 0.0), dtype, "scale")
    close(K.lrelu_fwd(ag), E.lrelu_fwd(aa), dtype, "lrelu")
    close(K.mask_mul(ag, bg), E.mask_mul(aa, ba), dtype, "mask_mul")
    if dtype == torch.float32:      # ragged length through the scalar tail
        v = rnd(1003, seed=16)
        close(K.lincomb(v.cuda(), v.cuda(), 0.5, 0.25), v * 0.75, dtype, "lincomb tail")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [8, 16, 48])
@pytest.mark.parametrize("lrelu_after", [False, True])
def test_pixelnorm(dtype, c, lrelu_after):
    xa, xg = act(2, c, 2, 4, 8, dtype, seed=17)
    ga, gg = act(2, c, 2, 4, 8, dtype, seed=18)
    close(K.pixelnorm_fwd(xg, c, lrelu_after), E.pixelnorm_fwd(xa, c, lrelu_after), dtype, "pn fwd")
    close(K.pixelnorm_bwd(xg, gg, c, lrelu_after), E.pixelnorm_bwd(xa, ga, c, lrelu_after), dtype, "pn bwd")


def test_gp_helpers():
    real, fake = rnd(5, 1, 2, 8, 8, seed=19), rnd(5, 1, 2, 8, 8, seed=20)
    eps = torch.rand(5, generator=torch.Generator().manual_seed(21))
    close(K.interp(real.cuda(), fake.cuda(), eps.cuda()), E.interp(real, fake, eps), torch.float32, "interp")
    x = rnd(5, 1, 4, 16, 16, seed=22)
    close(K.sumsq_rows(x.cuda()), E.sumsq_rows(x), torch.float32, "sumsq")
    close(K.rowscale(x.cuda(), eps.cuda()), E.rowscale(x, eps), torch.float32, "rowscale")


@pytest.mark.parametrize("b,fin,fout", [(4, 512, 32), (8, 32, 1), (32, 32, 512), (3, 100, 17)])
def test_linear(b, fin, fout):
    x, wt, bias, g = rnd(b, fin, seed=23), rnd(fout, fin, seed=24), rnd(fout, seed=25), rnd(b, fout, seed=26)
    close(K.linear_fwd(x.cuda(), wt.cuda(), bias.cuda(), 0.1, True), E.linear_fwd(x, wt, bias, 0.1, True),
          torch.float32, "linear fwd")
    close(K.linear_dgrad(g.cuda(), wt.cuda(), 0.1), E.linear_dgrad(g, wt, 0.1), torch.float32, "linear dgrad")
    gw, gb = K.linear_wgrad(g.cuda(), x.cuda(), 0.1, True)
    ew, eb = E.linear_wgrad(g, x, 0.1, True)
    close(gw, ew, torch.float32, "linear wgrad")
    close(gb, eb, torch.float32, "linear bias grad")


def test_error_reporting():
    x = torch.zeros(8, device="cuda")
    with pytest.raises(RuntimeError, match="odd extent"):
        K._lib.call("sg_down2", x, x, 1, 1, 1, 1, 3, 2, 2, 1.0)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        K.lrelu_fwd(torch.zeros(8))


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_mask_variants(dtype):
    """LeakyReLU-backward masks fused into the avg-pool backward and the pixel-norm backward."""
    xa, xg = act(2, 24, 2, 4, 8, dtype, seed=30)
    ra, rg = act(2, 24, 4, 8, 16, dtype, seed=31)
    close(K.up2(xg, 0.125, dtype, rg), E.up2(xa, 0.125, dtype, ra), dtype, "up2 + mask")
    pa, pg = act(2, 16, 2, 4, 8, dtype, seed=32)
    ga, gg = act(2, 16, 2, 4, 8, dtype, seed=33)
    for lrelu_after in (False, True):
        close(K.pixelnorm_bwd(pg, gg, 16, lrelu_after, True), E.pixelnorm_bwd(pa, ga, 16, lrelu_after, True), dtype,
              "pixelnorm bwd + input mask")


@pytest.mark.parametrize("b,group,sub", [(4, 4, 1), (8, 4, 1), (6, 6, 1), (2, 2, 1), (3, 3, 1), (8, 4, 2), (24, 4, 3)])
def test_mbstd_kernels(b, group, sub):
    """forward, backward and double backward of minibatch-stddev against torch autograd through the
    reference formula (tests/cpu_emul.py does not use the hand-derived expressions of the kernels)."""
    c = 40
    x = rnd(b, c, 1, 4, 4, seed=40) + 0.3 * rnd(1, c, 1, 4, 4, seed=41)
    out, s = K.mbstd_fwd(x.cuda(), group, sub)
    eo, es = E.mbstd_fwd(x, group, sub)
    close(out, eo, torch.float32, "mbstd fwd")
    close(s, es, torch.float32, "mbstd s")
    gout = rnd(b, c + 1, 1, 4, 4, seed=42)
    gx, gt = K.mbstd_bwd(gout.cuda(), out, s, group, sub)
    egx, egt = E.mbstd_bwd(gout, eo, es, group, sub)
    close(gx, egx, torch.float32, "mbstd bwd")
    close(gt, egt, torch.float32, "mbstd gt")
    u = rnd(b, c, 1, 4, 4, seed=43)
    d_gout, d_x = K.mbstd_bwdbwd(u.cuda(), gt, out, s, group, sub)
    e_gout, e_x = E.mbstd_bwdbwd(u, egt, eo, es, group, sub)
    assert rel_err(d_gout.cpu(), e_gout) < 1e-4, "mbstd bwdbwd d_gout"
    # the second derivative of sqrt(mean xc^2 + 1e-8) blows up where a feature's group spread is ~0
    # (ill-conditioned in any arithmetic): compare where the stddev is not tiny, and measure the
    # error against the scale of the inputs too (for a group of 2 the exact result is ~0)
    m = b // (group * sub)
    ok = (es.reshape(sub, 1, m, c, 1, 4, 4) > 3e-2).expand(sub, group, m, c, 1, 4, 4).reshape(b, c, 1, 4, 4)
    scale = float((e_x * ok).norm()) + 1e-3 * float(u.norm()) * float(egt.abs().max())
    assert float(((d_x.cpu() - e_x) * ok).norm()) < 2e-4 * scale, "mbstd bwdbwd d_x"
