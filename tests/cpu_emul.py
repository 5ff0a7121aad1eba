"""CPU emulation of saragan_b200.kernels for the `not gpu` tests (TEST INFRASTRUCTURE).

Each function restates one C-ABI entry point with torch CPU ops on the same blocked layout, so
that the host logic above the ABI -- the autograd wiring, the double backward of the gradient
penalty, the module plumbing, the gradient bucketing -- can be checked against the oracle on a
box without a GPU.  It is injected with monkeypatch by tests/conftest.py's `cpu_kernels`
fixture; nothing under saragan_b200/ knows it exists.
"""
import torch
import torch.nn.functional as F


LEAK = 0.2   # sg_set_leaky_slope / sg_get_leaky_slope (process-wide, like the library's)


def set_leaky_slope(slope):
    global LEAK
    assert 0.0 <= slope <= 1.0
    LEAK = float(slope)


def get_leaky_slope():
    return LEAK


def chunks(c):
    return 2 * ((c + 15) // 16)


def plain_to_act(plain, dtype):
    n, c, d, h, w = plain.shape
    cc = chunks(c)
    pad = torch.zeros((n, cc * 8, d, h, w), dtype=torch.float32)
    pad[:, :c] = plain
    return pad.view(n, cc, 8, d, h, w).permute(0, 1, 3, 4, 5, 2).contiguous().to(dtype)


def act_to_plain(act, c):
    n, cc, d, h, w, _ = act.shape
    return act.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cc * 8, d, h, w)[:, :c].contiguous()


def _tf32(t):
    """round to nearest to 10 mantissa bits (cvt.rna.tf32.f32: ties away from zero in magnitude)"""
    i = t.detach().float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _Packed:
    """stand-in for the packed weight buffer: keeps the logical (Cout,Cin,3,3,3) tensor, rounded like the packing"""

    def __init__(self, w, kind, flip):
        self.w = _tf32(w) if kind == "tf32" else w.detach().to(kind).float()
        self.kind = kind
        self.flip = flip


def conv_tf32_supported(n, cin, cout, d, h, w):
    return w % 8 == 0 and h >= 8 and h % min(h, 16) == 0


def pack_conv_weight(w, dtype, flip, out=None):
    return _Packed(w, dtype, flip)


def pack_conv_weights_multi(jobs):
    return [_Packed(w, dtype, flip) for w, dtype, flip, _ in jobs]


def _mask(ref):
    return torch.where(ref.float() > 0, 1.0, LEAK)


def conv3d_fprop(x, wp, bias, mask_src, cin, cout, scale, lrelu, impl=0):
    xp = act_to_plain(x, cin)
    assert (impl == 3) == (wp.kind == "tf32"), "SG_IMPL_TF32 goes with the SG_TF32 packing"
    if impl == 3:       # kind::tf32: the kernel rounds the landed activations to tf32
        assert x.dtype == torch.float32
        xp = _tf32(xp)
    if wp.flip:   # dgrad packing: contraction over the weight's Cout, flipped taps
        w = wp.w.flip(2, 3, 4).transpose(0, 1)
    else:
        w = wp.w
    y = F.conv3d(xp, w, None, 1, 1) * scale
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if lrelu:
        y = F.leaky_relu(y, LEAK)
    out = plain_to_act(y, x.dtype)
    if mask_src is not None:
        out = (out.float() * _mask(mask_src)).to(x.dtype)
    return out


def conv_pixelnorm_supported(x, cin, cout):
    """mirror of sg_conv3d_pixelnorm_supported: bf16, H % 16 == 0, W % 8 == 0, the resident weights of ONE N tile fit"""
    n, _, d, h, w, _ = x.shape
    coutp, ccin = 16 * ((cout + 15) // 16), chunks(cin)
    return (x.dtype == torch.bfloat16 and h % 16 == 0 and w % 8 == 0 and coutp in (16, 32, 64)
            and 27 * ccin * coutp * 16 <= 120 * 1024 and d >= 2 and d % 2 == 0)


def conv3d_fprop_pixelnorm(x, wp, bias, cin, cout, scale, lrelu, lrelu_after):
    """the fused kernel normalises the fp32 accumulator values (y itself is rounded to bf16 only when stored)"""
    xp = act_to_plain(x, cin)
    y = F.conv3d(xp, wp.w, None, 1, 1) * scale
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if lrelu:
        y = F.leaky_relu(y, LEAK)
    yn = y * torch.rsqrt(torch.mean(y ** 2, dim=1, keepdim=True) + 1e-8)
    if lrelu_after:
        yn = F.leaky_relu(yn, LEAK)
    return plain_to_act(y, x.dtype), plain_to_act(yn, x.dtype)


def conv_pool_supported(x, cin, cout):
    """mirror of sg_conv3d_pool_supported: like the resident kernel's coverage, even depth"""
    return conv_pixelnorm_supported(x, cin, cout) or (
        x.dtype == torch.bfloat16 and x.shape[3] % 16 == 0 and x.shape[4] % 8 == 0 and x.shape[2] % 2 == 0
        and 16 * ((cout + 15) // 16) in (16, 32, 64) and 27 * chunks(cin) * 16 * ((cout + 15) // 16) * 16 <= 120 * 1024)


def conv3d_fprop_pool(x, wp, bias, cin, cout, scale, lrelu, pool_scale):
    """the fused kernel pools the fp32 values (y itself is rounded to bf16 only when stored)"""
    xp = act_to_plain(x, cin)
    y = F.conv3d(xp, wp.w, None, 1, 1) * scale
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if lrelu:
        y = F.leaky_relu(y, LEAK)
    return plain_to_act(y, x.dtype), plain_to_act(F.avg_pool3d(y, 2) * (8.0 * pool_scale), x.dtype)


def conv3d_wgrad(x, gy, cin, cout, scale, want_bias, impl=0):
    xp = act_to_plain(x, cin)
    gp = act_to_plain(gy, cout)
    wg = lambda a, b: torch.nn.grad.conv3d_weight(a, (cout, cin, 3, 3, 3), b, stride=1, padding=1)     # noqa: E731
    if impl == 3:
        # the fp32 levels' tensor-core wgrad: bf16 hi/lo halves, g (x) x ~ (g_hi + g_lo) (x) x_hi + g_hi (x) x_lo
        x_hi, g_hi = xp.bfloat16().float(), gp.bfloat16().float()
        x_lo, g_lo = (xp - x_hi).bfloat16().float(), (gp - g_hi).bfloat16().float()
        gw = (wg(x_hi, g_hi + g_lo) + wg(x_lo, g_hi)) * scale
        gp = g_hi + g_lo
    elif impl == 4:
        gp = gp.bfloat16().float()      # fp32 tensors, bf16 operands (the bias gradient sums the rounded gy)
        gw = wg(xp.bfloat16().float(), gp) * scale
    else:
        gw = wg(xp, gp) * scale
    gb = gp.sum(dim=(0, 2, 3, 4)) if want_bias else None
    return gw.contiguous(), gb


def pw_expand(img, w, bias, dtype, c, scale, lrelu, mask_ref=None):
    y = img * (w.view(1, -1, 1, 1, 1) * scale)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if lrelu:
        y = F.leaky_relu(y, LEAK)
    out = plain_to_act(y, dtype)
    if mask_ref is not None:
        out = (out.float() * _mask(mask_ref)).to(dtype)
    return out


def pw_reduce(x, w, bias, c, scale):
    xp = act_to_plain(x, c)
    img = (xp * w.view(1, -1, 1, 1, 1)).sum(1, keepdim=True) * scale
    if bias is not None:
        img = img + bias.view(1, 1, 1, 1, 1)
    return img.contiguous()


def pw_wgrad(g, img, c, scale, want_w, want_b):
    gp = act_to_plain(g, c)
    gw = (gp * img).sum(dim=(0, 2, 3, 4)) * scale if want_w else None
    gb = gp.sum(dim=(0, 2, 3, 4)) if want_b else None
    return gw, gb


def _to5(x):
    if x.dim() == 6:
        n, cc, d, h, w, _ = x.shape
        return x.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cc * 8, d, h, w), True
    return x, False


def _from5(y, blocked, dtype):
    if blocked:
        n, c8, d, h, w = y.shape
        return y.view(n, c8 // 8, 8, d, h, w).permute(0, 1, 3, 4, 5, 2).contiguous().to(dtype)
    return y.contiguous()


def down2(x, scale, out_dtype=None):
    y, blocked = _to5(x)
    return _from5(F.avg_pool3d(y, 2) * (8.0 * scale), blocked, out_dtype or x.dtype)


def up2(x, scale, out_dtype=None, mask_ref=None):
    y, blocked = _to5(x)
    out = _from5(F.interpolate(y, scale_factor=2, mode="nearest") * scale, blocked, out_dtype or x.dtype)
    if mask_ref is not None:
        out = (out.float() * _mask(mask_ref)).to(out.dtype)
    return out


def lincomb(a, b, alpha, beta=None):
    if isinstance(alpha, torch.Tensor):      # device pair {alpha, beta} (sg_lincomb_dev)
        alpha, beta = float(alpha[0]), float(alpha[1])
    y = a.float() * alpha
    if b is not None:
        y = y + b.float() * beta
    return y.to(a.dtype)


def lrelu_fwd(x):
    return F.leaky_relu(x.float(), LEAK).to(x.dtype)


def mask_mul(g, ref):
    return (g.float() * _mask(ref)).to(g.dtype)


def pixelnorm_fwd(x, c, lrelu_after):
    xp = act_to_plain(x, c)
    y = xp * torch.rsqrt(torch.mean(xp ** 2, dim=1, keepdim=True) + 1e-8)
    if lrelu_after:
        y = F.leaky_relu(y, LEAK)
    return plain_to_act(y, x.dtype)


def pixelnorm_bwd(x, gy, c, lrelu_after, mask_input=False):
    xp = act_to_plain(x, c).requires_grad_(True)
    with torch.enable_grad():
        y = xp * torch.rsqrt(torch.mean(xp ** 2, dim=1, keepdim=True) + 1e-8)
        if lrelu_after:
            y = F.leaky_relu(y, LEAK)
        (gx,) = torch.autograd.grad(y, xp, act_to_plain(gy, c))
    if mask_input:
        gx = gx * _mask(xp.detach())
    return plain_to_act(gx, x.dtype)


def interp(real, fake, eps):
    e = eps.view(-1, 1, 1, 1, 1)
    return e * real + (1 - e) * fake


def sumsq_rows(x):
    return (x.reshape(x.shape[0], -1) ** 2).sum(1)


def rowscale(x, s):
    return x * s.view(-1, *([1] * (x.dim() - 1)))


def linear_fwd(x, w, bias, scale, lrelu):
    y = F.linear(x, w) * scale
    if bias is not None:
        y = y + bias
    return F.leaky_relu(y, LEAK) if lrelu else y


def linear_dgrad(g, w, scale):
    return (g @ w) * scale


def linear_wgrad(g, x, scale, want_bias):
    return (g.t() @ x) * scale, (g.sum(0) if want_bias else None)


def _mbstd_ref(x, group, sub=1):
    """the reference formula (network.py:118-133) applied to each of `sub` stacked minibatches,
    differentiable by torch autograd"""
    outs, sds = [], []
    for xs in x.chunk(sub, dim=0):
        b, c, d, h, w = xs.shape
        y = xs.reshape(group, -1, c, d, h, w)
        yc = y - torch.mean(y, dim=0, keepdim=True)
        sd = torch.sqrt(torch.mean(yc ** 2, dim=0) + 1e-8)
        t = torch.mean(sd, dim=[1, 2, 3, 4], keepdim=True).repeat([group, 1, d, h, w])
        outs.append(torch.cat([yc.reshape(b, c, d, h, w), t], dim=1))
        sds.append(sd.reshape(sd.shape[0], -1))
    return torch.cat(outs), torch.cat(sds)


def mbstd_fwd(x, group, sub_batches=1):
    return _mbstd_ref(x, group, sub_batches)


def _recover_x(out):
    # any x with the same group-centred values gives the same derivatives
    return out[:, :-1].detach().clone()


def _stat_grad(gout, group, sub):
    b = gout.shape[0]
    m = b // (group * sub)
    return gout[:, -1].reshape(sub, group, m, -1).sum(dim=(1, 3)).reshape(-1)


def mbstd_bwd(gout, out, s, group, sub_batches=1):
    """independent of the hand-derived kernel formulas: torch autograd through the reference"""
    x = _recover_x(out).requires_grad_(True)
    with torch.enable_grad():
        o, _ = _mbstd_ref(x, group, sub_batches)
        (gx,) = torch.autograd.grad(o, x, gout)
    return gx, _stat_grad(gout, group, sub_batches)


def mbstd_bwdbwd(u, gt, out, s, group, sub_batches=1):
    """torch double backward through the reference; gout's stat channel is rebuilt from gt"""
    b, c1, d, h, w = out.shape
    m = b // (group * sub_batches)
    x = _recover_x(out).requires_grad_(True)
    gout = torch.zeros_like(out)
    per_pos = (gt / (group * d * h * w)).reshape(sub_batches, 1, m).expand(sub_batches, group, m).reshape(b)
    gout[:, -1] = per_pos.view(b, 1, 1, 1)
    gout.requires_grad_(True)
    with torch.enable_grad():
        o, _ = _mbstd_ref(x, group, sub_batches)
        (gx,) = torch.autograd.grad(o, x, gout, create_graph=True)
        d_gout, d_x = torch.autograd.grad(gx, [gout, x], u)
    return d_gout, d_x


ALL = [n for n in dir() if not n.startswith("_") and n not in ("torch", "F", "chunks", "ALL", "LEAK")]


# ------------------------------------------------------------------ evaluation metrics (csrc/metrics.cu)
# restated with numpy / scipy directly from the kernels' contracts in include/saragan_b200.h
def _binom3():
    import numpy as np
    f = np.array([1, 4, 6, 4, 1], dtype=np.float64) / 16
    return f[:, None, None] * f[None, :, None] * f[None, None, :]


def pyr_down(x):
    import scipy.ndimage
    y = scipy.ndimage.convolve(x.numpy().astype("float64"), _binom3()[None, None], mode="mirror")
    return torch.from_numpy(y[:, :, ::2, ::2, ::2].astype("float32"))


def pyr_up_sub(fine, coarse):
    import numpy as np
    import scipy.ndimage
    n, c, cd, ch, cw = coarse.shape
    assert tuple(fine.shape) == (n, c, 2 * cd, 2 * ch, 2 * cw)
    z = np.zeros(tuple(fine.shape), dtype=np.float64)
    z[:, :, ::2, ::2, ::2] = coarse.numpy()
    up = scipy.ndimage.convolve(z, 4.0 * _binom3()[None, None], mode="mirror").astype("float32")
    return fine - torch.from_numpy(up)


def swd_descriptors(level, pos_z, pos_y, pos_x, out):
    b = level.shape[0]
    dz, dx, dy = torch.meshgrid(torch.arange(-1, 2), torch.arange(-4, 5), torch.arange(-4, 5), indexing="ij")
    for j in range(pos_z.numel()):
        patch = level[:, 0, (int(pos_z[j]) + dz), (int(pos_y[j]) + dy), (int(pos_x[j]) + dx)].reshape(b, -1).double()
        patch = (patch - patch.mean()).float().double()
        patch = (patch / patch.std(unbiased=False).float().double()).float()
        out[:, j * 243:(j + 1) * 243] = patch


def swd_project(a, dirs, want_colsq):
    return a @ dirs, ((dirs * dirs).sum(0) if want_colsq else None)


def swd_finish(p, colsq, out):
    b = p.shape[0] // 2
    if colsq is not None:
        p = p * torch.rsqrt(colsq)
    out[0] = (p[:b].sort(0).values - p[b:].sort(0).values).abs().mean()


def value_hist(x, intercept, lo, hi):
    import numpy as np
    v = ((x.numpy() * np.float32(intercept)) + np.float32(intercept)).astype(np.int64).clip(lo, hi)
    return torch.from_numpy(np.stack([np.bincount(r - lo, minlength=hi - lo + 1) for r in v]).astype(np.int32))


ALL = [n for n in dir() if not n.startswith("_") and n not in ("torch", "F", "chunks", "ALL", "LEAK")]
