"""Autograd wiring of the CUDA kernels.

Every differentiable op of the PGAN step is a ``torch.autograd.Function`` whose
``backward`` is itself written in terms of these Functions (the closed set
fprop <-> dgrad <-> wgrad, avg-pool <-> nearest-upsample, expand <-> reduce), so
``torch.autograd.grad(..., create_graph=True)`` -- the WGAN-GP double backward of
loss.py:17-24 -- differentiates through them to any order.

torch is the tensor/stream/autograd host only; all arithmetic on activations happens in
libsaragan_b200.so.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import config
from . import kernels as K
from ._lib import IMPL_AUTO

_weight_grads_enabled = True


@contextlib.contextmanager
def no_weight_gradients():
    """Skip wgrad work inside the block.  Used around the gradient penalty's
    ``autograd.grad(outputs, inputs=interpolates, only_inputs=True)`` (loss.py:17-24): custom
    Functions cannot see that only the input gradient is wanted, and computing every layer's
    weight gradient there would cost a full extra wgrad pass."""
    global _weight_grads_enabled
    old = _weight_grads_enabled
    _weight_grads_enabled = False
    try:
        yield
    finally:
        _weight_grads_enabled = old


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


class PackedWeight:
    """Per-parameter cache of the packings of a (Cout,Cin,3,3,3) weight -- (dtype, flip) -> buffer -- re-packed IN PLACE
    when the optimiser bumps the parameter's version counter (persistent buffers: stable pointers for captured graphs
    and for the multi-tensor re-pack of `prepack`)."""

    def __init__(self, weight: torch.Tensor, cache: bool = True, known=None):
        self.weight = weight
        self.cache = cache
        self._bufs = {}
        self._ver = {}
        self.known = set() if known is None else known   # every (dtype, flip) ever asked for

    def _cur(self):
        w = self.weight
        return (w._version, w.data_ptr(), w.device)

    def stale(self, k) -> bool:
        return self._ver.get(k) != self._cur()

    def _reusable(self, k):
        buf = self._bufs.get(k)
        return buf if (buf is not None and getattr(buf, "device", None) == self.weight.device) else None

    def get(self, dtype: torch.dtype, flip: bool) -> torch.Tensor:
        k = (dtype, flip)
        self.known.add(k)
        if self.stale(k):
            self._bufs[k] = K.pack_conv_weight(self.weight.detach().contiguous(), dtype, flip, self._reusable(k))
            self._ver[k] = self._cur()
        return self._bufs[k]

    def stale_jobs(self):
        return [(k, (self.weight.detach().contiguous(), k[0], k[1], self._reusable(k))) for k in self.known if self.stale(k)]

    def adopt(self, k, buf) -> None:
        self._bufs[k], self._ver[k] = buf, self._cur()

    def prepack(self) -> None:
        for k in tuple(self.known):
            self.get(*k)


def prepack(net) -> None:
    """Re-pack every stale packing of every conv layer of `net` on the CURRENT stream with ONE launch
    (`sg_pack_conv_weights_multi`): after an optimiser step, and before work forks onto a second stream."""
    todo = []
    for m in net.modules():
        pw = getattr(m, "_packed", None)
        if pw is None:
            continue
        if getattr(pw, "weight", None) is not getattr(m, "weight", None):    # parameter re-bound (.to(), load)
            pw = m._packed = PackedWeight(m.weight, known=pw.known)
        todo += [(pw, k, job) for k, job in pw.stale_jobs()]
    if len(todo) == 1:
        todo[0][0].get(*todo[0][1])
    elif todo:
        for (pw, k, _), buf in zip(todo, K.pack_conv_weights_multi([job for _, _, job in todo])):
            pw.adopt(k, buf)


def packs_settled(net) -> bool:
    """True once `net` has run a complete pass at its current phase, i.e. every packing the pass needs exists and
    `prepack` refreshes it.  A layer's FIRST packing is made lazily at its first use and then cached; with the D
    phase forked onto two streams that first use would be on whichever stream gets there first while the other
    stream reads the same buffer un-ordered -- train.d_phase keeps such a pass (first step, after grow() or a
    change of `.phase`) on one stream."""
    return getattr(net, "_packs_phase", None) == (getattr(net, "phase", None), id(net), config.policy_key())


def mark_packs_settled(net) -> None:
    net._packs_phase = (getattr(net, "phase", None), id(net), config.policy_key())


# ================================================================================ conv
class Conv3x3(Function):
    """y = [m(out_mask) *] [lrelu](std * conv3d(x, w, pad=1) + b)   -- network.py:54-56 (+ :89 etc. fused).

    LeakyReLU-mask fusion (the mask of a layer is the sign of its OUTPUT):
      premasked       -- the consumer of y already multiplied the incoming gradient by m(y)
                         (a following conv's dgrad epilogue, an avg-pool backward, a pixel-norm
                         backward), so this op must not do it again;
      mask_input_grad -- x is the LeakyReLU output of the producing op: multiply the returned
                         input gradient by m(x) in the dgrad epilogue (the producer is premasked);
      dd_fuse         -- the same pairing one order up, for the gradient penalty's double backward: the backward of this
                         op's dgrad node multiplies what it returns by m(y) in ITS conv epilogue (`out_mask`), and does
                         not mask what it receives by m(x), because the producer of x -- wired with dd_fuse too --
                         did.  Without it each of those masks is a stand-alone MaskMul pass over a full-resolution
                         tensor.  Both sides of a producer/consumer pair must be wired alike (network.py's
                         Discriminator does so; stand-alone modules leave it off)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pw: Optional[PackedWeight], std: float, lrelu: bool,
                premasked: bool = False, mask_input_grad: bool = False, out_mask=None, dd_fuse: bool = False):
        cout, cin = weight.shape[0], weight.shape[1]
        pw = pw if pw is not None else PackedWeight(weight, cache=False)
        impl, kind = config.conv_route(x, cin, cout)
        y = K.conv3d_fprop(x, pw.get(kind, False), bias, out_mask, cin, cout, std, lrelu, impl)
        ctx.save_for_backward(x, weight, y if (lrelu and (dd_fuse or not premasked)) else None, out_mask)
        ctx.pw, ctx.std, ctx.lrelu, ctx.has_bias = pw, std, lrelu, bias is not None
        ctx.premasked, ctx.mask_input_grad, ctx.dd_fuse = premasked, mask_input_grad, dd_fuse
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y, out_mask = ctx.saved_tensors
        g = _c(gy)
        if out_mask is not None:
            g = MaskMul.apply(g, out_mask)
        if ctx.lrelu and not ctx.premasked:
            g = MaskMul.apply(g, y)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(g, weight, ctx.pw, ctx.std, x if ctx.mask_input_grad else None,
                                 y if (ctx.lrelu and ctx.premasked and ctx.dd_fuse) else None,
                                 ctx.dd_fuse and ctx.mask_input_grad)
        want_w = ctx.needs_input_grad[1] and _weight_grads_enabled
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled
        if want_w:
            gw, gb_ = ConvWgrad.apply(x, g, ctx.std, weight.shape[1], weight.shape[0], want_b)
            gb = gb_ if want_b else None
        elif want_b:
            gb = ChanSum.apply(g, weight.shape[0])
        return gx, gw, gb, None, None, None, None, None, None, None


class ConvDgrad(Function):
    """gx = [m(mask_ref) *] std * dgrad(g, w): the same implicit GEMM on the flipped/transposed
    packing; mask_ref fuses the producing layer's LeakyReLU backward into the epilogue.
    Double backward (dd_fuse wiring of Conv3x3): g_mask = the output of the layer this node belongs to, whose mask the
    incoming g already carries -- the gradient w.r.t. g is multiplied by m(g_mask) in the epilogue of the conv that
    computes it; ggx_premasked = the incoming second-order gradient already carries m(mask_ref)."""

    @staticmethod
    def forward(ctx, g, weight, pw: Optional[PackedWeight], std: float, mask_ref=None, g_mask=None,
                ggx_premasked: bool = False):
        cout, cin = weight.shape[0], weight.shape[1]
        pw = pw if pw is not None else PackedWeight(weight, cache=False)
        impl, kind = config.conv_route(g, cout, cin)
        gx = K.conv3d_fprop(g, pw.get(kind, True), None, mask_ref, cout, cin, std, False, impl)
        ctx.save_for_backward(g, weight, mask_ref, g_mask)
        ctx.pw, ctx.std, ctx.ggx_premasked = pw, std, ggx_premasked
        return gx

    @staticmethod
    def backward(ctx, ggx):
        g, weight, mask_ref, g_mask = ctx.saved_tensors
        ggx = _c(ggx)
        if mask_ref is not None and not ctx.ggx_premasked:
            ggx = MaskMul.apply(ggx, mask_ref)
        gg = gw = None
        if ctx.needs_input_grad[0]:
            gg = Conv3x3.apply(ggx, weight, None, ctx.pw, ctx.std, False, False, False, g_mask)
        if ctx.needs_input_grad[1] and _weight_grads_enabled:
            gw, _ = ConvWgrad.apply(ggx, g, ctx.std, weight.shape[1], weight.shape[0], False)
        return gg, gw, None, None, None, None, None


class ConvPool(Function):
    """pool = 1/8 * (2x2x2 block sums of lrelu(std * conv3d(x, w) + b)): the discriminator block's conv2 -> LeakyReLU ->
    AvgPool3d(2) (network.py:88-90) in ONE tcgen05 kernel.  The backward is exactly the graph the separate ops build with
    the block's fusion flags -- Up2 (masked by the conv output y) -> ConvDgrad / ConvWgrad -- so the gradient penalty's
    double backward differentiates through it like before."""

    @staticmethod
    def forward(ctx, x, weight, bias, pw: Optional[PackedWeight], std: float, mask_input_grad: bool, dd_fuse: bool):
        cout, cin = weight.shape[0], weight.shape[1]
        pw = pw if pw is not None else PackedWeight(weight, cache=False)
        y, y_pool = K.conv3d_fprop_pool(x, pw.get(x.dtype, False), bias, cin, cout, std, True, 0.125)
        ctx.save_for_backward(x, weight, y)
        ctx.pw, ctx.std, ctx.has_bias, ctx.mask_input_grad, ctx.dd_fuse = pw, std, bias is not None, mask_input_grad, dd_fuse
        return y_pool

    @staticmethod
    def backward(ctx, gpool):
        x, weight, y = ctx.saved_tensors
        g = Up2.apply(_c(gpool), 0.125, y.dtype, y, ctx.dd_fuse)          # avg-pool backward with the LeakyReLU mask
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(g, weight, ctx.pw, ctx.std, x if ctx.mask_input_grad else None,
                                 y if ctx.dd_fuse else None, ctx.dd_fuse and ctx.mask_input_grad)
        want_w = ctx.needs_input_grad[1] and _weight_grads_enabled
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled
        if want_w:
            gw, gb_ = ConvWgrad.apply(x, g, ctx.std, weight.shape[1], weight.shape[0], want_b)
            gb = gb_ if want_b else None
        elif want_b:
            gb = ChanSum.apply(g, weight.shape[0])
        return gx, gw, gb, None, None, None, None


class ConvPixelNorm(Function):
    """y = [lrelu_after](pixel_norm([lrelu](std * conv3d(x, w) + b))): a generator-block convolution with its
    ChannelNormalization (and the LeakyReLU on either side, network.py:204-216) in ONE tcgen05 kernel -- the epilogue
    thread of a voxel holds all its channels.  First order only (the generator is never differentiated twice); the
    backward is pixel-norm backward (with the inner LeakyReLU's mask) -> dgrad / wgrad."""

    @staticmethod
    def forward(ctx, x, weight, bias, pw: Optional[PackedWeight], std: float, lrelu: bool, lrelu_after: bool):
        cout, cin = weight.shape[0], weight.shape[1]
        pw = pw if pw is not None else PackedWeight(weight, cache=False)
        y, y_norm = K.conv3d_fprop_pixelnorm(x, pw.get(x.dtype, False), bias, cin, cout, std, lrelu, lrelu_after)
        ctx.save_for_backward(x, weight, y)
        ctx.pw, ctx.std, ctx.lrelu, ctx.lrelu_after, ctx.has_bias = pw, std, lrelu, lrelu_after, bias is not None
        return y_norm

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        cout, cin = weight.shape[0], weight.shape[1]
        # through [lrelu_after] and the normalisation, then through the conv's own LeakyReLU (mask_input: y is its output)
        g = K.pixelnorm_bwd(y, _c(gy), cout, ctx.lrelu_after, ctx.lrelu)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            impl, kind = config.conv_route(g, cout, cin)
            gx = K.conv3d_fprop(g, ctx.pw.get(kind, True), None, None, cout, cin, ctx.std, False, impl)
        want_w = ctx.needs_input_grad[1] and _weight_grads_enabled
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled
        if want_w:
            gw, gb_ = K.conv3d_wgrad(x, g, cin, cout, ctx.std, want_b, config.wgrad_impl(x, cin, cout))
            gb = gb_ if want_b else None
        elif want_b:
            gb = K.pw_wgrad(g, None, cout, 1.0, False, True)[1]
        return gx, gw, gb, None, None, None, None


class ConvWgrad(Function):
    """gw = std * sum_p g (x) x,  gb = sum_p g   (fp32)."""

    @staticmethod
    def forward(ctx, x, g, std: float, cin: int, cout: int, want_bias: bool):
        gw, gb = K.conv3d_wgrad(x, g, cin, cout, std, want_bias, config.wgrad_impl(x, cin, cout))
        ctx.save_for_backward(x, g)
        ctx.std, ctx.cin, ctx.cout = std, cin, cout
        if gb is None:
            gb = gw.new_zeros(())
            ctx.mark_non_differentiable(gb)
        return gw, gb

    @staticmethod
    def backward(ctx, ggw, ggb):
        x, g = ctx.saved_tensors
        ggw = _c(ggw)
        gx = gg = None
        if ctx.needs_input_grad[0]:
            gx = ConvDgrad.apply(g, ggw, None, ctx.std)
        if ctx.needs_input_grad[1]:
            gg = Conv3x3.apply(x, ggw, _c(ggb) if ggb is not None and ggb.dim() == 1 else None, None,
                               ctx.std, False)
        return gx, gg, None, None, None, None


class ChanSum(Function):
    """out[c] = sum over batch and voxels (bias gradient)."""

    @staticmethod
    def forward(ctx, g, c: int):
        ctx.shape, ctx.dtype, ctx.c = g.shape, g.dtype, c
        return K.pw_wgrad(g, None, c, 1.0, False, True)[1]

    @staticmethod
    def backward(ctx, gb):
        n, cc, d, h, w, _ = ctx.shape
        ones = torch.ones((n, 1, d, h, w), dtype=torch.float32, device=gb.device)
        return PwExpand.apply(ones, _c(gb), None, 1.0, False, ctx.c, ctx.dtype), None


class MaskMul(Function):
    """y = g * (ref > 0 ? 1 : 0.2): LeakyReLU backward from the sign of the layer OUTPUT; linear
    in g, so its own backward is the same op (LeakyReluBackwardBackward)."""

    @staticmethod
    def forward(ctx, g, ref):
        ctx.save_for_backward(ref)
        return K.mask_mul(g, ref)

    @staticmethod
    def backward(ctx, gg):
        (ref,) = ctx.saved_tensors
        return MaskMul.apply(_c(gg), ref), None


class LeakyRelu(Function):
    @staticmethod
    def forward(ctx, x):
        y = K.lrelu_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(_c(gy), y)


# ========================================================================== resampling
class Down2(Function):
    """y = scale * (2x2x2 block sum): AvgPool3d(2) with scale 1/8 (network.py:90,154)."""

    @staticmethod
    def forward(ctx, x, scale: float, out_dtype: Optional[torch.dtype] = None, premask: bool = False,
                dd_fuse: bool = False):
        """premask: x is the LeakyReLU output of a (premasked) conv; fold that conv's mask into
        this op's backward (the up-sampling kernel multiplies by m(x)).  dd_fuse: see Conv3x3."""
        ctx.scale, ctx.in_dtype, ctx.dd_fuse = scale, x.dtype, dd_fuse
        ctx.save_for_backward(x if premask else None)
        return K.down2(x, scale, out_dtype)

    @staticmethod
    def backward(ctx, gy):
        (ref,) = ctx.saved_tensors
        return Up2.apply(_c(gy), ctx.scale, ctx.in_dtype, ref, ctx.dd_fuse and ref is not None), None, None, None, None


class Up2(Function):
    """y[child] = scale * x[parent]: nearest Upsample(2) with scale 1 (network.py:203,265).
    gy_premasked: the gradient arriving in backward already carries m(mask_ref) (dd_fuse wiring of Conv3x3)."""

    @staticmethod
    def forward(ctx, x, scale: float, out_dtype: Optional[torch.dtype] = None, mask_ref=None,
                gy_premasked: bool = False):
        ctx.scale, ctx.in_dtype, ctx.gy_premasked = scale, x.dtype, gy_premasked
        ctx.save_for_backward(mask_ref)
        return K.up2(x, scale, out_dtype, mask_ref)

    @staticmethod
    def backward(ctx, gy):
        (ref,) = ctx.saved_tensors
        gy = _c(gy)
        if ref is not None and not ctx.gy_premasked:
            gy = MaskMul.apply(gy, ref)
        return Down2.apply(gy, ctx.scale, ctx.in_dtype), None, None, None, None


class Lincomb(Function):
    """y = alpha*a + beta*b: the fade-in blend (network.py:185,281).  alpha, beta are Python floats, or -- for a tensor
    alpha -- `alpha` = the device pair {alpha, beta} and `beta` = the swapped pair {beta, alpha} (`blend_coef`)."""

    @staticmethod
    def forward(ctx, a, b, alpha, beta):
        ctx.alpha, ctx.beta, ctx.has_b = alpha, beta, b is not None
        return K.lincomb(a, b, alpha, beta)

    @staticmethod
    def backward(ctx, gy):
        gy = _c(gy)
        dev = isinstance(ctx.alpha, torch.Tensor)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = Lincomb.apply(gy, None, ctx.alpha, ctx.beta if dev else 0.0)
        if ctx.has_b and ctx.needs_input_grad[1]:
            gb = Lincomb.apply(gy, None, ctx.beta, ctx.alpha if dev else 0.0)
        return ga, gb, None, None


def blend_coef(alpha, device):
    """(alpha, 1 - alpha) as the arguments of Lincomb: floats for a Python number, device pairs for a tensor alpha
    (no `.item()` host sync; inside a captured graph the kernels then read the current value on every replay)."""
    if isinstance(alpha, torch.Tensor):
        a = alpha.detach().to(device=device, dtype=torch.float32).reshape(())
        pair = torch.stack([a, 1.0 - a])
        return pair, pair.flip(0)
    return float(alpha), 1.0 - float(alpha)


# =============================================================================== 1x1x1
class PwExpand(Function):
    """FromRGB (network.py:101-110): y[n,c,v] = [m(out_mask) *] [lrelu](std*w[c]*img[n,v] + b[c]).
    premasked / dd_fuse: as in Conv3x3 (the consumer is the top block's conv1)."""

    @staticmethod
    def forward(ctx, img, w, bias, std: float, lrelu: bool, c: int, dtype: torch.dtype,
                premasked: bool = False, dd_fuse: bool = False, out_mask=None):
        y = K.pw_expand(img, w, bias, dtype, c, std, lrelu, out_mask)
        ctx.save_for_backward(img, w, y if (lrelu and (dd_fuse or not premasked)) else None, out_mask)
        ctx.std, ctx.lrelu, ctx.c, ctx.has_bias, ctx.premasked, ctx.dd_fuse = std, lrelu, c, bias is not None, premasked, dd_fuse
        return y

    @staticmethod
    def backward(ctx, gy):
        img, w, y, out_mask = ctx.saved_tensors
        g = _c(gy)
        if out_mask is not None:
            g = MaskMul.apply(g, out_mask)
        if ctx.lrelu and not ctx.premasked:
            g = MaskMul.apply(g, y)
        gimg = gw = gb = None
        if ctx.needs_input_grad[0]:
            gimg = PwReduce.apply(g, w, None, ctx.std, ctx.c, y if (ctx.lrelu and ctx.premasked and ctx.dd_fuse) else None)
        want_w = ctx.needs_input_grad[1] and _weight_grads_enabled
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled
        if want_w or want_b:
            gw_, gb_ = PwWgrad.apply(g, img, ctx.std, ctx.c)
            gw = gw_ if want_w else None
            gb = gb_ if want_b else None
        return gimg, gw, gb, None, None, None, None, None, None, None


class PwReduce(Function):
    """ToRGB (network.py:219-225): img[n,v] = std*sum_c w[c]*x[n,c,v] + b.
    x_mask (dd_fuse wiring): x already carries m(x_mask); the gradient w.r.t. x is masked by it where it is computed."""

    @staticmethod
    def forward(ctx, x, w, bias, std: float, c: int, x_mask=None):
        img = K.pw_reduce(x, w, bias, c, std)
        ctx.save_for_backward(x, w, x_mask)
        ctx.std, ctx.c, ctx.has_bias = std, c, bias is not None
        return img

    @staticmethod
    def backward(ctx, gimg):
        x, w, x_mask = ctx.saved_tensors
        gimg = _c(gimg)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = PwExpand.apply(gimg, w, None, ctx.std, False, ctx.c, x.dtype, False, False, x_mask)
        if ctx.needs_input_grad[1] and _weight_grads_enabled:
            gw, _ = PwWgrad.apply(x, gimg, ctx.std, ctx.c)
        if ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled:
            gb = gimg.sum().reshape(1)
        return gx, gw, gb, None, None, None


class PwWgrad(Function):
    """gw[c] = std*sum g[n,c,v]*img[n,v],  gb[c] = sum g[n,c,v]."""

    @staticmethod
    def forward(ctx, g, img, std: float, c: int):
        gw, gb = K.pw_wgrad(g, img, c, std, True, True)
        ctx.save_for_backward(g, img)
        ctx.std, ctx.c = std, c
        return gw, gb

    @staticmethod
    def backward(ctx, ggw, ggb):
        g, img = ctx.saved_tensors
        gg = gimg = None
        if ctx.needs_input_grad[0]:
            gg = PwExpand.apply(img, _c(ggw), _c(ggb), ctx.std, False, ctx.c, g.dtype)
        if ctx.needs_input_grad[1]:
            gimg = PwReduce.apply(g, _c(ggw), None, ctx.std, ctx.c)
        return gg, gimg, None, None


# ========================================================================== pixel-norm
class PixelNorm(Function):
    """ChannelNormalization (network.py:192-197) [+ LeakyReLU]; generator only, first order."""

    @staticmethod
    def forward(ctx, x, c: int, lrelu_after: bool, mask_input: bool = False):
        """mask_input: x is the LeakyReLU output of a (premasked) conv; the backward kernel also
        applies that conv's mask m(x)."""
        ctx.save_for_backward(x)
        ctx.c, ctx.lrelu_after, ctx.mask_input = c, lrelu_after, mask_input
        return K.pixelnorm_fwd(x, c, lrelu_after)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return K.pixelnorm_bwd(x, _c(gy), ctx.c, ctx.lrelu_after, ctx.mask_input), None, None, None


# ============================================================================== layout
class ToAct(Function):
    """plain fp32 NCDHW -> blocked activation layout."""

    @staticmethod
    def forward(ctx, plain, dtype: torch.dtype):
        ctx.c = plain.shape[1]
        return K.plain_to_act(_c(plain), dtype)

    @staticmethod
    def backward(ctx, g):
        return ToPlain.apply(_c(g), ctx.c), None


class ToPlain(Function):
    """blocked activation layout -> plain fp32 NCDHW."""

    @staticmethod
    def forward(ctx, act, c: int):
        ctx.dtype = act.dtype
        return K.act_to_plain(act, c)

    @staticmethod
    def backward(ctx, g):
        return ToAct.apply(_c(g), ctx.dtype), None


# ============================================================================== linear
class Linear(Function):
    """y = [lrelu](std * x @ w.T + b)   -- network.py:76-77."""

    @staticmethod
    def forward(ctx, x, w, bias, std: float, lrelu: bool):
        y = K.linear_fwd(x, w, bias, std, lrelu)
        ctx.save_for_backward(x, w, y if lrelu else None)
        ctx.std, ctx.lrelu, ctx.has_bias = std, lrelu, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        g = _c(gy)
        if ctx.lrelu:
            g = MaskMul.apply(g, y)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = LinearDgrad.apply(g, w, ctx.std)
        want_w = ctx.needs_input_grad[1] and _weight_grads_enabled
        want_b = ctx.has_bias and ctx.needs_input_grad[2] and _weight_grads_enabled
        if want_w or want_b:
            gw_, gb_ = LinearWgrad.apply(g, x, ctx.std)
            gw = gw_ if want_w else None
            gb = gb_ if want_b else None
        return gx, gw, gb, None, None


class LinearDgrad(Function):
    @staticmethod
    def forward(ctx, g, w, std: float):
        ctx.save_for_backward(g, w)
        ctx.std = std
        return K.linear_dgrad(g, w, std)

    @staticmethod
    def backward(ctx, ggx):
        g, w = ctx.saved_tensors
        ggx = _c(ggx)
        gg = gw = None
        if ctx.needs_input_grad[0]:
            gg = Linear.apply(ggx, w, None, ctx.std, False)
        if ctx.needs_input_grad[1] and _weight_grads_enabled:
            gw, _ = LinearWgrad.apply(g, ggx, ctx.std)
        return gg, gw, None


class LinearWgrad(Function):
    @staticmethod
    def forward(ctx, g, x, std: float):
        ctx.save_for_backward(g, x)
        ctx.std = std
        return K.linear_wgrad(g, x, std, True)

    @staticmethod
    def backward(ctx, ggw, ggb):
        g, x = ctx.saved_tensors
        gg = gx = None
        if ctx.needs_input_grad[0]:
            gg = Linear.apply(x, _c(ggw), _c(ggb), ctx.std, False)
        if ctx.needs_input_grad[1]:
            gx = LinearDgrad.apply(g, _c(ggw), ctx.std)
        return gg, gx, None


# ================================================================== minibatch stddev
class Mbstd(Function):
    """MinibatchStandardDeviation (network.py:113-133) on the plain fp32 base-level tensor.  The only
    non-piecewise-linear op of D: its backward is a Function with its own backward (the second
    derivative of sqrt(mean xc^2) that the gradient penalty's double backward needs)."""

    @staticmethod
    def forward(ctx, x, group: int, sub_batches: int = 1):
        x = _c(x)
        out, s = K.mbstd_fwd(x, group, sub_batches)
        ctx.save_for_backward(x, out, s)
        ctx.group, ctx.sub = group, sub_batches
        return out

    @staticmethod
    def backward(ctx, gout):
        x, out, s = ctx.saved_tensors
        return MbstdBwd.apply(_c(gout), x, out.detach(), s, ctx.group, ctx.sub), None, None


class MbstdBwd(Function):
    """gx of Mbstd; `x` is passed only as the handle the second-order gradient flows back to
    (`out` and `s` are functions of it and enter as constants)."""

    @staticmethod
    def forward(ctx, gout, x, out, s, group: int, sub_batches: int = 1):
        gx, gt = K.mbstd_bwd(gout, out, s, group, sub_batches)
        ctx.save_for_backward(gt, out, s)
        ctx.group, ctx.sub = group, sub_batches
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        gt, out, s = ctx.saved_tensors
        d_gout, d_x = K.mbstd_bwdbwd(_c(u), gt, out, s, ctx.group, ctx.sub)
        return d_gout, d_x, None, None, None, None


# ==================================================================== gradient penalty
class RowNorm(Function):
    """norm[n] = ||x[n].flatten()||_2   (loss.py:25-26 `gradients.norm(2, dim=1)`).  The
    backward seeds the double-backward chain: g_hat[n] = gn[n] * x[n] / norm[n] (0 at norm 0,
    like torch's norm backward)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        norm = torch.sqrt(K.sumsq_rows(x))
        ctx.save_for_backward(x, norm)
        return norm

    @staticmethod
    @once_differentiable
    def backward(ctx, gn):
        x, norm = ctx.saved_tensors
        coef = torch.where(norm > 0, gn / norm, torch.zeros_like(norm)).contiguous()
        return K.rowscale(x, coef)


def interpolate(real: torch.Tensor, fake: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """eps*real + (1-eps)*fake  (loss.py:13); not differentiated (the result is the leaf)."""
    return K.interp(_c(real), _c(fake), _c(eps.reshape(-1).float()))
