"""Tensor-level wrappers over the C ABI: allocate outputs, launch, return.

Every function here maps 1:1 onto an entry point of include/saragan_b200.h.  Shapes:
  act   (N, CC, D, H, W, 8)  bf16 | fp32, CC = 2*ceil(C/16)
  img   (N, 1, D, H, W)      fp32
  plain (N, C, D, H, W)      fp32
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, chunks

EPS_PN = 1e-8

# Optional per-launch timing of the convolution kernels (bench.py installs a ConvProbe to time the
# dominant kernel with CUDA events on the launching stream inside the timed region).
conv_probe = None


class ConvProbe:
    """Records CUDA-event pairs around the conv launches whose (kind, N, Cin, Cout, D, H, W) is in
    `select`; `durations_ms(key)` after a synchronize."""

    def __init__(self, select):
        self.select = {tuple(k) for k in select}
        self.pairs = {k: [] for k in self.select}

    def start(self, kind, n, cin, cout, d, h, w):
        key = (kind, n, cin, cout, d, h, w)
        if key not in self.select:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        return key, e0

    def stop(self, tok):
        key, e0 = tok
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(torch.cuda.current_stream())
        self.pairs[key].append((e0, e1))

    def durations_ms(self, key):
        return [a.elapsed_time(b) for a, b in self.pairs[tuple(key)]]


def set_leaky_slope(slope: float) -> None:
    """Negative slope of every fused LeakyReLU and LeakyReLU-backward mask (process-wide; 0.2 = network.py,
    0.3 / 0.0 = network_dict.py's 'leaky_relu' / 'relu').  Synchronises the device when the value changes."""
    lib = _lib.load()
    rc = lib.sg_set_leaky_slope(float(slope))
    if rc != 0:
        msg = lib.sg_last_error()
        raise RuntimeError(f"sg_set_leaky_slope failed (status {rc}): {msg.decode() if msg else ''}")


def get_leaky_slope() -> float:
    return float(_lib.load().sg_get_leaky_slope())


def ensure_leaky_slope(want: float) -> None:
    """Set the library's slope if it differs (a float compare on the host when it does not)."""
    if abs(get_leaky_slope() - want) > 1e-7:
        set_leaky_slope(want)


def _vox(t: torch.Tensor) -> int:
    return t.shape[2] * t.shape[3] * t.shape[4]


# ----------------------------------------------------------------------------- layout
def plain_to_act(plain: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    n, c, d, h, w = plain.shape
    out = torch.empty((n, chunks(c), d, h, w, 8), dtype=dtype, device=plain.device)
    call("sg_plain_to_act", plain, out, _lib.dtype_code(out), n, c, d * h * w)
    return out


def act_to_plain(act: torch.Tensor, c: int) -> torch.Tensor:
    n, cc, d, h, w, _ = act.shape
    out = torch.empty((n, c, d, h, w), dtype=torch.float32, device=act.device)
    call("sg_act_to_plain", act, out, _lib.dtype_code(act), n, c, d * h * w)
    return out


# ------------------------------------------------------------------------------- conv
def _pack_kind(kind):
    """packed-weight kind -> (buffer dtype, ABI dtype code): torch.bfloat16, torch.float32, or "tf32" (fp32 values
    rounded to tf32 in the 4-channel-chunk layout of the kind::tf32 kernels)"""
    if kind == "tf32":
        return torch.float32, _lib.TF32_PACK
    return kind, (_lib.BF16 if kind == torch.bfloat16 else _lib.F32)


def pack_conv_weight(w: torch.Tensor, kind, flip: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`out`: a buffer from an earlier call for the same parameter (re-packed in place after an optimiser step)."""
    cout, cin = w.shape[0], w.shape[1]
    dtype, code = _pack_kind(kind)
    if out is None:
        out = torch.empty(_lib.packed_weight_elems(cout, cin, int(flip)), dtype=dtype, device=w.device)
    call("sg_pack_conv_weight", w, out, code, cout, cin, int(flip))
    return out


def conv_tf32_supported(n, cin, cout, d, h, w) -> bool:
    return bool(_lib.load().sg_conv3d_tf32_supported(n, cin, cout, d, h, w))


_pack_tables = {}


def pack_conv_weights_multi(jobs) -> list:
    """jobs: [(w, dtype, flip, out or None)] -> the packed buffers, all written by ONE launch
    (sg_pack_conv_weights_multi).  The job / block tables live on the device, cached on the pointers involved; a new
    table cannot be built while a CUDA graph is being captured (pinned staging), then the jobs run one by one."""
    outs = []
    for w, kind, flip, out in jobs:
        if out is None:
            out = torch.empty(_lib.packed_weight_elems(w.shape[0], w.shape[1], int(flip)), dtype=_pack_kind(kind)[0],
                              device=w.device)
        outs.append(out)
    key = tuple((w.data_ptr(), o.data_ptr(), int(flip), str(kind)) for (w, kind, flip, _), o in zip(jobs, outs))
    tab = _pack_tables.get(key)
    if tab is None:
        if torch.cuda.is_current_stream_capturing():
            for (w, dtype, flip, _), o in zip(jobs, outs):
                pack_conv_weight(w, dtype, flip, o)
            return outs
        rows, bj, bk, br = [], [], [], []
        for ji, ((w, dtype, flip, _), o) in enumerate(zip(jobs, outs)):
            cout, cin = w.shape[0], w.shape[1]
            k, r = (cout, cin) if flip else (cin, cout)
            # struct SgPackJob { const float* w; void* dst; int Cout, Cin, flip, dtype; } as four int64 words
            rows.append([w.data_ptr(), o.data_ptr(), cout | (cin << 32), int(flip) | (_pack_kind(dtype)[1] << 32)])
            for kc in range(chunks(k)):
                for r0 in range(0, 16 * ((r + 15) // 16), 32):
                    bj.append(ji), bk.append(kc), br.append(r0)
        dev = outs[0].device
        tab = dict(jobs=torch.tensor(rows, dtype=torch.int64).to(dev), n=len(bj),
                   bj=torch.tensor(bj, dtype=torch.int32).to(dev), bk=torch.tensor(bk, dtype=torch.int32).to(dev),
                   br=torch.tensor(br, dtype=torch.int32).to(dev))
        if len(_pack_tables) > 64:
            _pack_tables.clear()
        _pack_tables[key] = tab
    call("sg_pack_conv_weights_multi", tab["jobs"], tab["bj"], tab["bk"], tab["br"], tab["n"])
    return outs


def _workspace(kind: int, x: torch.Tensor, n, cin, cout, d, h, w):
    """Caller-owned scratch for the conv entry points (the library never allocates)."""
    nbytes = _lib.conv_workspace_bytes(kind, _lib.dtype_code(x), n, cin, cout, d, h, w)
    if nbytes == 0:
        return None, 0
    return torch.empty((nbytes,), dtype=torch.uint8, device=x.device), nbytes


def conv3d_fprop(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor],
                 mask_src: Optional[torch.Tensor], cin: int, cout: int, scale: float, lrelu: bool,
                 impl: int = _lib.IMPL_AUTO) -> torch.Tensor:
    n, cc, d, h, w, _ = x.shape
    assert cc == chunks(cin), (cc, cin)
    y = torch.empty((n, chunks(cout), d, h, w, 8), dtype=x.dtype, device=x.device)
    ws, nbytes = _workspace(0, x, n, cin, cout, d, h, w)
    probe = conv_probe
    tok = probe.start("fprop", n, cin, cout, d, h, w) if probe is not None else None
    call("sg_conv3d_fprop", x, wp, bias, mask_src, y, _lib.dtype_code(x), n, cin, cout, d, h, w,
         float(scale), int(lrelu), impl, ws, nbytes)
    if tok is not None:
        probe.stop(tok)
    return y


def conv_pixelnorm_supported(x: torch.Tensor, cin: int, cout: int) -> bool:
    """True when conv3d_fprop_pixelnorm covers this call (bf16, weight-resident kernel, all channels in one N tile)."""
    if x.dtype != torch.bfloat16:
        return False
    n, _, d, h, w, _ = x.shape
    return bool(_lib.load().sg_conv3d_pixelnorm_supported(n, cin, cout, d, h, w))


def conv3d_fprop_pixelnorm(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], cin: int, cout: int,
                           scale: float, lrelu: bool, lrelu_after: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """(y, y_norm): y = [lrelu](scale*conv(x) + bias), y_norm = [lrelu_after](pixel-norm(y)) from one kernel."""
    n, cc, d, h, w, _ = x.shape
    y = torch.empty((n, chunks(cout), d, h, w, 8), dtype=x.dtype, device=x.device)
    y_norm = torch.empty_like(y)
    call("sg_conv3d_fprop_pixelnorm", x, wp, bias, y, y_norm, _lib.dtype_code(x), n, cin, cout, d, h, w, float(scale),
         int(lrelu), int(lrelu_after), EPS_PN)
    return y, y_norm


def conv_pool_supported(x: torch.Tensor, cin: int, cout: int) -> bool:
    """True when conv3d_fprop_pool covers this call (bf16 in and out, weight-resident kernel, even planes per tile)."""
    if x.dtype != torch.bfloat16:
        return False
    n, _, d, h, w, _ = x.shape
    return bool(_lib.load().sg_conv3d_pool_supported(n, cin, cout, d, h, w))


def conv3d_fprop_pool(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], cin: int, cout: int, scale: float,
                      lrelu: bool, pool_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(y, y_pool): y = [lrelu](scale*conv(x) + bias), y_pool = pool_scale * (2x2x2 block sums of y) from one kernel."""
    n, cc, d, h, w, _ = x.shape
    y = torch.empty((n, chunks(cout), d, h, w, 8), dtype=x.dtype, device=x.device)
    y_pool = torch.empty((n, chunks(cout), d // 2, h // 2, w // 2, 8), dtype=x.dtype, device=x.device)
    call("sg_conv3d_fprop_pool", x, wp, bias, y, y_pool, _lib.dtype_code(x), n, cin, cout, d, h, w, float(scale), int(lrelu),
         float(pool_scale))
    return y, y_pool


def conv3d_wgrad(x: torch.Tensor, gy: torch.Tensor, cin: int, cout: int, scale: float,
                 want_bias: bool, impl: int = _lib.IMPL_AUTO
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    n, cc, d, h, w, _ = x.shape
    assert cc == chunks(cin) and gy.shape[1] == chunks(cout)
    gw = torch.empty((cout, cin, 3, 3, 3), dtype=torch.float32, device=x.device)
    gb = torch.empty((cout,), dtype=torch.float32, device=x.device) if want_bias else None
    ws, nbytes = _workspace(1, x, n, cin, cout, d, h, w)
    probe = conv_probe
    tok = probe.start("wgrad", n, cin, cout, d, h, w) if probe is not None else None
    call("sg_conv3d_wgrad", x, gy, gw, gb, _lib.dtype_code(x), n, cin, cout, d, h, w, float(scale), impl,
         ws, nbytes)
    if tok is not None:
        probe.stop(tok)
    return gw, gb


# ----------------------------------------------------------------------------- 1x1x1
def pw_expand(img: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], dtype: torch.dtype,
              c: int, scale: float, lrelu: bool, mask_ref: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, _, d, h, wd = img.shape
    y = torch.empty((n, chunks(c), d, h, wd, 8), dtype=dtype, device=img.device)
    if mask_ref is None:
        call("sg_pw_expand", img, w, bias, y, _lib.dtype_code(y), n, c, d * h * wd, float(scale), int(lrelu))
    else:
        assert mask_ref.shape == y.shape and mask_ref.dtype == y.dtype
        call("sg_pw_expand_masked", img, w, bias, mask_ref, y, _lib.dtype_code(y), n, c, d * h * wd, float(scale),
             int(lrelu))
    return y


def pw_reduce(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], c: int,
              scale: float) -> torch.Tensor:
    n, cc, d, h, wd, _ = x.shape
    img = torch.empty((n, 1, d, h, wd), dtype=torch.float32, device=x.device)
    call("sg_pw_reduce", x, w, bias, img, _lib.dtype_code(x), n, c, d * h * wd, float(scale))
    return img


def pw_wgrad(g: torch.Tensor, img: Optional[torch.Tensor], c: int, scale: float, want_w: bool,
             want_b: bool) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    n, cc, d, h, wd, _ = g.shape
    gw = torch.empty((c,), dtype=torch.float32, device=g.device) if want_w else None
    gb = torch.empty((c,), dtype=torch.float32, device=g.device) if want_b else None
    call("sg_pw_wgrad", g, img if want_w else None, gw, gb, _lib.dtype_code(g), n, c, d * h * wd,
         float(scale))
    return gw, gb


# ------------------------------------------------------------------------- resampling
def _resample(name: str, x: torch.Tensor, scale: float, factor_num: int, factor_den: int,
              out_dtype: Optional[torch.dtype] = None, mask_ref: Optional[torch.Tensor] = None):
    if x.dim() == 6:
        n, cc, d, h, w, _ = x.shape
        shape = (n, cc, d * factor_num // factor_den, h * factor_num // factor_den,
                 w * factor_num // factor_den, 8)
        vec, planes = 8, n * cc
    else:
        n, c1, d, h, w = x.shape
        assert x.dtype == torch.float32
        shape = (n, c1, d * factor_num // factor_den, h * factor_num // factor_den,
                 w * factor_num // factor_den)
        vec, planes = 1, n * c1
    y = torch.empty(shape, dtype=out_dtype or x.dtype, device=x.device)
    if name == "sg_up2":
        assert mask_ref is None or (mask_ref.shape == y.shape and mask_ref.dtype == y.dtype)
        call(name, x, y, mask_ref, _lib.dtype_code(x), _lib.dtype_code(y), vec, planes, d, h, w, float(scale))
    else:
        call(name, x, y, _lib.dtype_code(x), _lib.dtype_code(y), vec, planes, d, h, w, float(scale))
    return y


def down2(x: torch.Tensor, scale: float, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    return _resample("sg_down2", x, scale, 1, 2, out_dtype)


def up2(x: torch.Tensor, scale: float, out_dtype: Optional[torch.dtype] = None,
        mask_ref: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _resample("sg_up2", x, scale, 2, 1, out_dtype, mask_ref)


# ------------------------------------------------------------------------ elementwise
def lincomb(a: torch.Tensor, b: Optional[torch.Tensor], alpha, beta=None) -> torch.Tensor:
    """y = alpha*a + beta*b.  alpha, beta: Python floats, or `alpha` a 2-element fp32 DEVICE tensor {alpha, beta}
    (read by the kernel: no host sync, and a captured graph follows the values)."""
    y = torch.empty_like(a)
    if isinstance(alpha, torch.Tensor):
        assert alpha.dtype == torch.float32 and alpha.numel() == 2 and alpha.device == a.device
        call("sg_lincomb_dev", a, b, y, _lib.dtype_code(a), a.numel(), alpha)
    else:
        call("sg_lincomb", a, b, y, _lib.dtype_code(a), a.numel(), float(alpha), float(beta))
    return y


def lrelu_fwd(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    call("sg_lrelu_fwd", x, y, _lib.dtype_code(x), x.numel())
    return y


def mask_mul(g: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(g)
    call("sg_mask_mul", g, ref, y, _lib.dtype_code(g), g.numel())
    return y


def pixelnorm_fwd(x: torch.Tensor, c: int, lrelu_after: bool) -> torch.Tensor:
    y = torch.empty_like(x)
    call("sg_pixelnorm_fwd", x, y, _lib.dtype_code(x), x.shape[0], c, _vox(x), EPS_PN, int(lrelu_after))
    return y


def pixelnorm_bwd(x: torch.Tensor, gy: torch.Tensor, c: int, lrelu_after: bool,
                  mask_input: bool = False) -> torch.Tensor:
    gx = torch.empty_like(x)
    call("sg_pixelnorm_bwd", x, gy, gx, _lib.dtype_code(x), x.shape[0], c, _vox(x), EPS_PN,
         int(lrelu_after), int(mask_input))
    return gx


# ------------------------------------------------------------------- gradient penalty
def interp(real: torch.Tensor, fake: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(real)
    n = real.shape[0]
    call("sg_interp", real, fake, eps, out, n, real.numel() // max(n, 1))
    return out


def sumsq_rows(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=x.device)
    call("sg_sumsq_rows", x, out, n, x.numel() // max(n, 1))
    return out


def rowscale(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    n = x.shape[0]
    call("sg_rowscale", x, s, y, n, x.numel() // max(n, 1))
    return y


# ----------------------------------------------------------------------------- linear
def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], scale: float,
               lrelu: bool) -> torch.Tensor:
    b, fin = x.shape
    y = torch.empty((b, w.shape[0]), dtype=torch.float32, device=x.device)
    call("sg_linear_fwd", x, w, bias, y, b, fin, w.shape[0], float(scale), int(lrelu))
    return y


def linear_dgrad(g: torch.Tensor, w: torch.Tensor, scale: float) -> torch.Tensor:
    b, fout = g.shape
    gx = torch.empty((b, w.shape[1]), dtype=torch.float32, device=g.device)
    call("sg_linear_dgrad", g, w, gx, b, w.shape[1], fout, float(scale))
    return gx


def linear_wgrad(g: torch.Tensor, x: torch.Tensor, scale: float, want_bias: bool
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    b, fout = g.shape
    fin = x.shape[1]
    gw = torch.empty((fout, fin), dtype=torch.float32, device=g.device)
    gb = torch.empty((fout,), dtype=torch.float32, device=g.device) if want_bias else None
    call("sg_linear_wgrad", g, x, gw, gb, b, fin, fout, float(scale))
    return gw, gb


# ------------------------------------------------------------------ minibatch stddev
def mbstd_fwd(x: torch.Tensor, group: int, sub_batches: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """x plain fp32 (B,C,D,H,W) -> out (B,C+1,D,H,W) = cat(group-centred x, stat channel), s (S*M, C*V).
    B = sub_batches * group * M: `sub_batches` independent minibatches stacked along the batch axis."""
    b, c, d, h, w = x.shape
    m, v = b // (group * sub_batches), d * h * w
    out = torch.empty((b, c + 1, d, h, w), dtype=torch.float32, device=x.device)
    s = torch.empty((sub_batches * m, c * v), dtype=torch.float32, device=x.device)
    t = torch.empty((sub_batches * m,), dtype=torch.float32, device=x.device)
    call("sg_mbstd_fwd", x, out, s, t, sub_batches, group, m, c, v, 1e-8)
    return out, s


def mbstd_bwd(gout: torch.Tensor, out: torch.Tensor, s: torch.Tensor, group: int, sub_batches: int = 1
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    b, c1, d, h, w = out.shape
    c, m, v = c1 - 1, b // (group * sub_batches), d * h * w
    gx = torch.empty((b, c, d, h, w), dtype=torch.float32, device=out.device)
    gt = torch.empty((sub_batches * m,), dtype=torch.float32, device=out.device)
    call("sg_mbstd_bwd", gout, out, s, gt, gx, sub_batches, group, m, c, v)
    return gx, gt


def mbstd_bwdbwd(u: torch.Tensor, gt: torch.Tensor, out: torch.Tensor, s: torch.Tensor, group: int,
                 sub_batches: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    b, c1, d, h, w = out.shape
    c, m, v = c1 - 1, b // (group * sub_batches), d * h * w
    d_gout = torch.empty_like(out)
    d_gt = torch.empty((sub_batches * m,), dtype=torch.float32, device=out.device)
    d_x = torch.empty((b, c, d, h, w), dtype=torch.float32, device=out.device)
    call("sg_mbstd_bwdbwd", u, gt, out, s, d_gout, d_gt, d_x, sub_batches, group, m, c, v)
    return d_gout, d_x


# ------------------------------------------------------------------ evaluation metrics
def pyr_down(x: torch.Tensor) -> torch.Tensor:
    """swd.py:61-63 on a (N, C, D, H, W) fp32 volume."""
    n, c, d, h, w = x.shape
    y = torch.empty((n, c, (d + 1) // 2, (h + 1) // 2, (w + 1) // 2), dtype=torch.float32, device=x.device)
    call("sg_pyr_down", x, y, n * c, d, h, w)
    return y


def pyr_up_sub(fine: torch.Tensor, coarse: torch.Tensor) -> torch.Tensor:
    """fine - pyr_up(coarse)  (swd.py:65-78)."""
    n, c, cd, ch, cw = coarse.shape
    if tuple(fine.shape) != (n, c, 2 * cd, 2 * ch, 2 * cw):
        raise ValueError(f"pyr_up_sub: {tuple(fine.shape)} is not twice {tuple(coarse.shape)} (odd extents have no "
                         "Laplacian level in the reference either: its subtraction cannot broadcast)")
    lap = torch.empty_like(fine)
    call("sg_pyr_up_sub", fine, coarse, lap, n * c, cd, ch, cw)
    return lap


def swd_descriptors(level: torch.Tensor, pos_z: torch.Tensor, pos_y: torch.Tensor, pos_x: torch.Tensor,
                    out: torch.Tensor) -> None:
    """Standardised 3x9x9 descriptors of `level` (B, 1, D, H, W) at the N positions into out[b, j*243 + e]
    (out: a (B, N*243) row-major view, e.g. one arm's rows of the stacked real/fake matrix)."""
    b, c, d, h, w = level.shape
    assert c == 1 and out.shape[0] == b and out.stride(1) == 1
    n = pos_z.numel()
    call("sg_swd_descriptors", level, pos_z, pos_y, pos_x, out, b, d, h, w, n, out.stride(0))


def swd_project(a: torch.Tensor, dirs: torch.Tensor, want_colsq: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """p = a @ dirs (a: (R, K), dirs: (K, 128)); colsq = squared column norms of dirs when asked for."""
    r, k = a.shape
    assert dirs.shape == (k, 128)
    p = torch.empty((r, 128), dtype=torch.float32, device=a.device)
    colsq = torch.empty(128, dtype=torch.float32, device=a.device) if want_colsq else None
    call("sg_swd_project", a, dirs, p, colsq, r, k, a.stride(0))
    return p, colsq


def swd_finish(p: torch.Tensor, colsq: Optional[torch.Tensor], out: torch.Tensor) -> None:
    """out[0] = mean over directions and sorted ranks of |real - fake| for one repeat (p: (2B, 128))."""
    call("sg_swd_finish", p, colsq, out, p.shape[0] // 2)


def value_hist(x: torch.Tensor, intercept: float, lo: int, hi: int) -> torch.Tensor:
    """Per-image counts of clip(int(x*intercept + intercept), lo, hi): (N, hi-lo+1) int32 (kms.py:6-10)."""
    n = x.shape[0]
    v = x[0].numel() if n else 0
    hist = torch.empty((n, hi - lo + 1), dtype=torch.int32, device=x.device)
    call("sg_value_hist", x, hist, n, v, float(intercept), lo, hi)
    return hist
