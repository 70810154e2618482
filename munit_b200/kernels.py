"""Thin torch-tensor wrappers over the C-ABI (munit_b200/_lib.py).  PyTorch supplies device memory and
the current CUDA stream only; every call below launches hand-written sm_100a kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT, NORM, TapGemmDesc, WgradDesc, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _f16(t) -> int:
    """1 when a raw conv output is stored as fp16 (ops.ConvFn in front of a norm), 0 for bf16."""
    assert t.dtype in (torch.bfloat16, torch.float16)
    return int(t.dtype == torch.float16)


def _count(n=1):
    _lib.launches += n


# Profiling aid (never set in production): MUNIT_SKIP="family,family" makes those kernel families no-ops so the
# step-time delta measures their true in-graph cost (results are garbage while it is set).
import os as _os

_SKIP = set(filter(None, _os.environ.get("MUNIT_SKIP", "").split(",")))
if _SKIP:
    import ctypes as _ct

    class _SkipLib:
        def __init__(self, real):
            self._real = real

        def __getattr__(self, name):
            fn = getattr(self._real, name)
            if name.replace("munit_", "") in _SKIP:
                return lambda *a, **kw: 0
            return fn

    lib = _SkipLib(lib)


def _fill5(dst, src, fill=0):
    for i in range(5):
        dst[i] = src[i] if i < len(src) else fill


def stats_splits(plan, kind: int) -> int:
    """Partials per sample the conv epilogue writes for `plan` (include/munit_b200.h, munit_tapgemm_desc.stats),
    or 0 when the plan cannot produce statistics (partial tiles, several images per tile, phases, bn < 64)."""
    if (plan.bn < 64 or plan.phases != 1 or plan.tn != 1 or plan.out_w % plan.tw or plan.out_h % plan.th
            or plan.n_store < plan.b_rows or plan.b_rows % 64):
        return 0
    tiles = (plan.out_w // plan.tw) * (plan.out_h // plan.th)
    return 4 * tiles * (plan.b_rows // 64 if kind == 2 else 1)


def auto_ksplit(plan) -> int:
    """Split-K factor for launches whose output tiles cannot fill the GPU (the deepest discriminator layers:
    <= 8 CTAs each walking >= 64 K blocks one after the other).  1 = no split.  Wider thresholds were measured
    slower in the two-stream step: the memset + finish launches and the fp32 atomics cost more than the short
    kernels, which already overlap with the other branch (38.39 ms unsplit, 39.33 ms at <= 48 CTAs / >= 16 blocks,
    37.92 ms at <= 8 / >= 64)."""
    if getattr(plan, "halo", 0) or not _AUTO_KSPLIT:
        return 1
    gx, gy, gz = plan.grid
    ctas = gx * gy * gz
    num_kb = plan.num_taps * plan.chunks
    if ctas == 0 or ctas > _KSPLIT_MAX_CTAS or num_kb < _KSPLIT_MIN_KB:
        return 1
    return max(1, min(num_kb // 4, 16, 296 // ctas))


# CTA-pair (cta_group::2) tap-GEMM for the bn 128 / 256 launches (include/munit_b200.h, munit_tapgemm_desc.pair)
PAIR = _os.environ.get("MUNIT_PAIR", "1") != "0"
_AUTO_KSPLIT = _os.environ.get("MUNIT_KSPLIT", "1") != "0"
_KSPLIT_MAX_CTAS = int(_os.environ.get("MUNIT_KSPLIT_CTAS", "8"))
_KSPLIT_MIN_KB = int(_os.environ.get("MUNIT_KSPLIT_KB", "64"))


def tapgemm(plan, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, bias=None, act="none", stages=0,
            stats=None, stats_kind=0, ksplit=0, pair=None):
    """out (bf16) <- act(tapconv(a; b) + bias) as described by `plan` (geometry.TapGemmPlan).  `stats` (fp32,
    n_img * stats_splits(plan, kind) * (C if kind == 1 else 1) * 2 floats) receives the norm partials."""
    _lib.init()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and out.dtype in (torch.bfloat16, torch.float16)
    assert a.is_cuda and a.is_contiguous() and b.is_contiguous()
    assert b.numel() == plan.b_rows * plan.b_k, (b.shape, plan.b_rows, plan.b_k)
    d = TapGemmDesc()
    d.a, d.a_rank = a.data_ptr(), plan.a_rank
    _fill5(d.a_dim, plan.a_dim, 1)
    _fill5(d.a_stride, plan.a_stride, 0)
    _fill5(d.a_box, plan.a_box, 1)
    d.b, d.b_rows, d.b_k, d.bn = b.data_ptr(), plan.b_rows, plan.b_k, plan.bn
    d.tw, d.th, d.tn = plan.tw, plan.th, plan.tn
    d.out_w, d.out_h, d.n_img = plan.out_w, plan.out_h, plan.n_img
    _fill5(d.mx, plan.mx)
    _fill5(d.my, plan.my)
    _fill5(d.mn, plan.mn)
    d.num_taps, d.chunks = plan.num_taps, plan.chunks
    for t, off in enumerate(plan.tap_off):
        for i in range(5):
            d.tap_off[t][i] = off[i] if i < len(off) else 0
    d.phases = plan.phases
    for i in range(plan.phases):
        d.b_k0[i], d.o_yoff[i], d.o_xoff[i] = plan.b_k0[i], plan.o_yoff[i], plan.o_xoff[i]
    d.out = out.data_ptr()
    d.o_sn, d.o_sy, d.o_sx, d.o_ymul, d.o_xmul = plan.o_sn, plan.o_sy, plan.o_sx, plan.o_ymul, plan.o_xmul
    d.n_store = plan.n_store
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= plan.b_rows
    d.bias, d.act, d.stages = _ptr(bias), ACT[act], stages
    d.out_f16 = _f16(out)
    d.pair = int(PAIR if pair is None else pair)
    d.halo = int(getattr(plan, "halo", 0))
    ksplit = ksplit or (auto_ksplit(plan) if stats is None else 1)
    scratch = None
    if ksplit > 1:
        assert out.is_contiguous() and plan.n_store == plan.b_rows and out.shape[-1] == plan.b_rows
        scratch = torch.zeros(out.shape, dtype=torch.float32, device=out.device)
        d.ksplit, d.scratch = ksplit, scratch.data_ptr()
    if stats is not None:
        assert stats.dtype == torch.float32 and stats_kind in (1, 2)
        assert stats.numel() == plan.n_img * stats_splits(plan, stats_kind) * (plan.b_rows if stats_kind == 1 else 1) * 2
        d.stats, d.stats_kind = stats.data_ptr(), stats_kind
    check(lib.munit_tapgemm(C.byref(d), _stream()), "munit_tapgemm")
    _count()
    if scratch is not None:
        check(lib.munit_splitk_finish(scratch.data_ptr(), _ptr(bias), ACT[act], out.data_ptr(), _f16(out), out.numel(),
                                      out.shape[-1], _stream()), "splitk_finish")
        _count()
    return out


def wgrad(plan, dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, ksplit=0, stages=0):
    """dw (fp32, +=) <- sum_pix dy[pix, co] * x[pix@tap, ci] as described by `plan` (geometry.WgradPlan).
    A swapped plan (tap_on_a) feeds x as the 128-row A operand and dy as B."""
    _lib.init()
    a, b = (x, dy) if getattr(plan, "tap_on_a", 0) else (dy, x)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and dw.dtype == torch.float32
    d = WgradDesc()
    d.a, d.a_rank = a.data_ptr(), plan.a_rank
    _fill5(d.a_dim, plan.a_dim, 1)
    _fill5(d.a_stride, plan.a_stride, 0)
    _fill5(d.a_box, plan.a_box, 1)
    _fill5(d.a_mx, plan.a_mx)
    _fill5(d.a_my, plan.a_my)
    _fill5(d.a_mn, plan.a_mn)
    d.b, d.b_rank = b.data_ptr(), plan.b_rank
    _fill5(d.b_dim, plan.b_dim, 1)
    _fill5(d.b_stride, plan.b_stride, 0)
    _fill5(d.b_box, plan.b_box, 1)
    _fill5(d.b_mx, plan.b_mx)
    _fill5(d.b_my, plan.b_my)
    _fill5(d.b_mn, plan.b_mn)
    d.pw, d.ph, d.pn = plan.pw, plan.ph, plan.pn
    d.out_w, d.out_h, d.n_img = plan.out_w, plan.out_h, plan.n_img
    d.m_total, d.n_total, d.bn, d.num_taps = plan.m_total, plan.n_total, plan.bn, plan.num_taps
    for t, off in enumerate(plan.tap_off):
        for i in range(5):
            d.tap_off[t][i] = off[i] if i < len(off) else 0
    d.dw, d.s_m, d.s_t, d.s_n = dw.data_ptr(), plan.s_m, plan.s_t, plan.s_n
    d.ksplit, d.stages, d.tap_on_a = ksplit, stages, int(getattr(plan, "tap_on_a", 0))
    check(lib.munit_wgrad(C.byref(d), _stream()), "munit_wgrad")
    _count()
    return dw


# ---------------------------------------------------------------- layout
def image_to_act(x, pad, cp):
    n, c, h, w = x.shape
    act = torch.empty(n, h + 2 * pad, w + 2 * pad, cp, dtype=torch.bfloat16, device=x.device)
    check(lib.munit_image_to_act(x.data_ptr(), act.data_ptr(), n, c, h, w, pad, cp, _stream()), "image_to_act")
    _count()
    return act


def image_to_kwexp(x, pad, kw, sx, kwp, cp):
    n, c, h, w = x.shape
    wo = (w + 2 * pad - kw) // sx + 1
    e = torch.empty(n, h + 2 * pad, wo, kwp * cp, dtype=torch.bfloat16, device=x.device)
    check(lib.munit_image_to_kwexp(x.data_ptr(), e.data_ptr(), n, c, h, w, pad, kw, sx, wo, kwp, cp, _stream()),
          "image_to_kwexp")
    _count()
    return e


def kwexp_to_image_grad(de, n, c, h, w, pad, kw, sx, kwp, cp):
    wo = (w + 2 * pad - kw) // sx + 1
    dx = torch.empty(n, c, h, w, dtype=torch.float32, device=de.device)
    check(lib.munit_kwexp_to_image_grad(de.data_ptr(), dx.data_ptr(), n, c, h, w, pad, kw, sx, wo, kwp, cp, _stream()),
          "kwexp_to_image_grad")
    _count()
    return dx


def u8_crop_normalize(img_u8, top, left, flip, out):
    """img_u8 [H, W, 3] uint8 (device) -> out [3, ch, cw] fp32 in [-1, 1]: crop at (top, left), optional mirror,
    ToTensor + Normalize(0.5, 0.5) (utils.py:218-240)."""
    _lib.init()
    assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.is_contiguous() and img_u8.shape[2] == 3
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == 3
    ih, iw, _ = img_u8.shape
    _, ch, cw = out.shape
    check(lib.munit_u8_crop_normalize(img_u8.data_ptr(), ih, iw, int(top), int(left), int(bool(flip)), out.data_ptr(), ch,
                                      cw, _stream()), "u8_crop_normalize")
    _count()
    return out


def act_to_nchw(act, c, pad):
    n, hp, wp, cp = act.shape
    h, w = hp - 2 * pad, wp - 2 * pad
    y = torch.empty(n, c, h, w, dtype=torch.float32, device=act.device)
    check(lib.munit_act_to_nchw(act.data_ptr(), y.data_ptr(), n, c, h, w, pad, cp, _stream()), "act_to_nchw")
    _count()
    return y


def nchw_to_act(x, pad, cp=None, fill_halo=True):
    n, c, h, w = x.shape
    cp = cp or c
    act = torch.empty(n, h + 2 * pad, w + 2 * pad, cp, dtype=torch.bfloat16, device=x.device)
    check(lib.munit_nchw_to_act(x.data_ptr(), act.data_ptr(), n, c, h, w, pad, cp, _stream()), "nchw_to_act")
    _count()
    if fill_halo and pad:
        halo_fill(act, pad)
    return act


def halo_fill(act, pad):
    n, hp, wp, c = act.shape
    if pad:
        check(lib.munit_halo_fill(act.data_ptr(), n, hp - 2 * pad, wp - 2 * pad, c, pad, _stream()), "halo_fill")
        _count()
    return act


# ---------------------------------------------------------------- norms
def norm_stats(y):
    n, h, w, c = y.shape
    splits = lib.munit_norm_splits(h * w, c)
    assert splits > 0, f"unsupported channel count {c}"
    stats = torch.empty(n, splits, c, 2, dtype=torch.float32, device=y.device)
    shift = torch.empty(n, c, dtype=torch.float32, device=y.device)
    check(lib.munit_norm_stats(y.data_ptr(), _f16(y), stats.data_ptr(), shift.data_ptr(), n, h * w, c, _stream()), "norm_stats")
    _count()
    return stats, shift


def norm_finalize(stats, shift, mode, p_w, p_b, ldw, hw, eps=1e-5):
    n, c = shift.shape
    o = torch.empty(4, n, c, dtype=torch.float32, device=stats.device)
    check(lib.munit_norm_finalize(stats.data_ptr(), shift.data_ptr(), NORM[mode], _ptr(p_w), _ptr(p_b), ldw, eps,
                                  o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), n, hw, c,
                                  _stream()), "norm_finalize")
    _count()
    return o  # mean, rinv, a, b


def norm_apply(y, a, b, relu, residual, res_pad, out_pad, upsample):
    n, h, w, c = y.shape
    out = torch.empty(n, h * upsample + 2 * out_pad, w * upsample + 2 * out_pad, c, dtype=torch.bfloat16,
                      device=y.device)
    check(lib.munit_norm_apply(y.data_ptr(), _f16(y), a.data_ptr(), b.data_ptr(), int(relu), _ptr(residual), res_pad,
                               out.data_ptr(), out_pad, upsample, n, h, w, c, _stream()), "norm_apply")
    _count()
    return out


def norm_finalize_parts(part, kind, mode, p_w, p_b, ldw, n, hw, c, eps=1e-5):
    """coef (mean, rinv, a, b) from the partials a conv epilogue left (tapgemm(..., stats=...))."""
    o = torch.empty(4, n, c, dtype=torch.float32, device=part.device)
    splits = part.numel() // (n * 2 * (c if kind == 1 else 1))
    check(lib.munit_norm_finalize_parts(part.data_ptr(), splits, kind, NORM[mode], _ptr(p_w), _ptr(p_b), ldw, eps,
                                        o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), n, hw, c,
                                        _stream()), "norm_finalize_parts")
    _count()
    return o


_TICKETS = {}


def _tickets(n, device):
    """n zero int32 tickets for one last-block-done launch (munit_norm_stats_finalize / _bwd_reduce_finalize).  Taken
    from a rotating pool: launches that may be in flight at the same time (other streams, later nodes of a captured
    graph) hold different slots; the kernels return their tickets to zero."""
    key = (device.type, device.index)
    st = _TICKETS.get(key)
    if st is None:
        st = _TICKETS[key] = [torch.zeros(1 << 16, dtype=torch.int32, device=device), 0]
    buf, pos = st
    if pos + n > buf.numel():
        pos = 0
    st[1] = pos + n
    return buf[pos:pos + n]


# Statistics + finalize (and backward reduce + finalize) as ONE launch each, the finalize running in the last block of
# every sample.  Measured on B200 inside the captured step: 37.1 ms/step against 37.0 with the separate ~2 us finalize
# launches (the serial tail of the last block costs what the removed launch boundary saves), so the separate launches
# stay the default; MUNIT_NORM_LASTBLOCK=1 selects the fused form (same results bit for bit, tests/test_simt_gpu.py).
FUSE_FINALIZE = _os.environ.get("MUNIT_NORM_LASTBLOCK", "0") != "0"


def norm_stats_finalize(y, mode, p_w, p_b, ldw, eps=1e-5):
    """coef[4][N][C] = (mean, rinv, a, b) of y in ONE launch (statistics, then the finalize in each sample's last block)."""
    n, h, w, c = y.shape
    splits = lib.munit_norm_splits(h * w, c)
    assert splits > 0, f"unsupported channel count {c}"
    stats = torch.empty(n, splits, c, 2, dtype=torch.float32, device=y.device)
    shift = torch.empty(n, c, dtype=torch.float32, device=y.device)
    o = torch.empty(4, n, c, dtype=torch.float32, device=y.device)
    check(lib.munit_norm_stats_finalize(y.data_ptr(), _f16(y), stats.data_ptr(), shift.data_ptr(),
                                        _tickets(n, y.device).data_ptr(), NORM[mode], _ptr(p_w), _ptr(p_b), ldw, eps,
                                        o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), n, h * w, c,
                                        _stream()), "norm_stats_finalize")
    _count()
    return o


def norm_fwd(y, mode, p_w, p_b, ldw, eps, relu, residual, res_pad, out_pad, upsample, part=None):
    """Normalise + affine + ReLU + residual + upsample + halo.  Returns (out_act, coef[4][N][C]) with
    coef = (mean, rinv, a, b).  `part`: statistics partials from the producing conv's epilogue (skips the
    statistics pass over y)."""
    n, h, w, c = y.shape
    if part is not None and part.numel():
        coef = norm_finalize_parts(part, 2 if mode == "ln" else 1, mode, p_w, p_b, ldw, n, h * w, c, eps)
    elif FUSE_FINALIZE:
        coef = norm_stats_finalize(y, mode, p_w, p_b, ldw, eps)
    else:
        stats, shift = norm_stats(y)
        coef = norm_finalize(stats, shift, mode, p_w, p_b, ldw, h * w, eps)
    return norm_apply(y, coef[2], coef[3], relu, residual, res_pad, out_pad, upsample), coef


def norm_bwd_reduce(g_out, out_pad, upsample, y, coef, relu):
    """sums [N][S][C][2] = split partials of {sum dz, sum dz*(y - mean)}, dz = fold(g_out) * relu'(a*y + b)."""
    n, h, w, c = y.shape
    mean, rinv, a, b = coef[0], coef[1], coef[2], coef[3]
    sums = torch.empty(n, lib.munit_norm_splits(h * w, c), c, 2, dtype=torch.float32, device=y.device)
    check(lib.munit_norm_bwd_reduce(g_out.data_ptr(), out_pad, upsample, y.data_ptr(), _f16(y), a.data_ptr(), b.data_ptr(),
                                    int(relu), mean.data_ptr(), rinv.data_ptr(), sums.data_ptr(), n, h, w, c,
                                    _stream()), "norm_bwd_reduce")
    _count()
    return sums


def norm_bwd_finalize(sums, coef, mode, p_w, ldw, g_w, g_b, ldg, hw, eps=1e-5):
    """k[3][N][C] = (ca, cb, cc) with dx = ca*dz + cb*xhat + cc; AdaIN / LayerNorm parameter gradients."""
    n, _, c, _ = sums.shape
    k = torch.empty(3, n, c, dtype=torch.float32, device=sums.device)
    check(lib.munit_norm_bwd_finalize(sums.data_ptr(), NORM[mode], _ptr(p_w), ldw, coef[1].data_ptr(), eps,
                                      k[0].data_ptr(), k[1].data_ptr(), k[2].data_ptr(), _ptr(g_w), _ptr(g_b), ldg, n,
                                      hw, c, _stream()), "norm_bwd_finalize")
    _count()
    return k


def norm_bwd_apply(g_out, out_pad, upsample, y, coef, relu, k, want_res, res_pad):
    n, h, w, c = y.shape
    mean, rinv, a, b = coef[0], coef[1], coef[2], coef[3]
    dy = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device)
    g_res = None
    if want_res:
        g_res = torch.empty(n, h + 2 * res_pad, w + 2 * res_pad, c, dtype=torch.bfloat16, device=y.device)  # halo zeroed by the kernel
    check(lib.munit_norm_bwd_apply(g_out.data_ptr(), out_pad, upsample, y.data_ptr(), _f16(y), a.data_ptr(), b.data_ptr(),
                                   int(relu), mean.data_ptr(), rinv.data_ptr(), k[0].data_ptr(), k[1].data_ptr(),
                                   k[2].data_ptr(), dy.data_ptr(), _ptr(g_res), res_pad, n, h, w, c, _stream()),
          "norm_bwd_apply")
    _count()
    return dy, g_res


def norm_bwd_reduce_finalize(g_out, out_pad, upsample, y, coef, relu, mode, p_w, ldw, g_w, g_b, ldg, eps=1e-5):
    """k[3][N][C] = (ca, cb, cc) in ONE launch (reduce, then the finalize in each sample's last block)."""
    n, h, w, c = y.shape
    mean, rinv, a, b = coef[0], coef[1], coef[2], coef[3]
    sums = torch.empty(n, lib.munit_norm_splits(h * w, c), c, 2, dtype=torch.float32, device=y.device)
    k = torch.empty(3, n, c, dtype=torch.float32, device=y.device)
    check(lib.munit_norm_bwd_reduce_finalize(g_out.data_ptr(), out_pad, upsample, y.data_ptr(), _f16(y), a.data_ptr(),
                                             b.data_ptr(), int(relu), mean.data_ptr(), rinv.data_ptr(), sums.data_ptr(),
                                             _tickets(n, y.device).data_ptr(), NORM[mode], _ptr(p_w), ldw, eps,
                                             k[0].data_ptr(), k[1].data_ptr(), k[2].data_ptr(), _ptr(g_w), _ptr(g_b), ldg,
                                             n, h, w, c, _stream()), "norm_bwd_reduce_finalize")
    _count()
    return k


def norm_bwd(g_out, out_pad, upsample, y, coef, relu, mode, p_w, ldw, g_w, g_b, ldg, want_res, res_pad, eps=1e-5):
    """Returns (dy, g_res).  coef = (mean, rinv, a, b) from norm_finalize."""
    n, h, w, c = y.shape
    if FUSE_FINALIZE:
        k = norm_bwd_reduce_finalize(g_out, out_pad, upsample, y, coef, relu, mode, p_w, ldw, g_w, g_b, ldg, eps)
    else:
        sums = norm_bwd_reduce(g_out, out_pad, upsample, y, coef, relu)
        k = norm_bwd_finalize(sums, coef, mode, p_w, ldw, g_w, g_b, ldg, h * w, eps)
    return norm_bwd_apply(g_out, out_pad, upsample, y, coef, relu, k, want_res, res_pad)


def act_bwd_takes_bias(c):
    return 256 % (c // 8) == 0


def act_bwd(g_out, out_act, pad, act, dbias=None, c_out=None):
    """dbias (fp32, optional): += column sums of the result over the first c_out channels (the bias gradient)."""
    n, hp, wp, c = out_act.shape
    h, w = hp - 2 * pad, wp - 2 * pad
    dy = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=g_out.device)
    assert dbias is None or act_bwd_takes_bias(c)
    check(lib.munit_act_bwd(g_out.data_ptr(), out_act.data_ptr(), pad, ACT[act], dy.data_ptr(), n, h, w, c,
                            _ptr(dbias), c if c_out is None else c_out, _stream()), "act_bwd")
    _count()
    return dy


def colsum(dy, out, c_out=None):
    c = dy.shape[-1]
    check(lib.munit_colsum(dy.data_ptr(), out.data_ptr(), dy.numel() // c, c, c if c_out is None else c_out,
                           _stream()), "colsum")
    _count()
    return out


def rspace_combine(r, bias, cout, kw, act):
    n, h, wp, c = r.shape
    assert c == 32
    w = wp - kw + 1
    out = torch.empty(n, cout, h, w, dtype=torch.float32, device=r.device)
    check(lib.munit_rspace_combine(r.data_ptr(), _ptr(bias), out.data_ptr(), n, cout, h, w, kw, ACT[act], _stream()),
          "rspace_combine")
    _count()
    return out


def rspace_expand(g, out, kw, act, dbias):
    n, cout, h, w = out.shape
    dr = torch.empty(n, h, w + kw - 1, 32, dtype=torch.bfloat16, device=out.device)
    check(lib.munit_rspace_expand(g.data_ptr(), out.data_ptr(), dr.data_ptr(), _ptr(dbias), n, cout, h, w, kw, ACT[act],
                                  _stream()), "rspace_expand")
    _count()
    return dr


# ---------------------------------------------------------------- weights
def gather_cast(src, idx, dst):
    check(lib.munit_gather_cast(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), dst.numel(), _stream()), "gather_cast")
    _count()
    return dst


def gather_add(src, idx, dst):
    check(lib.munit_gather_add(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), dst.numel(), _stream()), "gather_add")
    _count()
    return dst


GATHER_BLOCK = 2048  # MUNIT_GATHER_BLOCK


def gather_seg_table(segs, device):
    """segs: list of (src fp32 tensor, idx int32 tensor or None, dst bf16 tensor) -> (device table, nseg, nblocks) for
    gather_cast_multi (munit_gather_seg: src, idx, dst, n, block0 as five 64-bit words)."""
    rows, b0 = [], 0
    for src, idx, dst in segs:
        n = dst.numel()
        assert idx is None or idx.numel() == n
        rows.append([src.data_ptr(), 0 if idx is None else idx.data_ptr(), dst.data_ptr(), n, b0])
        b0 += (n + GATHER_BLOCK - 1) // GATHER_BLOCK
    return torch.tensor(rows, dtype=torch.int64).to(device), len(rows), b0


def gather_cast_multi(table, nseg, nblocks):
    check(lib.munit_gather_cast_multi(table.data_ptr(), nseg, nblocks, _stream()), "gather_cast_multi")
    _count()


def cast_bf16(src, dst):
    check(lib.munit_cast_bf16(src.data_ptr(), dst.data_ptr(), dst.numel(), _stream()), "cast_bf16")
    _count()
    return dst


# ---------------------------------------------------------------- small fp32 ops
def linear_fwd(x, w, bias, relu):
    b, i = x.shape
    o = w.shape[0]
    y = torch.empty(b, o, dtype=torch.float32, device=x.device)
    check(lib.munit_linear_fwd(x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), b, i, o, int(relu), _stream()),
          "linear_fwd")
    _count()
    return y


def mlp3_fwd(x, w1, b1, w2, b2, w3, b3):
    """relu(W1 x + b1) -> relu(W2 . + b2) -> W3 . + b3 in one launch; returns (h1, h2, y)."""
    b, i = x.shape
    d, o = w1.shape[0], w3.shape[0]
    assert w1.shape == (d, i) and w2.shape == (d, d) and w3.shape == (o, d)
    h1 = torch.empty(b, d, dtype=torch.float32, device=x.device)
    h2 = torch.empty(b, d, dtype=torch.float32, device=x.device)
    y = torch.empty(b, o, dtype=torch.float32, device=x.device)
    check(lib.munit_mlp3_fwd(x.data_ptr(), w1.data_ptr(), _ptr(b1), w2.data_ptr(), _ptr(b2), w3.data_ptr(), _ptr(b3),
                             h1.data_ptr(), h2.data_ptr(), y.data_ptr(), b, i, d, o, _stream()), "mlp3_fwd")
    _count()
    return h1, h2, y


def linear_bwd(x, w, y, dy, relu, need_dx, dw, db):
    b, i = x.shape
    o = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    check(lib.munit_linear_bwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), dy.data_ptr(), int(relu), _ptr(dx), _ptr(dw),
                               _ptr(db), b, i, o, _stream()), "linear_bwd")
    _count(2)
    return dx


def gap_fwd(y):
    n, h, w, c = y.shape
    out = torch.empty(n, c, dtype=torch.float32, device=y.device)
    check(lib.munit_gap_fwd(y.data_ptr(), out.data_ptr(), n, h * w, c, _stream()), "gap_fwd")
    _count()
    return out


def gap_bwd(g, shape):
    n, h, w, c = shape
    dy = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=g.device)
    check(lib.munit_gap_bwd(g.data_ptr(), dy.data_ptr(), n, h * w, c, _stream()), "gap_bwd")
    _count()
    return dy


def dis_head_fwd(y, w, bias, target, loss, scale):
    c = y.shape[-1]
    npix = y.numel() // c
    out = torch.empty(npix, dtype=torch.float32, device=y.device)
    check(lib.munit_dis_head_fwd(y.data_ptr(), w.data_ptr(), bias.data_ptr(), float(target), out.data_ptr(), _ptr(loss),
                                 float(scale), npix, c, _stream()), "dis_head_fwd")
    _count()
    return out


def dis_head_bwd(y, w, out, target, gscale_dev, gscale, dw, db):
    c = y.shape[-1]
    npix = y.numel() // c
    dy = torch.empty_like(y)
    check(lib.munit_dis_head_bwd(y.data_ptr(), w.data_ptr(), out.data_ptr(), float(target), _ptr(gscale_dev),
                                 float(gscale), dy.data_ptr(), _ptr(dw), _ptr(db), npix, c, _stream()), "dis_head_bwd")
    _count()
    return dy


def avgpool_fwd(x):
    n, c, h, w = x.shape
    y = torch.empty(n, c, (h + 1) // 2, (w + 1) // 2, dtype=torch.float32, device=x.device)
    check(lib.munit_avgpool3s2_fwd(x.data_ptr(), y.data_ptr(), n * c, h, w, _stream()), "avgpool_fwd")
    _count()
    return y


def avgpool_bwd(gy, gx):
    n, c, h, w = gx.shape
    check(lib.munit_avgpool3s2_bwd(gy.data_ptr(), gx.data_ptr(), n * c, h, w, _stream()), "avgpool_bwd")
    _count()
    return gx


def l1_fwd(a, b, loss, scale):
    fn = lib.munit_l1_bf16_fwd if a.dtype == torch.bfloat16 else lib.munit_l1_fwd
    check(fn(a.data_ptr(), b.data_ptr(), loss.data_ptr(), float(scale), a.numel(), _stream()), "l1_fwd")
    _count()


def l1_bwd(a, b, gscale_dev, scale, ga, gb):
    fn = lib.munit_l1_bf16_bwd if a.dtype == torch.bfloat16 else lib.munit_l1_bwd
    check(fn(a.data_ptr(), b.data_ptr(), _ptr(gscale_dev), float(scale), _ptr(ga), _ptr(gb), a.numel(), _stream()),
          "l1_bwd")
    _count()


def l1_masked_fwd(a, b, keep, loss, scale):
    n, c, h, w = a.shape
    check(lib.munit_l1_masked_fwd(a.data_ptr(), b.data_ptr(), keep.data_ptr(), loss.data_ptr(), float(scale), n, c, h * w,
                                  _stream()), "l1_masked_fwd")
    _count()


def l1_masked_bwd(a, b, keep, gscale_dev, scale, ga, gb):
    n, c, h, w = a.shape
    check(lib.munit_l1_masked_bwd(a.data_ptr(), b.data_ptr(), keep.data_ptr(), _ptr(gscale_dev), float(scale), _ptr(ga),
                                  _ptr(gb), n, c, h * w, _stream()), "l1_masked_bwd")
    _count()


def adam(p, g, m, v, p_saved, p_bf16, mode, save, lr, b1, b2, eps, wd, step, gscale=1.0, hyper_dev=None):
    check(lib.munit_adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p_saved), _ptr(p_bf16), p.numel(),
                         mode, int(save), lr, b1, b2, eps, wd, step, gscale, _ptr(hyper_dev), _stream()), "adam")
    _count()


def fill(t, v):
    check(lib.munit_fill_f32(t.data_ptr(), float(v), t.numel(), _stream()), "fill")
    _count()
    return t


def add_bf16(dst, src):
    check(lib.munit_add_bf16(dst.data_ptr(), src.data_ptr(), dst.numel(), _stream()), "add_bf16")
    _count()
    return dst


# ---------------------------------------------------------------- domain-adaptation heads (csrc/heads.cu)
def maxpool2_fwd(x, in_pad):
    n, hp, wp, c = x.shape
    h, w = hp - 2 * in_pad, wp - 2 * in_pad
    y = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=x.device)
    check(lib.munit_maxpool2_fwd(x.data_ptr(), in_pad, y.data_ptr(), n, h, w, c, _stream()), "maxpool2_fwd")
    _count()
    return y


def maxpool2_bwd(gy, x, in_pad):
    n, hp, wp, c = x.shape
    dx = torch.empty_like(x)
    check(lib.munit_maxpool2_bwd(gy.data_ptr(), x.data_ptr(), in_pad, dx.data_ptr(), n, hp - 2 * in_pad, wp - 2 * in_pad,
                                 c, _stream()), "maxpool2_bwd")
    _count()
    return dx


def _gather_ranks(t):
    """[n, ...] of every data-parallel rank stacked along dim 0 (synchronised BatchNorm); returns (all, n0)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not SYNC_BN:
        return t, 0
    out = torch.empty((dist.get_world_size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out, dist.get_rank() * t.shape[0]


# BatchNorm statistics over the global batch when torch.distributed is initialised (MUNIT_SYNC_BN=0: per rank).
SYNC_BN = _os.environ.get("MUNIT_SYNC_BN", "1") != "0"


def bn_fwd(y, gamma, beta, running_mean, running_var, momentum, eps, training, relu):
    """nn.BatchNorm2d (+ReLU) on y [N,H,W,C] bf16 -> (out [N,H,W,C], coef[4][N][C] = mean, rinv, a, b)."""
    n, h, w, c = y.shape
    coef = torch.empty(4, n, c, dtype=torch.float32, device=y.device)
    if training:
        stats, shift = norm_stats(y)
        splits = stats.shape[1]
        stats, _ = _gather_ranks(stats)
        shift, _ = _gather_ranks(shift)
        n_total = stats.shape[0]
        sp, shp = stats.data_ptr(), shift.data_ptr()
    else:
        splits, n_total, sp, shp = 0, 0, 0, 0
    check(lib.munit_bn_finalize(sp, splits, shp, n_total, n, _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                _ptr(running_var), float(momentum), float(eps), int(training), coef[0].data_ptr(),
                                coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), h * w, c, _stream()),
          "bn_finalize")
    _count()
    return norm_apply(y, coef[2], coef[3], relu, None, 0, 0, 1), coef


def bn_bwd(g_out, y, coef, relu, gamma, g_gamma, g_beta, training):
    """dy of bn_fwd; g_gamma / g_beta (+=, may be None)."""
    n, h, w, c = y.shape
    mean, rinv, a, b = coef[0], coef[1], coef[2], coef[3]
    splits = lib.munit_norm_splits(h * w, c)
    sums = torch.empty(n, splits, c, 2, dtype=torch.float32, device=y.device)
    check(lib.munit_norm_bwd_reduce(g_out.data_ptr(), 0, 1, y.data_ptr(), _f16(y), a.data_ptr(), b.data_ptr(), int(relu),
                                    mean.data_ptr(), rinv.data_ptr(), sums.data_ptr(), n, h, w, c, _stream()),
          "norm_bwd_reduce")
    sums_all, n0 = _gather_ranks(sums) if training else (sums, 0)
    k = torch.empty(3, n, c, dtype=torch.float32, device=y.device)
    check(lib.munit_bn_bwd_finalize(sums_all.data_ptr(), splits, sums_all.shape[0], n0, n, _ptr(gamma), rinv.data_ptr(),
                                    int(training), k[0].data_ptr(), k[1].data_ptr(), k[2].data_ptr(), _ptr(g_gamma),
                                    _ptr(g_beta), h * w, c, _stream()), "bn_bwd_finalize")
    dy = torch.empty_like(y)
    check(lib.munit_norm_bwd_apply(g_out.data_ptr(), 0, 1, y.data_ptr(), _f16(y), a.data_ptr(), b.data_ptr(), int(relu),
                                   mean.data_ptr(), rinv.data_ptr(), k[0].data_ptr(), k[1].data_ptr(), k[2].data_ptr(),
                                   dy.data_ptr(), 0, 0, n, h, w, c, _stream()), "norm_bwd_apply")
    _count(3)
    return dy


def add_relu(a, b):
    out = torch.empty_like(a)
    check(lib.munit_add_relu(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "add_relu")
    _count()
    return out


def mse_const_fwd(x, target, scale):
    loss = torch.empty(1, dtype=torch.float32, device=x.device)
    check(lib.munit_mse_const_fwd(x.data_ptr(), float(target), loss.data_ptr(), float(scale), x.numel(), _stream()),
          "mse_const_fwd")
    _count()
    return loss


def mse_const_bwd(x, target, gscale_dev, scale):
    dx = torch.empty_like(x)
    check(lib.munit_mse_const_bwd(x.data_ptr(), float(target), gscale_dev.data_ptr(), float(scale), dx.data_ptr(),
                                  x.numel(), _stream()), "mse_const_bwd")
    _count()
    return dx
