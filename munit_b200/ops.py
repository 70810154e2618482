"""Autograd glue: torch.autograd.Function wrappers that sequence the sm_100a kernels (kernels.py).

Internal activation format ("act"): torch bf16 tensor [N, H+2P, W+2P, C], NHWC, with the reflect halo
of width P that the *consuming* convolution needs already materialised (nn.ReflectionPad2d,
networks.py:643, is never executed as an op).  `Act` carries the tensor and its P.

Parameter gradients: when a parameter already owns a `.grad` buffer (the trainer pre-allocates them in
one flat arena) the kernels accumulate straight into it and the Function returns None for that input;
otherwise a fresh gradient tensor is returned and autograd accumulates as usual.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

import os
import weakref

from . import geometry as G
from . import kernels as K

class _WgradOverlap:
    """Weight gradients are consumed only by the optimizer step, so they need not sit on the critical chain
    dgrad -> norm backward -> dgrad ...: when enabled (trainer / StepRunner with two_streams), each wgrad is
    issued on a companion stream of the stream that runs the backward chain and joined before the optimizer step.
    The operands are kept alive until the join (the caching allocator would otherwise hand their memory to later
    kernels of the main stream)."""
    enabled = False
    streams: dict = {}   # backward-chain stream id -> its companion stream
    active: list = []    # companion streams that received work since the last join
    keep: list = []


WG = _WgradOverlap()


def wgrad_async(fn, *keep):
    """Run fn() (wgrad launches that accumulate into a parameter-gradient arena) on the companion stream."""
    if not WG.enabled:
        fn()
        return
    cur = torch.cuda.current_stream()
    ws = WG.streams.get(cur.cuda_stream)
    if ws is None:
        ws = WG.streams[cur.cuda_stream] = torch.cuda.Stream()
    ws.wait_stream(cur)
    with torch.cuda.stream(ws):
        fn()
    if ws not in WG.active:
        WG.active.append(ws)
    WG.keep.append(keep)


def wgrad_join():
    """Make the current stream wait for every outstanding companion-stream wgrad."""
    if WG.active:
        cur = torch.cuda.current_stream()
        for ws in WG.active:
            cur.wait_stream(ws)
    WG.active.clear()
    WG.keep.clear()


# Extra compute streams that received work in the current update (e.g. the discriminator's per-scale streams):
# the trainer orders the optimizer step after them once the backward pass is over (parameter gradients are written
# by kernels, not AccumulateGrad nodes, so autograd does not do it).
SIDE_ACTIVE: list = []


def note_side_stream(st):
    if st not in SIDE_ACTIVE:
        SIDE_ACTIVE.append(st)


def join_side_streams():
    if SIDE_ACTIVE:
        cur = torch.cuda.current_stream()
        for st in SIDE_ACTIVE:
            cur.wait_stream(st)
    SIDE_ACTIVE.clear()


# Data-parallel gradient synchronisation (dp.GradSync instances, one per optimiser arena): Functions that own
# parameters report their uses (forward) and the completion of their gradient launches (backward), so that a
# bucket of the flat gradient arena can be all-reduced as soon as it is final -- overlapped with the rest of the
# backward pass.  Empty list = single GPU, zero overhead.
SYNCS: list = []


def _track_use(ctx, needs, *bufs):
    """Forward: remember which pre-allocated gradient buffers the backward of this node will complete."""
    ctx.tracked = ()
    if SYNCS and needs:
        ctx.tracked = tuple(b for b in bufs if b is not None)
        for b in ctx.tracked:
            for s in SYNCS:
                s.note_use(b)


def _track_done(ctx):
    for b in getattr(ctx, "tracked", ()):
        for s in SYNCS:
            s.note_done(b)


def _active_wgrad_streams():
    return list(WG.active)


# Normalisation statistics in the conv epilogue (no separate pass over the conv output); MUNIT_EPI_STATS=0 restores
# the stand-alone statistics kernel.
EPI_STATS = os.environ.get("MUNIT_EPI_STATS", "0") != "0"


@dataclass
class Act:
    t: torch.Tensor  # [N, H+2P, W+2P, C] bf16
    pad: int

    @property
    def n(self):
        return self.t.shape[0]

    @property
    def h(self):
        return self.t.shape[1] - 2 * self.pad

    @property
    def w(self):
        return self.t.shape[2] - 2 * self.pad

    @property
    def c(self):
        return self.t.shape[3]


def _cl_weight(weight: torch.Tensor) -> torch.Tensor:
    """Flat view of a conv weight in [Cout][KH][KW][Cin] order (channels_last memory)."""
    if weight.dim() == 4 and not weight.is_contiguous(memory_format=torch.channels_last):
        weight = weight.contiguous(memory_format=torch.channels_last)
    return weight.permute(0, 2, 3, 1).reshape(-1) if weight.dim() == 4 else weight.reshape(-1)


def _grad_like_cl(weight: torch.Tensor) -> torch.Tensor:
    return torch.zeros(weight.shape, dtype=torch.float32, device=weight.device).contiguous(
        memory_format=torch.channels_last)


# Layers whose bf16 shadows exist, keyed by the address of their fp32 master.  An optimiser refreshes every shadow of
# its arena with ONE launch right after the parameter update (refresh_shadows) instead of one or two launches per layer
# at the next use (reference: the fp32 nn.Conv2d weights are used directly, networks.py:688).
_SHADOWED = weakref.WeakValueDictionary()
_SHADOW_GEN = [0]


def refresh_shadows(opt, params):
    """Called by optim.FlatAdam after it changed `params` (and bumped their versions).  No-op for layers not built
    yet; inside a graph capture the segment table must already exist (it is built by the eager warm-up steps)."""
    cache = getattr(opt, "_shadow_cache", None)
    if cache is None or cache[0] != _SHADOW_GEN[0]:
        if torch.cuda.is_current_stream_capturing():
            return  # stale table and no way to upload one here: the lazy per-layer path serves this capture
        layers, segs = [], []
        for p in params:
            layer = _SHADOWED.get(p.data_ptr())
            seg = layer.shadow_segments(p) if layer is not None else None
            if seg:
                layers.append((layer, p))
                segs += seg
        table = K.gather_seg_table(segs, params[0].device) if segs else None
        cache = opt._shadow_cache = (_SHADOW_GEN[0], layers, table, segs)
    _, layers, table, _ = cache
    if table is None:
        return
    K.gather_cast_multi(*table)
    for layer, p in layers:
        layer._wver = (p._version, p.data_ptr())


class ConvLayer:
    """Per-convolution host state: geometry, bf16 weight shadows, cached plans.  Not an nn.Module --
    owned by Conv2dBlock next to the nn.Conv2d that holds the fp32 master parameters."""

    def __init__(self, cin, cout, k, stride, pad, zpad=0):
        """pad: reflect halo the input act carries; zpad: implicit zero padding (nn.Conv2d(padding=zpad)) of an
        input act without halo -- out-of-range taps are TMA zero fill (geometry.plan_fwd)."""
        self.cin, self.cout, self.k, self.stride, self.pad, self.zpad = cin, cout, k, stride, pad, zpad
        assert not (pad and zpad)
        self.first = cin < 64  # image-space layer: kw-expanded GEMM (K = 64 per kh tap)
        # narrow-output layer (64->3 7x7 tanh): vertical GEMM with N = (kw, co) + horizontal combine
        self.last = (not self.first) and cout <= 4 and k <= 8 and stride == 1
        if self.first:
            assert cin <= 4 and k in (4, 7), "first-layer path supports 7x7 (s1) and 4x4 (s2) RGB convs"
            self.kwp, self.cp = (8, 8) if k == 7 else (4, 16)
            self.c_buf = 64
        else:
            assert cin % 64 == 0, f"input channels {cin} must be a multiple of 64 (or an image)"
            self.c_buf = cin
        self.co_rows = 32 if self.last else ((cout + 15) // 16) * 16
        self.ck = max(64, self.co_rows)  # K extent per tap of the dgrad matrix
        self._wver = self._bver = None
        self.w_fwd = self.w_dg = self.bias_p = None
        self._idx_fwd = self._idx_dg = self._idx_inv = None
        self._plans = {}

    # ---- weight shadows -------------------------------------------------------------------
    def _build_maps(self, dev):
        c, k = self, self.k
        if c.last:
            fwd, dg, inv = G.rspace_index_maps(c.cout, c.cin, k, k)
            self._idx_fwd, self._idx_dg, self._idx_inv = fwd.to(dev), dg.to(dev), inv.to(dev)
            self._identity_fwd = False
            return
        if c.first:
            fwd = G.fwd_index_map(c.cout, c.cin, k, k, c.co_rows, 64, c.kwp, c.cp)
            dg = G.dgrad_index_map(c.cout, c.cin, k, k, c.stride, 1, 64, c.ck, c.kwp, c.cp)
        else:
            fwd = G.fwd_index_map(c.cout, c.cin, k, k, c.co_rows, c.cin)
            dg = G.dgrad_index_map(c.cout, c.cin, k, k, c.stride, c.stride, c.cin, c.ck)
        self._idx_fwd, self._idx_dg = fwd.to(dev), dg.to(dev)
        if c.first:  # param element -> position in the padded fwd matrix (for the wgrad scatter-back)
            inv = torch.full((c.cout * k * k * c.cin,), -1, dtype=torch.int32)
            m = fwd >= 0
            inv[fwd[m].long()] = torch.arange(fwd.numel(), dtype=torch.int32)[m]
            self._idx_inv = inv.to(dev)
        self._identity_fwd = (not c.first) and c.co_rows == c.cout

    def refresh(self, weight, bias, force=False):
        """(Re)build the bf16 GEMM operands from the fp32 master weights when they changed.  After an optimiser step
        refresh_shadows() has normally done it already (one launch for the whole arena) and this is a no-op."""
        wver = (weight._version, weight.data_ptr())
        if force or wver != self._wver:
            dev = weight.device
            rebuilt = self._idx_fwd is None or self._idx_fwd.device != dev
            if rebuilt:
                self._build_maps(dev)
                self.w_fwd = torch.empty(self.co_rows, self._idx_fwd.numel() // self.co_rows, dtype=torch.bfloat16,
                                         device=dev)
                rows = 64 if self.first else self.cin
                self.w_dg = torch.empty(rows, self._idx_dg.numel() // rows, dtype=torch.bfloat16, device=dev)
                self.bias_p = torch.zeros(self.co_rows, dtype=torch.float32, device=dev)
            src = _cl_weight(weight.detach())
            if self._identity_fwd:
                K.cast_bf16(src, self.w_fwd)
            else:
                K.gather_cast(src, self._idx_fwd, self.w_fwd)
            K.gather_cast(src, self._idx_dg, self.w_dg)
            self._wver = wver
            if rebuilt or _SHADOWED.get(wver[1]) is not self:
                _SHADOWED[wver[1]] = self
                _SHADOW_GEN[0] += 1
        if bias is not None and not self.last:
            if self.co_rows == self.cout:
                self.bias_p = bias.detach()  # a view of the master: always current
            else:
                bver = (bias._version, bias.data_ptr())
                if force or bver != self._bver:
                    self.bias_p[: self.cout].copy_(bias.detach())
                    self._bver = bver

    def shadow_segments(self, weight):
        """(src, idx, dst) triples of this layer's shadows for kernels.gather_seg_table, or None when the master is not
        stored channels_last in place (then the lazy path of refresh() keeps serving it)."""
        src = _cl_weight(weight.detach())
        if src.data_ptr() != weight.data_ptr() or self._idx_fwd is None or self._idx_fwd.device != weight.device:
            return None
        return [(src, None if self._identity_fwd else self._idx_fwd, self.w_fwd), (src, self._idx_dg, self.w_dg)]

    # ---- plans ----------------------------------------------------------------------------
    def plans(self, n, hp, wp, out_pad):
        """hp, wp: extents of the GEMM input buffer (padded act, or the kw-expanded image)."""
        key = (n, hp, wp, out_pad)
        p = self._plans.get(key)
        if p is None:
            k, s = self.k, self.stride
            if self.first or self.last:
                kh, kw, sy, sx, c = k, 1, s, 1, 64
            else:
                kh, kw, sy, sx, c = k, k, s, s, self.cin
            zp = self.zpad
            ho, wo = G.conv_out(hp + 2 * zp, kh, sy), G.conv_out(wp + 2 * zp, kw, sx)
            hop, wop = ho + 2 * out_pad, wo + 2 * out_pad
            fwd = G.plan_fwd(n, hp, wp, c, kh, kw, sy, sx, self.co_rows,
                             (hop * wop * self.co_rows, wop * self.co_rows, self.co_rows, out_pad, out_pad),
                             halo=_want_halo(kh, kw, ho, wo), zpad=zp)
            dg = G.plan_dgrad(n, hp, wp, c, kh, kw, sy, sx, self.co_rows, halo=_want_halo(kh, kw, hp, wp), zpad=zp)
            if self.first:
                wg = G.plan_wgrad(n, hp, wp, 64, kh, 1, sy, 1, self.co_rows, self.cout, kh * 64, 64, 1)
            elif self.last:
                wg = G.plan_wgrad(n, hp, wp, 64, kh, 1, 1, 1, 32, 32, kh * 64, 64, 1)
            else:
                wg = G.plan_wgrad(n, hp, wp, c, kh, kw, sy, sx, self.co_rows, self.cout, kh * kw * c, c, 1, zpad=zp)
            # direct-form algorithmic work of this layer instance: 1 MAC = 2 FLOP
            flops = 2.0 * n * ho * wo * self.cout * self.k * self.k * self.cin
            fwd.alg_flops = dg.alg_flops = wg.alg_flops = flops
            p = (fwd, dg, wg, ho, wo)
            self._plans[key] = p
        return p


def _want_halo(kh, kw, out_h, out_w) -> int:
    """Use the halo-resident tap-GEMM variant where it measured faster (profiles/r1_halo.md): many taps (5x5 and
    up) and an output extent that 8 x 16 tiles cover with <= 10 % waste.  (The UMMA descriptor base offset stays 0:
    the 128B swizzle is a function of the absolute shared-memory address.)"""
    if kh * kw < 25:
        return 0
    cover = (-(-out_w // 8) * 8) * (-(-out_h // 16) * 16)
    return 1 if out_w * out_h >= 0.9 * cover else 0


def _param_grad_buf(p: Optional[torch.Tensor]):
    """The pre-allocated .grad of a leaf parameter, if any (kernels accumulate into it directly)."""
    if p is None or not isinstance(p, torch.nn.Parameter):
        return None
    return p.grad


class ConvFn(torch.autograd.Function):
    """Conv2dBlock with norm='none' (networks.py:695-701): reflect-padded input act -> conv + bias +
    activation -> output act with the next layer's halo.  Also used with act='none', out_pad=0 to
    produce the raw conv output that feeds a norm."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer: ConvLayer, act: str, out_pad: int, image_pad: int, stats_kind: int = 0):
        """stats_kind 1 (IN / AdaIN) or 2 (LayerNorm): the output feeds a normalisation.  It is then stored as IEEE
        fp16 -- only the norm kernels ever read it, and the 3 extra mantissa bits keep the sign of x - mean (the ReLU
        mask of networks.py:698-700) stable against the storage rounding -- but *typed* bf16 so that autograd does not
        cast the bf16 gradient that comes back (NormFn views it as float16 again: y_f16=True).  Also returns the
        normalisation partials the epilogue computed from the stored tile (an empty tensor when this plan cannot
        produce them)."""
        layer.refresh(weight, bias)
        if layer.first:  # x is an NCHW fp32 image
            n, c, h, w = x.shape
            gemm_in = K.image_to_kwexp(x.detach().contiguous(), image_pad, layer.k, layer.stride, layer.kwp, layer.cp)
            ctx.img_shape = (n, c, h, w)
        else:
            gemm_in = x
        n, hp, wp, _ = gemm_in.shape
        fwd, _, _, ho, wo = layer.plans(n, hp, wp, out_pad)
        out = torch.empty(n, ho + 2 * out_pad, wo + 2 * out_pad, layer.co_rows,
                          dtype=torch.float16 if stats_kind else torch.bfloat16, device=x.device)
        part = None
        if stats_kind:
            splits = K.stats_splits(fwd, stats_kind) if (EPI_STATS and out_pad == 0 and act == "none") else 0
            part = torch.empty(n * splits * (layer.co_rows if stats_kind == 1 else 1) * 2, dtype=torch.float32,
                               device=x.device)
        K.tapgemm(fwd, gemm_in, layer.w_fwd, out, layer.bias_p if bias is not None else None, act,
                  stats=part if (part is not None and part.numel()) else None, stats_kind=stats_kind)
        K.halo_fill(out, out_pad)
        ctx.layer, ctx.act, ctx.out_pad, ctx.image_pad = layer, act, out_pad, image_pad
        ctx.has_bias = bias is not None
        ctx.wbuf, ctx.bbuf = _param_grad_buf(weight), _param_grad_buf(bias)
        _track_use(ctx, ctx.needs_input_grad[1], ctx.wbuf, ctx.bbuf)
        ctx.save_for_backward(gemm_in, out if not stats_kind else out.new_empty(0), weight)
        if stats_kind:
            ctx.mark_non_differentiable(part)
            return out.view(torch.bfloat16), part
        return out

    @staticmethod
    def backward(ctx, g_out, g_part=None):
        layer: ConvLayer = ctx.layer
        gemm_in, out, weight = ctx.saved_tensors
        n, hp, wp, _ = gemm_in.shape
        _, dg, wg, ho, wo = layer.plans(n, hp, wp, ctx.out_pad)
        g_out = g_out.contiguous()
        gx = gw = gb = tgt = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            buf = ctx.bbuf
            tgt = buf if buf is not None else torch.zeros(layer.cout, dtype=torch.float32, device=g_out.device)
            gb = None if buf is not None else tgt
        if ctx.act != "none" or ctx.out_pad > 0:
            fused = tgt is not None and K.act_bwd_takes_bias(out.shape[3])  # bias gradient in the same pass
            dy = K.act_bwd(g_out, out, ctx.out_pad, ctx.act, tgt if fused else None, layer.cout)
            if fused:
                tgt = None
        else:
            dy = g_out
        if tgt is not None:
            K.colsum(dy, tgt, layer.cout)
        if ctx.needs_input_grad[0]:
            rows = 64 if layer.first else layer.cin
            dxp = torch.empty(n, hp, wp, rows, dtype=torch.bfloat16, device=dy.device)
            K.tapgemm(dg, dy, layer.w_dg, dxp)
            if layer.first:
                ni, ci, hi, wi = ctx.img_shape
                gx = K.kwexp_to_image_grad(dxp, ni, ci, hi, wi, ctx.image_pad, layer.k, layer.stride, layer.kwp,
                                           layer.cp)
            else:
                gx = dxp
        if ctx.needs_input_grad[1]:
            buf = ctx.wbuf
            tgt = buf if buf is not None else _grad_like_cl(weight)

            def run():
                if layer.first:
                    tmp = torch.zeros(layer.cout * layer.k * 64, dtype=torch.float32, device=dy.device)
                    K.wgrad(wg, dy, gemm_in, tmp)
                    K.gather_add(tmp, layer._idx_inv, _cl_weight(tgt))
                else:
                    K.wgrad(wg, dy, gemm_in, _cl_weight(tgt))

            if buf is not None:
                wgrad_async(run, dy, gemm_in, tgt)
            else:
                run()
            gw = None if buf is not None else tgt
        _track_done(ctx)
        return gx, gw, gb, None, None, None, None, None


class ConvOutFn(torch.autograd.Function):
    """Narrow-output Conv2dBlock (64 -> 3, 7x7, tanh; networks.py:548-559): act with halo -> NCHW fp32 image.
    Vertical taps on the tensor cores (7 taps, N = (kw, co) = 32), horizontal taps + bias + activation in
    one bandwidth kernel that writes the public layout directly."""

    @staticmethod
    def forward(ctx, x, weight, bias, layer: ConvLayer, act: str):
        layer.refresh(weight, bias)
        n, hp, wp, _ = x.shape
        fwd, _, _, ho, wo = layer.plans(n, hp, wp, 0)
        r = torch.empty(n, ho, wp, 32, dtype=torch.bfloat16, device=x.device)
        K.tapgemm(fwd, x, layer.w_fwd, r, None, "none")
        out = K.rspace_combine(r, bias.detach() if bias is not None else None, layer.cout, layer.k, act)
        ctx.layer, ctx.act = layer, act
        ctx.wbuf, ctx.bbuf = _param_grad_buf(weight), _param_grad_buf(bias)
        ctx.has_bias = bias is not None
        _track_use(ctx, ctx.needs_input_grad[1], ctx.wbuf, ctx.bbuf)
        ctx.save_for_backward(x, out, weight)
        return out

    @staticmethod
    def backward(ctx, g):
        layer: ConvLayer = ctx.layer
        x, out, weight = ctx.saved_tensors
        n, hp, wp, _ = x.shape
        _, dg, wg, ho, wo = layer.plans(n, hp, wp, 0)
        gb = gw = gx = None
        db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ctx.bbuf if ctx.bbuf is not None else torch.zeros(layer.cout, dtype=torch.float32, device=x.device)
            gb = None if ctx.bbuf is not None else db
        dr = K.rspace_expand(g.contiguous(), out, layer.k, ctx.act, db)
        if ctx.needs_input_grad[0]:
            gx = torch.empty(n, hp, wp, layer.cin, dtype=torch.bfloat16, device=x.device)
            K.tapgemm(dg, dr, layer.w_dg, gx)
        if ctx.needs_input_grad[1]:
            tgt = ctx.wbuf if ctx.wbuf is not None else _grad_like_cl(weight)

            def run():
                tmp = torch.zeros(32 * layer.k * 64, dtype=torch.float32, device=x.device)
                K.wgrad(wg, dr, x, tmp)
                K.gather_add(tmp, layer._idx_inv, _cl_weight(tgt))

            if ctx.wbuf is not None:
                wgrad_async(run, dr, x, tgt)
            else:
                run()
            gw = None if ctx.wbuf is not None else tgt
        _track_done(ctx)
        return gx, gw, gb, None, None


class NormFn(torch.autograd.Function):
    """InstanceNorm2d / AdaptiveInstanceNorm2d / MUNIT LayerNorm (networks.py:657,810-878) + ReLU +
    residual add (networks.py:623) + nearest-2x upsample (networks.py:534) + the next conv's reflect
    halo, as statistics -> finalize -> one apply pass."""

    @staticmethod
    def forward(ctx, y, p_w, p_b, residual, mode: str, relu: bool, res_pad: int, out_pad: int, upsample: int,
                eps: float, part=None, y_f16: bool = False):
        """y_f16: y holds IEEE fp16 bits behind its bf16 dtype (raw conv output from ConvFn(stats_kind != 0))."""
        n, h, w, c = y.shape
        y = y.contiguous()
        if y_f16:
            y = y.view(torch.float16)
        ldw = 0
        if mode == "adain":
            assert p_w.stride(1) == 1 and p_b.stride(1) == 1 and p_w.stride(0) == p_b.stride(0)
            ldw = p_w.stride(0)
        out, coef = K.norm_fwd(y, mode, p_w, p_b, ldw, eps, relu, residual, res_pad, out_pad, upsample, part)
        ctx.cfg = (mode, relu, res_pad, out_pad, upsample, eps, ldw, residual is not None)
        ctx.wbuf, ctx.bbuf = _param_grad_buf(p_w), _param_grad_buf(p_b)
        _track_use(ctx, mode == "ln" and ctx.needs_input_grad[1], ctx.wbuf, ctx.bbuf)
        ctx.save_for_backward(y, coef, p_w)
        return out

    @staticmethod
    def backward(ctx, g_out):
        mode, relu, res_pad, out_pad, upsample, eps, ldw, has_res = ctx.cfg
        y, coef, p_w = ctx.saved_tensors
        n, h, w, c = y.shape
        g_w = g_b = None
        ret_w = ret_b = None
        ldg = 0
        if mode == "adain":
            g_w = torch.empty(n, c, dtype=torch.float32, device=y.device)
            g_b = torch.empty(n, c, dtype=torch.float32, device=y.device)
            ldg, ret_w, ret_b = c, g_w, g_b
        elif mode == "ln":
            g_w = ctx.wbuf if ctx.wbuf is not None else torch.zeros(c, dtype=torch.float32, device=y.device)
            g_b = ctx.bbuf if ctx.bbuf is not None else torch.zeros(c, dtype=torch.float32, device=y.device)
            ret_w = None if ctx.wbuf is not None else g_w
            ret_b = None if ctx.bbuf is not None else g_b
        dy, g_res = K.norm_bwd(g_out.contiguous(), out_pad, upsample, y, coef, relu, mode, p_w, ldw, g_w, g_b, ldg,
                               has_res and ctx.needs_input_grad[3], res_pad, eps)
        _track_done(ctx)
        return dy, ret_w, ret_b, g_res, None, None, None, None, None, None, None, None


class AdainSplitFn(torch.autograd.Function):
    """assign_adain_params (networks.py:230-239) as ONE autograd node: per AdaIN layer the columns [:C] (bias) and
    [C:2C] (weight) of the MLP output, returned as strided views (the norm kernels index them in place).  The
    backward concatenates the 2L slice gradients with one kernel; the reference's chain of nested slices costs a
    zero-fill, a copy and an add per slice (~190 launches per generator update)."""

    @staticmethod
    def forward(ctx, params, sizes):
        outs, off = [], 0
        for c in sizes:
            outs.append(params[:, off:off + c])
            outs.append(params[:, off + c:off + 2 * c])
            off += 2 * c
        ctx.sizes, ctx.shape, ctx.used = tuple(sizes), tuple(params.shape), off
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        n, total = ctx.shape
        ref = next((g for g in grads if g is not None), None)
        if ref is None:
            return None, None
        parts = []
        for i, g in enumerate(grads):
            c = ctx.sizes[i // 2]
            parts.append(g if g is not None else torch.zeros(n, c, dtype=ref.dtype, device=ref.device))
        if ctx.used < total:
            parts.append(torch.zeros(n, total - ctx.used, dtype=ref.dtype, device=ref.device))
        return torch.cat(parts, 1), None


class ToActFn(torch.autograd.Function):
    """NCHW fp32 (public tensor format) -> act with reflect halo."""

    @staticmethod
    def forward(ctx, x, pad: int):
        ctx.pad = pad
        ctx.c = x.shape[1]
        return K.nchw_to_act(x.contiguous(), pad)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        folded = K.act_bwd(g, g, ctx.pad, "none") if ctx.pad else g
        return K.act_to_nchw(folded, ctx.c, 0), None


class FromActFn(torch.autograd.Function):
    """act interior (first `c` channels) -> NCHW fp32."""

    @staticmethod
    def forward(ctx, t, c: int, pad: int):
        ctx.pad, ctx.cp = pad, t.shape[3]
        return K.act_to_nchw(t, c, pad)

    @staticmethod
    def backward(ctx, g):
        # halo of the gradient buffer is zero (reads of the halo do not exist in the forward)
        gt = K.nchw_to_act(g.contiguous(), ctx.pad, ctx.cp, fill_halo=False)
        if ctx.pad:
            p = ctx.pad
            gt[:, :p].zero_(); gt[:, -p:].zero_(); gt[:, :, :p].zero_(); gt[:, :, -p:].zero_()
        return gt, None, None


class RepadFn(torch.autograd.Function):
    """Change the halo width of an act (rare: only when a caller hands an act to a layer that needs a
    different pad than its producer assumed)."""

    @staticmethod
    def forward(ctx, t, pad_in: int, pad_out: int):
        n, hp, wp, c = t.shape
        h, w = hp - 2 * pad_in, wp - 2 * pad_in
        out = torch.empty(n, h + 2 * pad_out, w + 2 * pad_out, c, dtype=t.dtype, device=t.device)
        out[:, pad_out:pad_out + h, pad_out:pad_out + w] = t[:, pad_in:pad_in + h, pad_in:pad_in + w]
        K.halo_fill(out, pad_out)
        ctx.pads = (pad_in, pad_out, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        pad_in, pad_out, h, w = ctx.pads
        g = g.contiguous()
        folded = K.act_bwd(g, g, pad_out, "none") if pad_out else g
        out = torch.zeros(g.shape[0], h + 2 * pad_in, w + 2 * pad_in, g.shape[3], dtype=g.dtype, device=g.device)
        out[:, pad_in:pad_in + h, pad_in:pad_in + w] = folded
        return out, None, None


class LinearFn(torch.autograd.Function):
    """nn.Linear (+ReLU) of the style MLP (networks.py:583-597,704-749) and the 1x1 style head
    (networks.py:472), fp32.  `weight` is [out, in] or a contiguous [out, in, 1, 1] conv weight."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu: bool):
        x = x.contiguous()
        assert weight.is_contiguous()
        y = K.linear_fwd(x, weight, bias, relu)
        ctx.relu = relu
        ctx.wbuf, ctx.bbuf = _param_grad_buf(weight), _param_grad_buf(bias)
        _track_use(ctx, ctx.needs_input_grad[1], ctx.wbuf, ctx.bbuf)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        need_w = ctx.needs_input_grad[1]
        dw = db = None
        if need_w:
            dw = ctx.wbuf if ctx.wbuf is not None else torch.zeros_like(weight)
            db = ctx.bbuf if ctx.bbuf is not None else torch.zeros(weight.shape[0], dtype=torch.float32,
                                                                      device=x.device)
        dx = K.linear_bwd(x, weight, y, gy.contiguous(), ctx.relu, ctx.needs_input_grad[0], dw, db)
        _track_done(ctx)
        return (dx, None if (not need_w or ctx.wbuf is not None) else dw,
                None if (not need_w or ctx.bbuf is not None) else db, None)


class Mlp3Fn(torch.autograd.Function):
    """The style MLP of AdaINGen (networks.py:583-597, n_blk = 3: Linear+ReLU, Linear+ReLU, Linear) that emits the AdaIN
    parameters: the forward pass is ONE kernel (munit_mlp3_fwd); the backward pass runs the per-layer linear
    backward kernels on the hidden activations the forward kernel kept."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        x = x.contiguous()
        h1, h2, y = K.mlp3_fwd(x, w1, b1, w2, b2, w3, b3)
        params = (w1, b1, w2, b2, w3, b3)
        ctx.bufs = tuple(_param_grad_buf(p) for p in params)
        ctx.need_w = any(ctx.needs_input_grad[1:])
        _track_use(ctx, ctx.need_w, *ctx.bufs)
        ctx.save_for_backward(x, h1, h2, y, w1, w2, w3)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, h1, h2, y, w1, w2, w3 = ctx.saved_tensors
        outs = [None] * 6

        def grads(i, w):
            if not ctx.need_w:
                return None, None
            dw = ctx.bufs[2 * i] if ctx.bufs[2 * i] is not None else torch.zeros_like(w)
            db = ctx.bufs[2 * i + 1] if ctx.bufs[2 * i + 1] is not None else torch.zeros(
                w.shape[0], dtype=torch.float32, device=w.device)
            outs[2 * i] = None if ctx.bufs[2 * i] is not None else dw
            outs[2 * i + 1] = None if ctx.bufs[2 * i + 1] is not None else db
            return dw, db

        dw, db = grads(2, w3)
        g2 = K.linear_bwd(h2, w3, y, gy.contiguous(), False, True, dw, db)
        dw, db = grads(1, w2)
        g1 = K.linear_bwd(h1, w2, h2, g2, True, True, dw, db)
        dw, db = grads(0, w1)
        gx = K.linear_bwd(x, w1, h1, g1, True, ctx.needs_input_grad[0], dw, db)
        _track_done(ctx)
        return (gx, *outs)


class GapFn(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(1) (networks.py:471) on an act with no halo -> fp32 [N, C]."""

    @staticmethod
    def forward(ctx, t):
        ctx.shape = tuple(t.shape)
        return K.gap_fwd(t.contiguous())

    @staticmethod
    def backward(ctx, g):
        return K.gap_bwd(g.contiguous(), ctx.shape)


class DisHeadFn(torch.autograd.Function):
    """MsImageDis head: 1x1 conv C->1 (networks.py:68) fused with its LSGAN term
    scale * mean((out - target)^2) (networks.py:91,109).  Returns (map fp32 [N,1,H,W], loss [1])."""

    @staticmethod
    def forward(ctx, t, weight, bias, target: float, scale: float):
        n, h, w, c = t.shape
        loss = torch.zeros(1, dtype=torch.float32, device=t.device)
        wv = weight.reshape(-1).contiguous()
        out = K.dis_head_fwd(t.contiguous(), wv, bias, target, loss, scale)
        ctx.cfg = (target, scale)
        ctx.wbuf, ctx.bbuf = _param_grad_buf(weight), _param_grad_buf(bias)
        _track_use(ctx, ctx.needs_input_grad[1], ctx.wbuf, ctx.bbuf)
        ctx.save_for_backward(t, weight, out)
        omap = out.view(n, 1, h, w)
        ctx.mark_non_differentiable(omap)
        return omap, loss

    @staticmethod
    def backward(ctx, g_map, g_loss):
        t, weight, out = ctx.saved_tensors
        target, scale = ctx.cfg
        need_w = ctx.needs_input_grad[1]
        dw = db = None
        if need_w:
            dw = ctx.wbuf if ctx.wbuf is not None else torch.zeros_like(weight)
            db = ctx.bbuf if ctx.bbuf is not None else torch.zeros(1, dtype=torch.float32, device=t.device)
        dy = K.dis_head_bwd(t, weight.reshape(-1).contiguous(), out, target, g_loss.contiguous(), scale,
                            dw.view(-1) if dw is not None else None, db)
        _track_done(ctx)
        return (dy if ctx.needs_input_grad[0] else None,
                None if (not need_w or ctx.wbuf is not None) else dw,
                None if (not need_w or ctx.bbuf is not None) else db, None, None)


class AvgPoolFn(torch.autograd.Function):
    """nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) (networks.py:32-34) on NCHW fp32."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return K.avgpool_fwd(x.contiguous())

    @staticmethod
    def backward(ctx, gy):
        gx = torch.zeros(ctx.shape, dtype=torch.float32, device=gy.device)
        return K.avgpool_bwd(gy.contiguous(), gx)


class L1MaskedFn(torch.autograd.Function):
    """recon_criterion_mask (trainer.py:292-305): mean over all elements of |(a - b) * keep|, keep = 1 - mask
    given per pixel ([N,1,H,W] or [N,H,W]) and broadcast over channels."""

    @staticmethod
    def forward(ctx, a, b, keep):
        a, b = a.contiguous(), b.contiguous()
        n, c, h, w = a.shape
        keep = keep.to(torch.float32).reshape(n, h * w).contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=a.device)
        K.l1_masked_fwd(a, b, keep, loss, 1.0 / a.numel())
        ctx.save_for_backward(a, b, keep)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b, keep = ctx.saved_tensors
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        K.l1_masked_bwd(a, b, keep, g.contiguous(), 1.0 / a.numel(), ga, gb)
        return ga, gb, None


class L1Fn(torch.autograd.Function):
    """recon_criterion (trainer.py:279-290): mean |a - b| over interiors.  a, b: fp32 tensors of the same
    shape, or acts with the same interior (halo widths pa, pb)."""

    @staticmethod
    def forward(ctx, a, b, pa: int, pb: int):
        if a.dtype == torch.bfloat16:
            n, hp, wp, c = a.shape
            h, w = hp - 2 * pa, wp - 2 * pa
            ai = a[:, pa:pa + h, pa:pa + w].contiguous() if pa else a.contiguous()
            bi = b[:, pb:pb + h, pb:pb + w].contiguous() if pb else b.contiguous()
        else:
            ai, bi = a.contiguous(), b.contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=a.device)
        K.l1_fwd(ai, bi, loss, 1.0 / ai.numel())
        ctx.pads = (pa, pb)
        ctx.save_for_backward(ai, bi)
        ctx.shapes = (a.shape, b.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        ai, bi = ctx.saved_tensors
        pa, pb = ctx.pads
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ga = torch.empty_like(ai) if need_a else None
        gb = torch.empty_like(bi) if need_b else None
        K.l1_bwd(ai, bi, g.contiguous(), 1.0 / ai.numel(), ga, gb)

        def embed(gi, shape, p):
            if gi is None or p == 0:
                return gi
            full = torch.zeros(shape, dtype=gi.dtype, device=gi.device)
            full[:, p:-p, p:-p] = gi
            return full

        return embed(ga, ctx.shapes[0], pa), embed(gb, ctx.shapes[1], pb), None, None


# ------------------------------------------------------------------ domain-adaptation heads (utils.py:1276-1392)
class MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2) (utils.py:1374,1376): interior of an act with halo `in_pad` -> act without halo.  The
    backward returns the gradient of the whole padded buffer (halo = 0), as the producer's backward expects."""

    @staticmethod
    def forward(ctx, x, in_pad: int):
        x = x.contiguous()
        ctx.in_pad = in_pad
        ctx.save_for_backward(x)
        return K.maxpool2_fwd(x, in_pad)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return K.maxpool2_bwd(gy.contiguous(), x, ctx.in_pad), None


class BatchNormFn(torch.autograd.Function):
    """nn.BatchNorm2d (+ReLU) of BasicBlock (utils.py:1300-1309,1316-1323) on a raw conv output [N,H,W,C]:
    per-channel batch statistics (training; running statistics updated in place) or running statistics (eval),
    then the generic normalise / backward-apply kernels."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, training: bool, relu: bool, momentum: float,
                eps: float):
        y = y.contiguous()
        out, coef = K.bn_fwd(y, gamma, beta, running_mean, running_var, momentum, eps, training, relu)
        ctx.cfg = (training, relu)
        ctx.wbuf, ctx.bbuf = _param_grad_buf(gamma), _param_grad_buf(beta)
        ctx.save_for_backward(y, coef, gamma)
        return out

    @staticmethod
    def backward(ctx, g_out):
        training, relu = ctx.cfg
        y, coef, gamma = ctx.saved_tensors
        c = y.shape[3]
        g_w = g_b = ret_w = ret_b = None
        if ctx.needs_input_grad[1]:
            g_w = ctx.wbuf if ctx.wbuf is not None else torch.zeros(c, dtype=torch.float32, device=y.device)
            ret_w = None if ctx.wbuf is not None else g_w
        if ctx.needs_input_grad[2]:
            g_b = ctx.bbuf if ctx.bbuf is not None else torch.zeros(c, dtype=torch.float32, device=y.device)
            ret_b = None if ctx.bbuf is not None else g_b
        dy = K.bn_bwd(g_out.contiguous(), y, coef, relu, gamma, g_w, g_b, training)
        return dy, ret_w, ret_b, None, None, None, None, None, None


class AddReluFn(torch.autograd.Function):
    """out = relu(a + b) (BasicBlock.forward, utils.py:1327-1329); both inputs receive g * (out > 0)."""

    @staticmethod
    def forward(ctx, a, b):
        out = K.add_relu(a.contiguous(), b.contiguous())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        d = K.act_bwd(g.contiguous(), out, 0, "relu")
        return d, d


class MseConstFn(torch.autograd.Function):
    """mean((x - target)^2) over all elements of an fp32 tensor (trainer.py:658-667)."""

    @staticmethod
    def forward(ctx, x, target: float):
        x = x.contiguous()
        ctx.target = float(target)
        ctx.save_for_backward(x)
        return K.mse_const_fwd(x, target, 1.0 / x.numel())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return K.mse_const_bwd(x, ctx.target, g.contiguous(), 1.0 / x.numel()), None
