"""Host-side planning: turns one convolution (forward, input-gradient, weight-gradient) into the
tap-GEMM / wgrad descriptors of include/munit_b200.h.  Pure Python, no device access -- the tap tables
built here are verified on the CPU against F.conv2d by tests/test_geometry.py.

Buffers (all NHWC bf16, contiguous):
  x   : [N, Hp, Wp, C]     conv input with its reflect halo already materialised (Hp = H + 2*pad)
  y   : [N, Ho, Wo, Co]    conv output, Ho = (Hp - KH)//SY + 1
Weight matrices (bf16, K contiguous):
  fwd   : [Co_rows][(kh*KW + kw)*C + c]                     (= the channels_last parameter memory)
  dgrad : [C_rows][phase][(a*NB + b)*Ck + co]               Ck = max(64, Co_pad), phase = py*SX + px
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple


def _pow2_floor(v: int) -> int:
    p = 1
    while p * 2 <= v:
        p *= 2
    return p


def _ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


def pick_tile(out_w: int, out_h: int, n_img: int, total: int) -> Tuple[int, int, int]:
    """(tw, th, tn) powers of two with tw*th*tn == total that waste the fewest lanes on ragged edges."""
    best, best_cost = None, None
    tw = 1
    while tw <= total:
        th = 1
        while tw * th <= total:
            tn = total // (tw * th)
            covered = (_ceil_div(out_w, tw) * tw) * (_ceil_div(out_h, th) * th) * (_ceil_div(n_img, tn) * tn)
            # prefer wide tiles on ties (longer contiguous TMA rows)
            cost = (covered, -tw, tn)
            if best_cost is None or cost < best_cost:
                best, best_cost = (tw, th, tn), cost
            th *= 2
        tw *= 2
    return best


def pick_bn(rows: int) -> int:
    for bn in (256, 128, 64, 32, 16):
        if rows % bn == 0:
            return bn
    raise ValueError(f"weight rows {rows} not a multiple of 16")


@dataclass
class TapGemmPlan:
    a_rank: int
    a_dim: List[int]
    a_stride: List[int]  # bytes
    a_box: List[int]
    b_rows: int
    b_k: int
    bn: int
    tw: int
    th: int
    tn: int
    out_w: int
    out_h: int
    n_img: int
    mx: List[int]
    my: List[int]
    mn: List[int]
    num_taps: int
    chunks: int
    tap_off: List[List[int]]
    phases: int
    b_k0: List[int]
    o_yoff: List[int]
    o_xoff: List[int]
    o_sn: int
    o_sy: int
    o_sx: int
    o_ymul: int
    o_xmul: int
    n_store: int
    halo: int = 0  # 1/2: halo-resident kernel variant (stride-1 only, 8x16 tiles)

    @property
    def grid(self):
        return (_ceil_div(self.out_w, self.tw) * _ceil_div(self.out_h, self.th) * _ceil_div(self.n_img, self.tn),
                self.b_rows // self.bn, self.phases)

    def flops(self) -> float:
        """MACs*2 actually issued to the tensor pipe (incl. ragged-tile waste)."""
        gx, gy, gz = self.grid
        return 2.0 * gx * 128 * gy * self.bn * gz * self.num_taps * self.chunks * 64


@dataclass
class WgradPlan:
    a_rank: int
    a_dim: List[int]
    a_stride: List[int]
    a_box: List[int]
    a_mx: List[int]
    a_my: List[int]
    a_mn: List[int]
    b_rank: int
    b_dim: List[int]
    b_stride: List[int]
    b_box: List[int]
    b_mx: List[int]
    b_my: List[int]
    b_mn: List[int]
    pw: int
    ph: int
    pn: int
    out_w: int
    out_h: int
    n_img: int
    m_total: int
    n_total: int
    bn: int
    num_taps: int
    tap_off: List[List[int]]
    s_m: int
    s_t: int
    s_n: int
    tap_on_a: int = 0


def _pad5(v, fill=0):
    return list(v) + [fill] * (5 - len(v))


def _x_view(n, hp, wp, c, sy, sx):
    """TMA view of the conv input for strides (sy, sx) in {1,2}: returns rank, dims, strides(bytes),
    (mx, my, mn) multipliers, a function tap->(offsets) and the box layout builder."""
    e = 2  # bytes per bf16
    if sy == 1 and sx == 1:
        dims = [c, wp, hp, n]
        strides = [e, c * e, wp * c * e, hp * wp * c * e]
        mx, my, mn = [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]
        off = lambda kh, kw: [0, kw, kh, 0]
        box = lambda tw, th, tn: [64, tw, th, tn]
        return 4, dims, strides, mx, my, mn, off, box
    if sy == 2 and sx == 2:
        assert hp % 2 == 0 and wp % 2 == 0, "stride-2 view needs even padded extents"
        dims = [2 * c, wp // 2, 2, hp // 2, n]
        strides = [e, 2 * c * e, wp * c * e, 2 * wp * c * e, hp * wp * c * e]
        mx, my, mn = [0, 1, 0, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1]
        off = lambda kh, kw: [(kw % 2) * c, kw // 2, kh % 2, kh // 2, 0]
        box = lambda tw, th, tn: [64, tw, 1, th, tn]
        return 5, dims, strides, mx, my, mn, off, box
    if sy == 2 and sx == 1:
        assert hp % 2 == 0
        dims = [c, wp, 2, hp // 2, n]
        strides = [e, c * e, wp * c * e, 2 * wp * c * e, hp * wp * c * e]
        mx, my, mn = [0, 1, 0, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1]
        off = lambda kh, kw: [0, kw, kh % 2, kh // 2, 0]
        box = lambda tw, th, tn: [64, tw, 1, th, tn]
        return 5, dims, strides, mx, my, mn, off, box
    raise ValueError(f"unsupported stride ({sy},{sx})")


def _shift_tap(t, zpad):
    """Tap offset of a stride-1 view [c, x, y, n] moved by the implicit zero padding."""
    return t if not zpad else [t[0], t[1] - zpad, t[2] - zpad] + list(t[3:])


def conv_out(hp, k, s):
    return (hp - k) // s + 1


def halo_ok(kh, kw, sy, sx, out_h, out_w, bn) -> bool:
    """Geometry the halo-resident kernel variant accepts."""
    return sy == 1 and sx == 1 and kw <= 9 and kh * kw > 1 and bn >= 64 and out_h >= 8 and out_w >= 8


def plan_fwd(n, hp, wp, c, kh, kw, sy, sx, co_rows, out_geom, n_store=None, halo=0, zpad=0) -> TapGemmPlan:
    """Forward conv: x [n,hp,wp,c] -> out.  out_geom = (o_sn, o_sy, o_sx, y_off, x_off) in elements:
    pixel (b, y, x) is stored at out + b*o_sn + (y+y_off)*o_sy + (x+x_off)*o_sx.
    zpad > 0: implicit zero padding of that width (nn.Conv2d(padding=zpad), utils.py:1238-1261) -- the buffer
    stays unpadded, the taps start at -zpad and TMA zero-fills what falls outside (stride 1 only)."""
    assert c % 64 == 0, "input channels must be a multiple of 64"
    assert zpad == 0 or (sy == 1 and sx == 1), "implicit zero padding is built for stride-1 convolutions"
    ho, wo = conv_out(hp + 2 * zpad, kh, sy), conv_out(wp + 2 * zpad, kw, sx)
    rank, dims, strides, mx, my, mn, off, box = _x_view(n, hp, wp, c, sy, sx)
    tw, th, tn = pick_tile(wo, ho, n, 128)
    taps = [_shift_tap(off(i, j), zpad) for i in range(kh) for j in range(kw)]
    halo = 0 if zpad else halo
    o_sn, o_sy, o_sx, y_off, x_off = out_geom
    bn = pick_bn(co_rows)
    halo = halo if halo_ok(kh, kw, sy, sx, ho, wo, bn) else 0
    if halo:
        tw, th, tn = 8, 16, 1
    return TapGemmPlan(
        halo=halo,
        a_rank=rank, a_dim=dims, a_stride=strides, a_box=box(tw, th, tn), b_rows=co_rows, b_k=kh * kw * c, bn=bn,
        tw=tw, th=th, tn=tn, out_w=wo, out_h=ho, n_img=n, mx=mx, my=my, mn=mn, num_taps=len(taps), chunks=c // 64,
        tap_off=taps, phases=1, b_k0=[0], o_yoff=[y_off], o_xoff=[x_off], o_sn=o_sn, o_sy=o_sy, o_sx=o_sx, o_ymul=1,
        o_xmul=1, n_store=co_rows if n_store is None else n_store)


def plan_dgrad(n, hp, wp, c_rows, kh, kw, sy, sx, co_c, halo=0, zpad=0) -> TapGemmPlan:
    """Input gradient: dy [n,ho,wo,co_c] (unpadded; out-of-range taps read zero through TMA) ->
    dxp [n,hp,wp,c_rows] (gradient w.r.t. the *padded* conv input, every position written).
    zpad > 0 (implicit zero padding, see plan_fwd): [hp, wp] is the unpadded input and the taps shift by +zpad."""
    assert zpad == 0 or (sy == 1 and sx == 1)
    ho, wo = conv_out(hp + 2 * zpad, kh, sy), conv_out(wp + 2 * zpad, kw, sx)
    halo = 0 if zpad else halo
    ck = max(64, co_c)
    assert ck % 64 == 0
    e = 2
    dims = [co_c, wo, ho, n]
    strides = [e, co_c * e, wo * co_c * e, ho * wo * co_c * e]
    na, nb = kh // sy, kw // sx
    assert na * sy == kh and nb * sx == kw, "kernel must be a multiple of the stride"
    assert hp % sy == 0 and wp % sx == 0
    oh, ow = hp // sy, wp // sx
    tw, th, tn = pick_tile(ow, oh, n, 128)
    taps = [[0, zpad - b, zpad - a, 0] for a in range(na) for b in range(nb)]
    phases = sy * sx
    kper = na * nb * ck
    halo = halo if halo_ok(kh, kw, sy, sx, oh, ow, pick_bn(c_rows)) else 0
    if halo:
        tw, th, tn = 8, 16, 1
    return TapGemmPlan(
        halo=halo,
        a_rank=4, a_dim=dims, a_stride=strides, a_box=[64, tw, th, tn], b_rows=c_rows, b_k=phases * kper,
        bn=pick_bn(c_rows), tw=tw, th=th, tn=tn, out_w=ow, out_h=oh, n_img=n, mx=[0, 1, 0, 0], my=[0, 0, 1, 0],
        mn=[0, 0, 0, 1], num_taps=len(taps), chunks=ck // 64, tap_off=taps, phases=phases,
        b_k0=[p * kper for p in range(phases)], o_yoff=[p // sx for p in range(phases)],
        o_xoff=[p % sx for p in range(phases)], o_sn=hp * wp * c_rows, o_sy=wp * c_rows, o_sx=c_rows, o_ymul=sy,
        o_xmul=sx, n_store=c_rows)


def plan_wgrad(n, hp, wp, c, kh, kw, sy, sx, co_c, m_total, s_m, s_t, s_n, n_total=None, swap=None,
               zpad=0) -> WgradPlan:
    """Weight gradient: dw[m, tap, cin] += sum_pix dy[pix, m] * x[pix@tap, cin].
    zpad > 0: x is unpadded and implicitly zero-padded (see plan_fwd)."""
    assert zpad == 0 or (sy == 1 and sx == 1)
    ho, wo = conv_out(hp + 2 * zpad, kh, sy), conv_out(wp + 2 * zpad, kw, sx)
    e = 2
    rank, dims, strides, mx, my, mn, off, box = _x_view(n, hp, wp, c, sy, sx)
    pw, ph, pn = pick_tile(wo, ho, n, 64)
    taps = [_shift_tap(off(i, j), zpad) for i in range(kh) for j in range(kw)]
    n_total = c if n_total is None else n_total
    if swap is None:  # put the wider channel count on the 128-row M operand
        swap = m_total <= 64 and n_total >= 128
    if swap:
        # A = x (M = input channels, shifted per tap), B = dy (N = output channels): dw[ci*s_n + tap*s_t + co*s_m]
        bn = 128 if m_total > 64 else 64
        return WgradPlan(
            a_rank=rank, a_dim=dims, a_stride=strides, a_box=box(pw, ph, pn), a_mx=mx, a_my=my, a_mn=mn,
            b_rank=4, b_dim=[co_c, wo, ho, n], b_stride=[e, co_c * e, wo * co_c * e, ho * wo * co_c * e],
            b_box=[64, pw, ph, pn], b_mx=[0, 1, 0, 0], b_my=[0, 0, 1, 0], b_mn=[0, 0, 0, 1],
            pw=pw, ph=ph, pn=pn, out_w=wo, out_h=ho, n_img=n, m_total=n_total, n_total=m_total, bn=bn,
            num_taps=len(taps), tap_off=taps, s_m=s_n, s_t=s_t, s_n=s_m, tap_on_a=1)
    bn = 128 if n_total % 128 == 0 else 64
    return WgradPlan(
        a_rank=4, a_dim=[co_c, wo, ho, n], a_stride=[e, co_c * e, wo * co_c * e, ho * wo * co_c * e],
        a_box=[64, pw, ph, pn], a_mx=[0, 1, 0, 0], a_my=[0, 0, 1, 0], a_mn=[0, 0, 0, 1],
        b_rank=rank, b_dim=dims, b_stride=strides, b_box=box(pw, ph, pn), b_mx=mx, b_my=my, b_mn=mn,
        pw=pw, ph=ph, pn=pn, out_w=wo, out_h=ho, n_img=n, m_total=m_total, n_total=n_total, bn=bn,
        num_taps=len(taps), tap_off=taps, s_m=s_m, s_t=s_t, s_n=s_n)


# ---------------------------------------------------------------------------- weight index maps
def fwd_index_map(cout, cin, kh, kw, co_rows, c_buf, kw_p=None, c_p=None):
    """Index map (into the channels_last parameter memory [cout][kh][kw][cin]) for the forward weight
    matrix [co_rows][K].  Default K order: (kh, kw, c_buf).  With kw_p/c_p given (kw-expanded first
    layers) K order is (kh, kw_p, c_p) -- one 64-wide chunk per kh tap."""
    import torch

    if kw_p is None:
        idx = torch.full((co_rows, kh, kw, c_buf), -1, dtype=torch.int32)
        src = torch.arange(cout * kh * kw * cin, dtype=torch.int32).view(cout, kh, kw, cin)
        idx[:cout, :, :, :cin] = src
    else:
        idx = torch.full((co_rows, kh, kw_p, c_p), -1, dtype=torch.int32)
        src = torch.arange(cout * kh * kw * cin, dtype=torch.int32).view(cout, kh, kw, cin)
        idx[:cout, :, :kw, :cin] = src
    return idx.reshape(-1)


def dgrad_index_map(cout, cin, kh, kw, sy, sx, c_rows, ck, kw_p=None, c_p=None):
    """Index map for the dgrad weight matrix [c_rows][phase][(a, b)][ck]:
    entry (ci, py*sx+px, a, b, co) <- w[co][sy*a+py][sx*b+px][ci].
    With kw_p/c_p (kw-expanded layers): rows are (kw_p, c_p) pairs and the kernel is (kh x 1)."""
    import torch

    src = torch.arange(cout * kh * kw * cin, dtype=torch.int32).view(cout, kh, kw, cin)
    if kw_p is None:
        na, nb = kh // sy, kw // sx
        idx = torch.full((c_rows, sy * sx, na, nb, ck), -1, dtype=torch.int32)
        for py in range(sy):
            for px in range(sx):
                for a in range(na):
                    for b in range(nb):
                        idx[:cin, py * sx + px, a, b, :cout] = src[:, sy * a + py, sx * b + px, :].t()
    else:
        na = kh // sy
        idx = torch.full((kw_p, c_p, sy, na, 1, ck), -1, dtype=torch.int32)
        for py in range(sy):
            for a in range(na):
                # rows (kw, c) <- w[co][sy*a+py][kw][c]
                idx[:kw, :cin, py, a, 0, :cout] = src[:, sy * a + py, :, :].permute(1, 2, 0)
        idx = idx.view(kw_p * c_p, sy, na, 1, ck)
    return idx.reshape(-1)


def rspace_index_maps(cout, cin, kh, kw, kwp=8, cop=4, ck=64):
    """Narrow-output layers (cout <= cop): weight matrices of the vertical (kh x 1) GEMM whose N index is
    (kw, co) = kw*cop + co.  Returns (fwd, dgrad, inv):
      fwd   [(kwp*cop)][kh][cin]          <- w[co][kh][kw][ci]
      dgrad [cin][kh][ck]  (K = (kw,co))  <- w[co][kh][kw][ci]
      inv   [cout*kh*kw*cin]              position of each parameter element inside `fwd` (wgrad scatter-back)."""
    import torch

    src = torch.arange(cout * kh * kw * cin, dtype=torch.int32).view(cout, kh, kw, cin)
    rows = kwp * cop
    fwd = torch.full((kwp, cop, kh, cin), -1, dtype=torch.int32)
    fwd[:kw, :cout] = src.permute(2, 0, 1, 3)
    dg = torch.full((cin, kh, ck), -1, dtype=torch.int32)
    dgv = dg[:, :, :rows].view(cin, kh, kwp, cop)
    dgv[:, :, :kw, :cout] = src.permute(3, 1, 2, 0)
    fwd = fwd.reshape(-1)
    inv = torch.full((cout * kh * kw * cin,), -1, dtype=torch.int32)
    m = fwd >= 0
    inv[fwd[m].long()] = torch.arange(fwd.numel(), dtype=torch.int32)[m]
    return fwd, dg.reshape(-1), inv
