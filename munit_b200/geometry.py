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


# ---------------------------------------------------------------------------- nearest-2x upsample + 5x5 conv as phase GEMMs
# nn.Upsample(scale_factor=2) -> ReflectionPad2d(2) -> Conv2d(k=5) (networks.py:534-545) reads every low-res pixel four
# times.  Per axis, output row 2i sees the low-res rows (i-1, i, i+1) with the tap sums (w0+w1, w2+w3, w4), row 2i+1
# sees them with (w0, w1+w2, w3+w4): four 3x3 "phase" convolutions on the low-res input, 9 MACs per output instead of
# 25.  With *replicate* padding of the low-res input this equals reflect padding of the up-sampled image everywhere
# except the outermost output row / column on each side (there the two mirrored up-sampled pixels belong to different
# low-res pixels), which take their own tap sums: top/left (0, w1+w2+w3, w0+w4), bottom/right mirrored.  So: 4 row
# types x 4 column types = 16 weight sets of 3x3 taps; one 4-phase launch writes every output with the interior sets,
# then four thin launches (rows 0 and 2H-1, columns 0 and 2W-1) and four one-pixel launches (corners) overwrite the
# ring.  EXPERIMENTAL (no-grad forward only, MUNIT_UPCONV_PHASE=1): the plans below are verified against F.conv2d by
# tests/test_geometry.py through the descriptor emulation; GPU validation is the next round's first step.
UP_ROW_TYPES = {
    # type -> 3 x 5 matrix A[d][k]: low-res row (i - 1 + d) collects tap k
    0: [[1, 1, 0, 0, 0], [0, 0, 1, 1, 0], [0, 0, 0, 0, 1]],   # even output row 2i (interior)
    1: [[1, 0, 0, 0, 0], [0, 1, 1, 0, 0], [0, 0, 0, 1, 1]],   # odd output row 2i+1 (interior)
    2: [[0, 0, 0, 0, 0], [0, 1, 1, 1, 0], [1, 0, 0, 0, 1]],   # output row 0
    3: [[1, 0, 0, 0, 1], [0, 1, 1, 1, 0], [0, 0, 0, 0, 0]],   # output row 2H-1
}


def upconv_phase_weights(weight):
    """weight [Co, Ci, 5, 5] (any float dtype) -> [Co, 16, 3, 3, Ci]: the tap sums of every (row type, column type),
    type index = 4 * row_type + col_type.  Pure tensor arithmetic (runs on whatever device `weight` lives on)."""
    import torch

    co, ci, kh, kw = weight.shape
    assert kh == 5 and kw == 5, "phase decomposition is built for the 5x5 decoder layers"
    a = torch.tensor([UP_ROW_TYPES[t] for t in range(4)], dtype=weight.dtype, device=weight.device)  # [4, 3, 5]
    w = weight.permute(0, 2, 3, 1)  # [Co, ky, kx, Ci]
    rows = (a[None, :, :, :, None, None] * w[:, None, None, :, :, :]).sum(3)          # [Co, rt, dy, kx, Ci]
    full = (a[None, None, None, :, :, :, None] * rows[:, :, :, None, None, :, :]).sum(5)  # [Co, rt, dy, ct, dx, Ci]
    return full.permute(0, 1, 3, 2, 4, 5).reshape(co, 16, 3, 3, ci).contiguous()


def plan_upconv_phases(n, h, w, c, co_rows, out_geom) -> List[TapGemmPlan]:
    """Launch list for y = conv5x5(reflect_pad2(nearest_up2(x))).  x is given as [n, h+2, w+2, c] with a *replicate*
    halo of 1; the weight matrix is [co_rows][16][3][3][c] (upconv_phase_weights, K contiguous); out_geom as in
    plan_fwd for the [2h, 2w] output.  Launch in list order (later launches overwrite the ring)."""
    assert c % 64 == 0 and h >= 2 and w >= 2
    e = 2
    dims = [c, w + 2, h + 2, n]
    strides = [e, c * e, (w + 2) * c * e, (h + 2) * (w + 2) * c * e]
    o_sn, o_sy, o_sx, y_off, x_off = out_geom
    bn = pick_bn(co_rows)
    kper = 9 * c

    def launch(i0, nh, j0, nw, phases):
        tw, th, tn = pick_tile(nw, nh, n, 128)
        taps = [[0, j0 + dx, i0 + dy, 0] for dy in range(3) for dx in range(3)]
        return TapGemmPlan(
            a_rank=4, a_dim=dims, a_stride=strides, a_box=[64, tw, th, tn], b_rows=co_rows, b_k=16 * kper, bn=bn,
            tw=tw, th=th, tn=tn, out_w=nw, out_h=nh, n_img=n, mx=[0, 1, 0, 0], my=[0, 0, 1, 0], mn=[0, 0, 0, 1],
            num_taps=9, chunks=c // 64, tap_off=taps, phases=len(phases),
            b_k0=[(4 * rt + ct) * kper for rt, ct, _, _ in phases],
            o_yoff=[2 * i0 + py + y_off for _, _, py, _ in phases], o_xoff=[2 * j0 + px + x_off for _, _, _, px in phases],
            o_sn=o_sn, o_sy=o_sy, o_sx=o_sx, o_ymul=2, o_xmul=2, n_store=co_rows)

    plans = [launch(0, h, 0, w, [(py, px, py, px) for py in (0, 1) for px in (0, 1)])]   # interior sets everywhere
    plans.append(launch(0, 1, 0, w, [(2, 0, 0, 0), (2, 1, 0, 1)]))                          # output row 0
    plans.append(launch(h - 1, 1, 0, w, [(3, 0, 1, 0), (3, 1, 1, 1)]))                      # output row 2h-1
    plans.append(launch(0, h, 0, 1, [(0, 2, 0, 0), (1, 2, 1, 0)]))                          # output column 0
    plans.append(launch(0, h, w - 1, 1, [(0, 3, 0, 1), (1, 3, 1, 1)]))                      # output column 2w-1
    for i0, rt, py in ((0, 2, 0), (h - 1, 3, 1)):                                           # corners
        for j0, ct, px in ((0, 2, 0), (w - 1, 3, 1)):
            plans.append(launch(i0, 1, j0, 1, [(rt, ct, py, px)]))
    flops = 2.0 * n * (2 * h) * (2 * w) * co_rows * 25 * c
    for p in plans:
        p.alg_flops = 0.0
    plans[0].alg_flops = flops  # direct-form work of the layer, booked on the main launch
    return plans


def upconv_dgrad_index_map(co, ci, c_rows, ck):
    """Index map into the [Co][16][3][3][Ci] phase-weight tensor for the interior dgrad matrix of the phase form:
    [c_rows (ci)][36 taps = (py, px, dy, dx)][ck (co)]  <-  wph[co, 4*py + px, dy, dx, ci]."""
    import torch

    src = torch.arange(co * 16 * 9 * ci, dtype=torch.int32).view(co, 16, 3, 3, ci)
    idx = torch.full((c_rows, 2, 2, 3, 3, ck), -1, dtype=torch.int32)
    for py in range(2):
        for px in range(2):
            idx[:ci, py, px, :, :, :co] = src[:, 4 * py + px].permute(3, 1, 2, 0)
    return idx.reshape(-1)


def plan_upconv_dgrad_interior(n, h, w, c_rows, co_c) -> TapGemmPlan:
    """Interior part of the phase-form input gradient: dy0 [n, 2h, 2w, co_c] (the gradient of the conv output with
    its outermost ring ZEROED -- the ring pixels use other tap sums and are handled separately) ->
    dxr [n, h+2, w+2, c_rows], the gradient w.r.t. the replicate-padded low-res input (every position written).
    dy0 is read through a space-to-depth view [2*co, w, 2, h, n]; 36 taps (py, px, dy, dx)."""
    assert co_c % 64 == 0
    e = 2
    dims = [2 * co_c, w, 2, h, n]
    strides = [e, 2 * co_c * e, 2 * w * co_c * e, 4 * w * co_c * e, 4 * h * w * co_c * e]
    oh, ow = h + 2, w + 2
    tw, th, tn = pick_tile(ow, oh, n, 128)
    taps = [[px * co_c, -dx, py, -dy, 0] for py in range(2) for px in range(2) for dy in range(3) for dx in range(3)]
    return TapGemmPlan(
        a_rank=5, a_dim=dims, a_stride=strides, a_box=[64, tw, 1, th, tn], b_rows=c_rows, b_k=36 * co_c,
        bn=pick_bn(c_rows), tw=tw, th=th, tn=tn, out_w=ow, out_h=oh, n_img=n, mx=[0, 1, 0, 0, 0], my=[0, 0, 0, 1, 0],
        mn=[0, 0, 0, 0, 1], num_taps=36, chunks=co_c // 64, tap_off=taps, phases=1, b_k0=[0], o_yoff=[0], o_xoff=[0],
        o_sn=oh * ow * c_rows, o_sy=ow * c_rows, o_sx=c_rows, o_ymul=1, o_xmul=1, n_store=c_rows)


def plan_upconv_wgrad_interior(n, h, w, c, co_c, py, px) -> WgradPlan:
    """Interior part of the phase-form weight gradient for output phase (py, px): accumulates
    dwph[co][4*py + px][dy][dx][ci] += sum_{i,j} dy0[2i+py, 2j+px][co] * xr[i+dy, j+dx][ci] into the phase-gradient
    scratch [co][16][3][3][ci] (dy0 ring-zeroed as in plan_upconv_dgrad_interior; xr = replicate-padded low-res
    input [n, h+2, w+2, c]).  The caller offsets the dy0 base pointer by (py*2w + px)*co_c elements; the view
    below then walks every second row / column."""
    e = 2
    pw, ph, pn = pick_tile(w, h, n, 64)
    taps = [[0, dx, dy, 0] for dy in range(3) for dx in range(3)]
    return WgradPlan(
        a_rank=4, a_dim=[co_c, w, h, n], a_stride=[e, 2 * co_c * e, 4 * w * co_c * e, 4 * h * w * co_c * e],
        a_box=[64, pw, ph, pn], a_mx=[0, 1, 0, 0], a_my=[0, 0, 1, 0], a_mn=[0, 0, 0, 1],
        b_rank=4, b_dim=[c, w + 2, h + 2, n], b_stride=[e, c * e, (w + 2) * c * e, (h + 2) * (w + 2) * c * e],
        b_box=[64, pw, ph, pn], b_mx=[0, 1, 0, 0], b_my=[0, 0, 1, 0], b_mn=[0, 0, 0, 1],
        pw=pw, ph=ph, pn=pn, out_w=w, out_h=h, n_img=n, m_total=co_c, n_total=c, bn=128 if c % 128 == 0 else 64,
        num_taps=9, tap_off=taps, s_m=16 * 9 * c, s_t=c, s_n=1)


# Ring share of the phase-form backward.  The outermost output row / column on each side is cut out of dY into
# strips with the four corner pixels zeroed (they have their own (first|last, first|last) types and are too few for
# a GEMM):   top / bottom strip [n, 2w, co]   (output rows 0 and 2h-1),   left / right strip [n, 2h, co].
# side: 0 top, 1 bottom, 2 left, 3 right; its row (or column) type is 2 (first) for top / left, 3 (last) otherwise.
def upconv_ring_dgrad_index_map(co, ci, c_rows, ck, side):
    """[c_rows (ci)][3 phases (the strip-normal offset d)][6 taps = (p, e)][ck (co)]: p the parity along the strip,
    e the tap along the strip  <-  wph[co, type, dy, dx, ci] with (dy, dx) = (d, e) for top / bottom, (e, d) for
    left / right and type = 4*rt + p (top / bottom) or 4*p + ct (left / right)."""
    import torch

    src = torch.arange(co * 16 * 9 * ci, dtype=torch.int32).view(co, 16, 3, 3, ci)
    t = 2 if side in (0, 2) else 3
    idx = torch.full((c_rows, 3, 2, 3, ck), -1, dtype=torch.int32)
    for d in range(3):
        for p in range(2):
            for e_ in range(3):
                if side < 2:
                    v = src[:, 4 * t + p, d, e_]
                else:
                    v = src[:, 4 * p + t, e_, d]
                idx[:ci, d, p, e_, :co] = v.t()
    return idx.reshape(-1)


def plan_upconv_dgrad_ring(n, h, w, c_rows, co_c, side) -> TapGemmPlan:
    """strip -> its contribution to the gradient of the replicate-padded low-res input, as a 3-wide band:
    top / bottom: band [n, 3, w+2, c_rows] = rows (0..2) / (h-1 .. h+1) of dxr; left / right: [n, h+2, 3, c_rows] =
    columns (0..2) / (w-1 .. w+1).  Phase d writes band line d; 6 taps (parity, tap along the strip)."""
    assert co_c % 64 == 0
    e = 2
    horiz = side < 2
    length = w if horiz else h            # low-res extent along the strip
    dims = [2 * co_c, length, 1, n]
    strides = [e, 2 * co_c * e, 2 * length * co_c * e, 2 * length * co_c * e]
    ol = length + 2
    taps = [[p * co_c, -e_, 0, 0] for p in range(2) for e_ in range(3)]
    if horiz:
        ow, oh = ol, 1
        o_sn, o_sy, o_sx = 3 * ol * c_rows, ol * c_rows, c_rows
        mx, my = [0, 1, 0, 0], [0, 0, 1, 0]
        yoff, xoff = [0, 1, 2], [0, 0, 0]
    else:
        ow, oh = 1, ol
        o_sn, o_sy, o_sx = ol * 3 * c_rows, 3 * c_rows, c_rows
        mx, my = [0, 0, 1, 0], [0, 1, 0, 0]   # the GEMM's y walks the strip
        yoff, xoff = [0, 0, 0], [0, 1, 2]
    tw, th, tn = pick_tile(ow, oh, n, 128)
    box = [64, tw, 1, tn] if horiz else [64, th, 1, tn]
    return TapGemmPlan(
        a_rank=4, a_dim=dims, a_stride=strides, a_box=box, b_rows=c_rows, b_k=3 * 6 * co_c, bn=pick_bn(c_rows),
        tw=tw, th=th, tn=tn, out_w=ow, out_h=oh, n_img=n, mx=mx, my=my, mn=[0, 0, 0, 1], num_taps=6,
        chunks=co_c // 64, tap_off=taps, phases=3, b_k0=[d * 6 * co_c for d in range(3)], o_yoff=yoff, o_xoff=xoff,
        o_sn=o_sn, o_sy=o_sy, o_sx=o_sx, o_ymul=1, o_xmul=1, n_store=c_rows)


def plan_upconv_wgrad_ring(n, h, w, c, co_c, side, p) -> WgradPlan:
    """Ring share of the phase-weight gradient: strip pixels of parity p along the strip against the three low-res
    rows (columns) they read.  Accumulates into the [co][16][3][3][ci] scratch at type 4*rt + p (top / bottom) or
    4*p + ct (left / right); the caller offsets the strip base pointer by p*co_c elements and the scratch pointer by
    type*9*c.  xr: replicate-padded low-res input [n, h+2, w+2, c]."""
    e = 2
    horiz = side < 2
    length = w if horiz else h
    line0 = 0 if side in (0, 2) else ((h - 1) if horiz else (w - 1))   # first of the three xr rows / columns read
    if horiz:
        pw, ph, pn = pick_tile(length, 1, n, 64)
        a_mx, a_my = [0, 1, 0, 0], [0, 0, 1, 0]
        taps = [[0, dx, line0 + dy, 0] for dy in range(3) for dx in range(3)]
        a_box = [64, pw, 1, pn]
        ow, oh = length, 1
    else:
        pw, ph, pn = pick_tile(1, length, n, 64)
        a_mx, a_my = [0, 0, 1, 0], [0, 1, 0, 0]
        taps = [[0, line0 + dx, dy, 0] for dy in range(3) for dx in range(3)]
        a_box = [64, ph, 1, pn]
        ow, oh = 1, length
    return WgradPlan(
        a_rank=4, a_dim=[co_c, length, 1, n], a_stride=[e, 2 * co_c * e, 2 * length * co_c * e, 2 * length * co_c * e],
        a_box=a_box, a_mx=a_mx, a_my=a_my, a_mn=[0, 0, 0, 1],
        b_rank=4, b_dim=[c, w + 2, h + 2, n], b_stride=[e, c * e, (w + 2) * c * e, (h + 2) * (w + 2) * c * e],
        b_box=[64, pw, ph, pn], b_mx=[0, 1, 0, 0], b_my=[0, 0, 1, 0], b_mn=[0, 0, 0, 1],
        pw=pw, ph=ph, pn=pn, out_w=ow, out_h=oh, n_img=n, m_total=co_c, n_total=c, bn=128 if c % 128 == 0 else 64,
        num_taps=9, tap_off=taps, s_m=16 * 9 * c, s_t=c, s_n=1)
