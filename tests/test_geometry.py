"""Host-side conv planning (munit_b200/geometry.py) checked against F.conv2d on the CPU through an
emulation of the descriptor semantics.  No GPU, no compute through the C-ABI."""
import pytest
import torch
import torch.nn.functional as F

from munit_b200 import geometry as G
from tests import emulate as E

CASES = [
    # n, h, w, cin, cout, k, s, pad
    (2, 8, 8, 64, 64, 3, 1, 1),
    (1, 6, 10, 64, 128, 5, 1, 2),
    (2, 8, 8, 64, 128, 4, 2, 1),
    (1, 4, 4, 128, 64, 4, 2, 1),
    (1, 9, 7, 64, 16, 7, 1, 3),
    (3, 4, 4, 64, 64, 1, 1, 0),
    (1, 6, 6, 128, 64, 5, 1, 2),  # cout < 128 <= cin: swapped wgrad orientation
    (1, 8, 8, 128, 64, 4, 2, 1),
]


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n,h,w,cin,cout,k,s,pad", CASES)
def test_fwd_dgrad_wgrad_plans(n, h, w, cin, cout, k, s, pad):
    torch.manual_seed(0)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) * 0.1
    bias = torch.randn(cout)
    xp = F.pad(x, (pad,) * 4, mode="reflect") if pad else x
    xp = xp.clone().requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y_ref = F.conv2d(xp, wt_r, bias, stride=s)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    hp, wp = h + 2 * pad, w + 2 * pad
    ho, wo = y_ref.shape[2:]
    # ---- forward
    plan = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0))
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, -1)  # channels_last memory == [co][kh][kw][ci]
    out = torch.zeros(n * ho * wo * cout)
    E.tapgemm(plan, nhwc(xp.detach()).reshape(-1), wmat, out, bias, "none")
    assert torch.allclose(out.view(n, ho, wo, cout), nhwc(y_ref.detach()), atol=1e-3, rtol=1e-4)
    # ---- dgrad (w.r.t. the padded input)
    ck = max(64, cout)
    dplan = G.plan_dgrad(n, hp, wp, cin, k, k, s, s, cout)
    idx = G.dgrad_index_map(cout, cin, k, k, s, s, cin, ck)
    wflat = wt.permute(0, 2, 3, 1).reshape(-1)
    wd = torch.where(idx >= 0, wflat[idx.clamp(min=0).long()], torch.zeros(())).view(cin, -1)
    dxp = torch.full((n * hp * wp * cin,), float("nan"))
    E.tapgemm(dplan, nhwc(gy).reshape(-1), wd, dxp)
    assert torch.allclose(dxp.view(n, hp, wp, cin), nhwc(xp.grad), atol=1e-3, rtol=1e-4)
    # ---- wgrad into the channels_last gradient layout [co][kh][kw][ci]
    wplan = G.plan_wgrad(n, hp, wp, cin, k, k, s, s, cout, cout, k * k * cin, cin, 1)
    dw = torch.zeros(cout * k * k * cin)
    E.wgrad(wplan, nhwc(gy).reshape(-1), nhwc(xp.detach()).reshape(-1), dw)
    assert torch.allclose(dw.view(cout, k, k, cin), wt_r.grad.permute(0, 2, 3, 1), atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout,k,zp", [(2, 8, 8, 64, 128, 3, 1), (1, 5, 7, 128, 64, 3, 1),
                                                  (2, 4, 4, 64, 64, 1, 0), (1, 6, 6, 64, 64, 5, 2)])
def test_zero_padded_conv_plans(n, h, w, cin, cout, k, zp):
    """conv3x3 / conv1x1 of the domain-classifier BasicBlock (utils.py:1238-1274): zero padding is implicit --
    unpadded buffers, taps starting at -zp, out-of-range reads are zero (TMA fill)."""
    torch.manual_seed(1)
    x = torch.randn(n, cin, h, w, requires_grad=True)
    wt = (torch.randn(cout, cin, k, k) * 0.1).requires_grad_(True)
    y_ref = F.conv2d(x, wt, None, padding=zp)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    ho, wo = y_ref.shape[2:]
    assert (ho, wo) == (h, w)
    plan = G.plan_fwd(n, h, w, cin, k, k, 1, 1, cout, (ho * wo * cout, wo * cout, cout, 0, 0), zpad=zp)
    wmat = wt.detach().permute(0, 2, 3, 1).reshape(cout, -1)
    out = torch.zeros(n * ho * wo * cout)
    E.tapgemm(plan, nhwc(x.detach()).reshape(-1), wmat, out)
    assert torch.allclose(out.view(n, ho, wo, cout), nhwc(y_ref.detach()), atol=1e-3, rtol=1e-4)
    ck = max(64, cout)
    dplan = G.plan_dgrad(n, h, w, cin, k, k, 1, 1, cout, zpad=zp)
    idx = G.dgrad_index_map(cout, cin, k, k, 1, 1, cin, ck)
    wflat = wt.detach().permute(0, 2, 3, 1).reshape(-1)
    wd = torch.where(idx >= 0, wflat[idx.clamp(min=0).long()], torch.zeros(())).view(cin, -1)
    dx = torch.full((n * h * w * cin,), float("nan"))
    E.tapgemm(dplan, nhwc(gy).reshape(-1), wd, dx)
    assert torch.allclose(dx.view(n, h, w, cin), nhwc(x.grad), atol=1e-3, rtol=1e-4)
    wplan = G.plan_wgrad(n, h, w, cin, k, k, 1, 1, cout, cout, k * k * cin, cin, 1, zpad=zp)
    dw = torch.zeros(cout * k * k * cin)
    E.wgrad(wplan, nhwc(gy).reshape(-1), nhwc(x.detach()).reshape(-1), dw)
    assert torch.allclose(dw.view(cout, k, k, cin), wt.grad.permute(0, 2, 3, 1), atol=1e-2, rtol=1e-3)


def test_fwd_into_padded_output():
    n, h, w, c, k, pad, po = 1, 8, 8, 64, 3, 1, 2
    x = torch.randn(n, c, h, w)
    wt = torch.randn(c, c, k, k) * 0.1
    xp = F.pad(x, (pad,) * 4, mode="reflect")
    y = F.conv2d(xp, wt)
    hop, wop = h + 2 * po, w + 2 * po
    plan = G.plan_fwd(n, h + 2, w + 2, c, k, k, 1, 1, c, (hop * wop * c, wop * c, c, po, po))
    out = torch.zeros(n * hop * wop * c)
    E.tapgemm(plan, nhwc(xp).reshape(-1), wt.permute(0, 2, 3, 1).reshape(c, -1), out)
    assert torch.allclose(out.view(n, hop, wop, c)[:, po:-po, po:-po], nhwc(y), atol=1e-3)


@pytest.mark.parametrize("k,s,pad,h", [(7, 1, 3, 8), (4, 2, 1, 8)])
def test_kwexp_first_layer(k, s, pad, h):
    """Cin=3 layers: image -> kw-expanded buffer E[n,yp,xo,kw,c] -> (kh x 1) tap GEMM with K=64."""
    n, cin, cout, w = 2, 3, 64, h
    kwp, cp = (8, 8) if k == 7 else (4, 16)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) * 0.1
    xp = F.pad(x, (pad,) * 4, mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y = F.conv2d(xp, wt_r, stride=s)
    gy = torch.randn_like(y)
    y.backward(gy)
    hp = h + 2 * pad
    ho, wo = y.shape[2:]
    e = torch.zeros(n, hp, wo, kwp, cp)
    for kw in range(k):
        e[:, :, :, kw, :cin] = xp.detach()[:, :, :, kw: kw + s * (wo - 1) + 1: s].permute(0, 2, 3, 1)
    plan = G.plan_fwd(n, hp, wo, 64, k, 1, s, 1, cout, (ho * wo * cout, wo * cout, cout, 0, 0))
    idx = G.fwd_index_map(cout, cin, k, k, cout, 64, kwp, cp)
    wflat = wt.permute(0, 2, 3, 1).reshape(-1)
    wm = torch.where(idx >= 0, wflat[idx.clamp(min=0).long()], torch.zeros(())).view(cout, -1)
    out = torch.zeros(n * ho * wo * cout)
    E.tapgemm(plan, e.reshape(-1), wm, out)
    assert torch.allclose(out.view(n, ho, wo, cout), nhwc(y.detach()), atol=1e-3)
    # dgrad in E space, then fold back to the padded image
    dplan = G.plan_dgrad(n, hp, wo, 64, k, 1, s, 1, cout)
    didx = G.dgrad_index_map(cout, cin, k, k, s, 1, 64, 64, kwp, cp)
    wd = torch.where(didx >= 0, wflat[didx.clamp(min=0).long()], torch.zeros(())).view(64, -1)
    de = torch.zeros(n * hp * wo * 64)
    E.tapgemm(dplan, nhwc(gy).reshape(-1), wd, de)
    de = de.view(n, hp, wo, kwp, cp)
    dxp = torch.zeros(n, cin, hp, w + 2 * pad)
    for kw in range(k):
        dxp[:, :, :, kw: kw + s * (wo - 1) + 1: s] += de[:, :, :, kw, :cin].permute(0, 3, 1, 2)
    assert torch.allclose(dxp, xp.grad, atol=1e-3)
    # wgrad in E space
    wplan = G.plan_wgrad(n, hp, wo, 64, k, 1, s, 1, cout, cout, k * 64, 64, 1)
    dwe = torch.zeros(cout * k * 64)
    E.wgrad(wplan, nhwc(gy).reshape(-1), e.reshape(-1), dwe)
    dw = torch.zeros(cout * k * k * cin)
    m = idx >= 0
    dw.index_add_(0, idx[m].long(), dwe[m])
    assert torch.allclose(dw.view(cout, k, k, cin), wt_r.grad.permute(0, 2, 3, 1), atol=1e-2, rtol=1e-3)


def test_pick_tile():
    assert G.pick_tile(64, 64, 8, 128) == (64, 2, 1)
    tw, th, tn = G.pick_tile(4, 4, 8, 128)
    assert tw * th * tn == 128 and tw <= 4 and th <= 4
    tw, th, tn = G.pick_tile(66, 66, 1, 128)
    assert tw * th * tn == 128


def test_rspace_narrow_output_layer():
    """64 -> 3 7x7 layer as a vertical (7x1) GEMM with N = (kw, co) plus a horizontal combine."""
    n, cin, cout, k, pad, h, w = 2, 64, 3, 7, 3, 6, 9
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, k, k) * 0.1
    xp = F.pad(x, (pad,) * 4, mode="reflect").requires_grad_(True)
    wt_r = wt.clone().requires_grad_(True)
    y = F.conv2d(xp, wt_r)
    gy = torch.randn_like(y)
    y.backward(gy)
    hp, wp = h + 2 * pad, w + 2 * pad
    fwd_i, dg_i, inv_i = G.rspace_index_maps(cout, cin, k, k)
    wflat = wt.permute(0, 2, 3, 1).reshape(-1)
    pick = lambda idx: torch.where(idx >= 0, wflat[idx.clamp(min=0).long()], torch.zeros(()))
    plan = G.plan_fwd(n, hp, wp, 64, k, 1, 1, 1, 32, (h * wp * 32, wp * 32, 32, 0, 0))
    r = torch.zeros(n * h * wp * 32)
    E.tapgemm(plan, nhwc(xp.detach()).reshape(-1), pick(fwd_i).view(32, -1), r)
    r = r.view(n, h, wp, 8, 4)
    out = torch.zeros(n, cout, h, w)
    for kw in range(k):
        out += r[:, :, kw:kw + w, kw, :cout].permute(0, 3, 1, 2)
    assert torch.allclose(out, y.detach(), atol=1e-3)
    # backward: expand g -> dR, vertical dgrad, wgrad + scatter-back
    dr = torch.zeros(n, h, wp, 8, 4)
    for kw in range(k):
        dr[:, :, kw:kw + w, kw, :cout] = gy.permute(0, 2, 3, 1)
    dplan = G.plan_dgrad(n, hp, wp, 64, k, 1, 1, 1, 32)
    dxp = torch.zeros(n * hp * wp * 64)
    E.tapgemm(dplan, dr.reshape(-1), pick(dg_i).view(64, -1), dxp)
    assert torch.allclose(dxp.view(n, hp, wp, 64), nhwc(xp.grad), atol=1e-3)
    wplan = G.plan_wgrad(n, hp, wp, 64, k, 1, 1, 1, 32, 32, k * 64, 64, 1)
    tmp = torch.zeros(32 * k * 64)
    E.wgrad(wplan, dr.reshape(-1), nhwc(xp.detach()).reshape(-1), tmp)
    dw = torch.where(inv_i >= 0, tmp[inv_i.clamp(min=0).long()], torch.zeros(()))
    assert torch.allclose(dw.view(cout, k, k, cin), wt_r.grad.permute(0, 2, 3, 1), atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 4, 6, 64, 64), (1, 8, 8, 128, 64), (3, 2, 2, 64, 128), (1, 5, 3, 64, 16)])
def test_upsample_5x5_as_phase_gemms(n, h, w, cin, cout):
    """nearest-2x upsample + reflect pad 2 + 5x5 conv (networks.py:534-545) == one 4-phase 3x3 launch on the
    replicate-padded low-res input plus eight ring launches with their own tap sums (geometry.plan_upconv_phases)."""
    torch.manual_seed(4)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 5, 5) * 0.1
    bias = torch.randn(cout)
    up = F.interpolate(x, scale_factor=2, mode="nearest")
    y_ref = F.conv2d(F.pad(up, (2,) * 4, mode="reflect"), wt, bias)
    xr = F.pad(x, (1,) * 4, mode="replicate")
    wph = G.upconv_phase_weights(wt)
    assert tuple(wph.shape) == (cout, 16, 3, 3, cin)
    # interior sets: even/even phase sums the 2x2 taps (0..1, 0..1) into its first entry
    assert torch.allclose(wph[:, 0, 0, 0], wt[:, :, 0:2, 0:2].sum((2, 3)), atol=1e-6)
    ho, wo = 2 * h, 2 * w
    plans = G.plan_upconv_phases(n, h, w, cin, cout, (ho * wo * cout, wo * cout, cout, 0, 0))
    assert len(plans) == 9 and plans[0].phases == 4
    out = torch.full((n * ho * wo * cout,), float("nan"))
    for p in plans:
        E.tapgemm(p, nhwc(xr).reshape(-1), wph.reshape(cout, -1), out, bias, "none")
    got = out.view(n, ho, wo, cout)
    assert torch.allclose(got, nhwc(y_ref), atol=2e-3, rtol=1e-4), float((got - nhwc(y_ref)).abs().max())
    # the interior launch alone is already exact away from the ring
    out1 = torch.zeros(n * ho * wo * cout)
    E.tapgemm(plans[0], nhwc(xr).reshape(-1), wph.reshape(cout, -1), out1, bias, "none")
    assert torch.allclose(out1.view(n, ho, wo, cout)[:, 1:-1, 1:-1], nhwc(y_ref)[:, 1:-1, 1:-1], atol=2e-3, rtol=1e-4)


def test_upsample_phase_gemm_issues_a_third_of_the_macs():
    """At the decoder's real shapes the phase form issues 36 % of the direct form's MACs (+ the thin ring launches)."""
    for n, h, c, co in ((8, 64, 256, 128), (8, 128, 128, 64)):
        plans = G.plan_upconv_phases(n, h, h, c, co, (4 * h * h * co, 2 * h * co, co, 0, 0))
        direct = 2.0 * n * (2 * h) ** 2 * co * 25 * c
        assert plans[0].alg_flops == direct
        assert abs(plans[0].flops() / direct - 0.36) < 1e-6
        assert sum(p.flops() for p in plans[1:]) < 0.04 * direct


def _ring_zeroed(g):
    g = g.clone()
    g[:, :, 0] = 0; g[:, :, -1] = 0; g[:, :, :, 0] = 0; g[:, :, :, -1] = 0
    return g


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 4, 6, 64, 64), (1, 5, 3, 128, 64), (1, 4, 4, 64, 128)])
def test_upsample_phase_form_interior_backward_plans(n, h, w, cin, cout):
    """Interior share of the phase-form backward (EXPERIMENTAL blueprint, see tests/test_upconv_math_cpu.py): with the
    ring of dY zeroed, one 36-tap launch over the space-to-depth view of dY gives the gradient of the replicate-padded
    low-res input, and four 9-tap wgrad launches give the interior phase-weight gradients -- both equal autograd of
    the direct formulation restricted to the interior output pixels."""
    torch.manual_seed(5)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 5, 5) * 0.1
    xr = F.pad(x, (1,) * 4, mode="replicate").requires_grad_(True)
    wph = G.upconv_phase_weights(wt).requires_grad_(True)                      # [co, 16, 3, 3, ci]
    # interior-type forward as plain convs (autograd reference for the interior share)
    y = torch.zeros(n, cout, 2 * h, 2 * w)
    parts = {}
    for py in (0, 1):
        for px in (0, 1):
            parts[(py, px)] = F.conv2d(xr, wph[:, 4 * py + px].permute(0, 3, 1, 2))
    y = torch.stack([torch.stack([parts[(py, px)] for px in (0, 1)], -1) for py in (0, 1)], -3)  # [n,co,h,2,w,2]
    y = y.reshape(n, cout, 2 * h, 2 * w)
    gy0 = _ring_zeroed(torch.randn(n, cout, 2 * h, 2 * w))
    y.backward(gy0)
    # ---- dgrad
    ck = max(64, cout)
    plan = G.plan_upconv_dgrad_interior(n, h, w, cin, cout)
    idx = G.upconv_dgrad_index_map(cout, cin, cin, ck)
    flat = wph.detach().reshape(-1)
    wd = torch.where(idx >= 0, flat[idx.clamp(min=0).long()], torch.zeros(())).view(cin, -1)
    assert ck == cout, "test shapes keep co a multiple of 64"
    dxr = torch.full((n * (h + 2) * (w + 2) * cin,), float("nan"))
    E.tapgemm(plan, nhwc(gy0).reshape(-1), wd, dxr)
    assert torch.allclose(dxr.view(n, h + 2, w + 2, cin), nhwc(xr.grad), atol=2e-3, rtol=1e-4)
    # ---- wgrad, one launch per output phase, into the [co][16][3][3][ci] scratch
    dw = torch.zeros(cout * 16 * 9 * cin)
    g_flat = nhwc(gy0).reshape(-1)
    for py in (0, 1):
        for px in (0, 1):
            wp_ = G.plan_upconv_wgrad_interior(n, h, w, cin, cout, py, px)
            base = (py * 2 * w + px) * cout
            E.wgrad(wp_, g_flat[base:], nhwc(xr.detach()).reshape(-1), dw[(4 * py + px) * 9 * cin:])
    got = dw.view(cout, 16, 3, 3, cin)
    assert torch.allclose(got[:, [0, 1, 4, 5]], wph.grad[:, [0, 1, 4, 5]], atol=1e-2, rtol=1e-3)
    assert float(got[:, [2, 3, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]].abs().max()) == 0  # ring types untouched here


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 4, 6, 64, 64), (1, 5, 3, 64, 128)])
def test_upsample_phase_form_ring_backward_plans(n, h, w, cin, cout):
    """Ring share of the phase-form backward (EXPERIMENTAL blueprint): the four strips (corners zeroed) through
    geometry.plan_upconv_dgrad_ring / plan_upconv_wgrad_ring against autograd of the type-form forward restricted to
    the non-corner ring pixels."""
    torch.manual_seed(6)
    x = torch.randn(n, cin, h, w)
    wt = torch.randn(cout, cin, 5, 5) * 0.1
    xr = F.pad(x, (1,) * 4, mode="replicate").requires_grad_(True)
    wph = G.upconv_phase_weights(wt).requires_grad_(True)
    gy = torch.randn(n, cout, 2 * h, 2 * w)
    ring = torch.zeros(2 * h, 2 * w, dtype=torch.bool)
    ring[0] = ring[-1] = True
    ring[:, 0] = ring[:, -1] = True
    for cy in (0, -1):
        for cx in (0, -1):
            ring[cy, cx] = False                                   # corners are handled apart
    gy_ring = gy * ring
    # type-form forward of the four strips (autograd reference)
    loss = 0
    def line(rt_or_ct, horiz, out_line, lo_line):
        """sum over the strip's pixels of gy * y, y through the type weights."""
        nonlocal loss
        L = w if horiz else h
        for p in (0, 1):
            t = (4 * rt_or_ct + p) if horiz else (4 * p + rt_or_ct)
            k = wph[:, t].permute(0, 3, 1, 2)                       # [co, ci, 3, 3]
            if horiz:
                yl = F.conv2d(xr[:, :, lo_line:lo_line + 3, :], k)  # [n, co, 1, w]
                g = gy_ring[:, :, out_line, p::2].unsqueeze(2)
            else:
                yl = F.conv2d(xr[:, :, :, lo_line:lo_line + 3], k)  # [n, co, h, 1]
                g = gy_ring[:, :, p::2, out_line].unsqueeze(3)
            loss = loss + (yl * g).sum()
    line(2, True, 0, 0); line(3, True, 2 * h - 1, h - 1); line(2, False, 0, 0); line(3, False, 2 * w - 1, w - 1)
    loss.backward()
    strips = {0: gy_ring[:, :, 0, :], 1: gy_ring[:, :, -1, :], 2: gy_ring[:, :, :, 0], 3: gy_ring[:, :, :, -1]}  # [n,co,L2]
    dxr = torch.zeros(n, h + 2, w + 2, cin)
    dw = torch.zeros(cout * 16 * 9 * cin)
    flat_w = wph.detach().reshape(-1)
    xr_flat = nhwc(xr.detach()).reshape(-1)
    for side, s in strips.items():
        sf = s.permute(0, 2, 1).contiguous().reshape(-1)            # [n, 2L, co]
        L = w if side < 2 else h
        plan = G.plan_upconv_dgrad_ring(n, h, w, cin, cout, side)
        idx = G.upconv_ring_dgrad_index_map(cout, cin, cin, max(64, cout), side)
        wd = torch.where(idx >= 0, flat_w[idx.clamp(min=0).long()], torch.zeros(())).view(cin, -1)
        band = torch.full((n * 3 * (L + 2) * cin,), float("nan"))
        E.tapgemm(plan, sf, wd, band)
        if side < 2:
            r0 = 0 if side == 0 else h - 1
            dxr[:, r0:r0 + 3] += band.view(n, 3, w + 2, cin)
        else:
            c0 = 0 if side == 2 else w - 1
            dxr[:, :, c0:c0 + 3] += band.view(n, h + 2, 3, cin)
        t = 2 if side in (0, 2) else 3
        for p in (0, 1):
            typ = (4 * t + p) if side < 2 else (4 * p + t)
            wp_ = G.plan_upconv_wgrad_ring(n, h, w, cin, cout, side, p)
            E.wgrad(wp_, sf[p * cout:], xr_flat, dw[typ * 9 * cin:])
    assert torch.allclose(dxr, nhwc(xr.grad), atol=2e-3, rtol=1e-4), float((dxr - nhwc(xr.grad)).abs().max())
    assert torch.allclose(dw.view(cout, 16, 3, 3, cin), wph.grad, atol=1e-2, rtol=1e-3)
