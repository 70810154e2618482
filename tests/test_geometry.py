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
