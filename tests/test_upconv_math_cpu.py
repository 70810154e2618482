"""The algebra behind the phase form of nn.Upsample(2) -> ReflectionPad2d(2) -> Conv2d(5x5) (networks.py:534-545),
forward AND backward, checked against autograd of the direct formulation on the CPU.  The forward launch plans
(geometry.plan_upconv_phases) implement the first part; this file is the verified blueprint for the backward half
(DESIGN.md s7 "next"):

  * every output pixel has a (row type, column type) in {even, odd, first, last}^2 and reads the 3x3 low-res
    neighbourhood of its low-res pixel on the replicate-padded input with that type's tap sums;
  * dW5 = sum over types of A_r^T dW_type A_c  (A = geometry.UP_ROW_TYPES, the 3x5 tap-collection matrices);
  * d(replicate-padded input) is the scatter of dY through the type weights, and the replicate halo folds onto the
    edge pixels (clamp adjoint).
"""
import pytest
import torch
import torch.nn.functional as F

from munit_b200 import geometry as G


def _types(h):
    """Output rows of each row type and their low-res row i: even 2i (i >= 1), odd 2i+1 (i <= h-2), first, last."""
    even = [(2 * i, i) for i in range(1, h)]
    odd = [(2 * i + 1, i) for i in range(0, h - 1)]
    return {0: even, 1: odd, 2: [(0, 0)], 3: [(2 * h - 1, h - 1)]}


def _phase_forward(xr, wph, h, w):
    n, c = xr.shape[:2]
    co = wph.shape[0]
    y = torch.zeros(n, co, 2 * h, 2 * w, dtype=xr.dtype)
    rows, cols = _types(h), _types(w)
    for rt, rl in rows.items():
        for ct, cl in cols.items():
            if not rl or not cl:
                continue
            oy = torch.tensor([o for o, _ in rl]); iy = torch.tensor([i for _, i in rl])
            ox = torch.tensor([o for o, _ in cl]); ix = torch.tensor([i for _, i in cl])
            acc = torch.zeros(n, co, len(rl), len(cl), dtype=xr.dtype)
            for dy in range(3):
                for dx in range(3):
                    patch = xr[:, :, iy + dy][:, :, :, ix + dx]                     # [n, ci, R, C]
                    acc += torch.einsum("oi,nirc->norc", wph[:, 4 * rt + ct, dy, dx], patch)
            y[:, :, oy[:, None], ox[None, :]] = acc
    return y


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 3, 4, 5, 4), (1, 2, 2, 3, 2), (1, 6, 5, 4, 3)])
def test_phase_form_forward_and_backward_match_autograd(n, h, w, ci, co):
    torch.manual_seed(0)
    dt = torch.float64
    x = torch.randn(n, ci, h, w, dtype=dt, requires_grad=True)
    wt = torch.randn(co, ci, 5, 5, dtype=dt, requires_grad=True)
    y_ref = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (2,) * 4, mode="reflect"), wt)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)

    # ---- forward through the 16 type weight sets on the replicate-padded low-res input
    xr = F.pad(x.detach(), (1,) * 4, mode="replicate")
    wph = G.upconv_phase_weights(wt.detach())                       # [co, 16, 3, 3, ci]
    y = _phase_forward(xr, wph, h, w)
    assert torch.allclose(y, y_ref.detach(), atol=1e-10)

    # ---- backward, explicit
    a = torch.tensor([G.UP_ROW_TYPES[t] for t in range(4)], dtype=dt)          # [4, 3, 5]
    rows, cols = _types(h), _types(w)
    dwph = torch.zeros_like(wph)
    dxr = torch.zeros_like(xr)
    for rt, rl in rows.items():
        for ct, cl in cols.items():
            if not rl or not cl:
                continue
            oy = torch.tensor([o for o, _ in rl]); iy = torch.tensor([i for _, i in rl])
            ox = torch.tensor([o for o, _ in cl]); ix = torch.tensor([i for _, i in cl])
            g = gy[:, :, oy[:, None], ox[None, :]]                               # [n, co, R, C]
            for dy in range(3):
                for dx in range(3):
                    patch = xr[:, :, iy + dy][:, :, :, ix + dx]
                    dwph[:, 4 * rt + ct, dy, dx] += torch.einsum("norc,nirc->oi", g, patch)
                    contrib = torch.einsum("oi,norc->nirc", wph[:, 4 * rt + ct, dy, dx], g)
                    # (rows / columns of one type are distinct: no duplicate indices inside one update)
                    dxr[:, :, (iy + dy)[:, None], (ix + dx)[None, :]] += contrib
    # weight gradient: dW5[ky][kx] = sum_types sum_{dy,dx} A_r[dy][ky] * A_c[dx][kx] * dW_type[dy][dx]
    d5 = torch.einsum("rak,cbl,orcabi->oikl", a, a, dwph.view(co, 4, 4, 3, 3, ci))
    assert torch.allclose(d5, wt.grad, atol=1e-9)
    # input gradient: fold the replicate halo (clamp adjoint) onto the edge pixels
    dx_ = dxr[:, :, 1:-1, 1:-1].clone()
    dx_[:, :, 0] += dxr[:, :, 0, 1:-1]
    dx_[:, :, -1] += dxr[:, :, -1, 1:-1]
    dx_[:, :, :, 0] += dxr[:, :, 1:-1, 0]
    dx_[:, :, :, -1] += dxr[:, :, 1:-1, -1]
    dx_[:, :, 0, 0] += dxr[:, :, 0, 0]
    dx_[:, :, 0, -1] += dxr[:, :, 0, -1]
    dx_[:, :, -1, 0] += dxr[:, :, -1, 0]
    dx_[:, :, -1, -1] += dxr[:, :, -1, -1]
    assert torch.allclose(dx_, x.grad, atol=1e-9)


def test_interior_types_cover_everything_but_the_ring():
    """A launch over ALL low-res pixels with the interior types (what plans[0] does) differs from the exact result
    only on the outermost output row / column on each side -- so zeroing that ring in dY makes the interior-type
    backward launches exact for the interior pixels' share of the gradients."""
    torch.manual_seed(1)
    n, h, w, ci, co = 1, 4, 5, 3, 2
    x = torch.randn(n, ci, h, w, dtype=torch.float64)
    wt = torch.randn(co, ci, 5, 5, dtype=torch.float64)
    y_ref = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (2,) * 4, mode="reflect"), wt)
    xr = F.pad(x, (1,) * 4, mode="replicate")
    wph = G.upconv_phase_weights(wt)
    y = torch.zeros_like(y_ref)
    for py in (0, 1):
        for px in (0, 1):
            k = wph[:, 4 * py + px].permute(0, 3, 1, 2)                         # [co, ci, 3, 3]
            y[:, :, py::2, px::2] = F.conv2d(xr, k)
    diff = (y - y_ref).abs()
    assert float(diff[:, :, 1:-1, 1:-1].max()) < 1e-10
    assert float(diff[:, :, 0].max()) > 1e-3 and float(diff[:, :, :, -1].max()) > 1e-3
