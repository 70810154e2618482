"""munit_b200/upconv.py (EXPERIMENTAL phase form of upsample + 5x5 conv): the whole forward / backward orchestration
-- launch plans, ring strips, bands, corner pixels, replicate-halo fold, phase-gradient fold -- run on the CPU through
the descriptor emulation (tests/emulate.py) in fp32 and compared with autograd of the direct formulation
nn.Upsample(2) -> ReflectionPad2d(2) -> Conv2d(5x5) (networks.py:534-545)."""
import pytest
import torch
import torch.nn.functional as F

from munit_b200 import geometry as G
from munit_b200 import upconv
from tests import emulate as E


class EmuLauncher:
    dtype = torch.float32

    def tapgemm(self, plan, a, b, out, bias=None):
        assert out.is_contiguous()
        E.tapgemm(plan, a.reshape(-1), b, out.view(-1), bias, "none")

    def wgrad(self, plan, dy, x, dw):
        E.wgrad(plan, dy.reshape(-1), x.reshape(-1), dw)

    def gather(self, src_flat, idx, rows):
        return torch.where(idx >= 0, src_flat[idx.clamp(min=0).long()], torch.zeros(())).view(rows, -1)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n,h,w,ci,co", [(2, 4, 6, 64, 64), (1, 3, 5, 128, 64), (1, 2, 2, 64, 128)])
def test_phase_form_forward_backward_through_the_emulated_launches(n, h, w, ci, co):
    torch.manual_seed(0)
    x = torch.randn(n, ci, h, w, requires_grad=True)
    wt = (torch.randn(co, ci, 5, 5) * 0.1).requires_grad_(True)
    bias = torch.randn(co)
    y_ref = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (2,) * 4, mode="reflect"), wt, bias)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    L = EmuLauncher()
    x_lo = nhwc(F.pad(x.detach(), (1,) * 4, mode="replicate"))
    wph = G.upconv_phase_weights(wt.detach())
    y = upconv.forward(L, x_lo, wph.reshape(co, -1), bias, co)
    assert torch.allclose(y, nhwc(y_ref.detach()), atol=2e-3, rtol=1e-4)
    g = nhwc(gy).clone()
    gx, dw5 = upconv.backward(L, g, x_lo, wph)
    assert float(g[:, 0].abs().max()) == 0 and float(g[:, :, -1].abs().max()) == 0    # ring consumed
    assert float(gx[:, 0].abs().max()) == 0 and float(gx[:, :, 0].abs().max()) == 0  # zero halo for the producer
    err_x = float((gx[:, 1:-1, 1:-1] - nhwc(x.grad)).abs().max())
    assert torch.allclose(gx[:, 1:-1, 1:-1], nhwc(x.grad), atol=5e-3, rtol=1e-3), err_x
    ref_w = wt.grad.permute(0, 2, 3, 1)                                               # [co, ky, kx, ci]
    assert torch.allclose(dw5, ref_w, atol=2e-2, rtol=1e-3), float((dw5 - ref_w).abs().max())
    # only one of the two gradients requested
    gx2, dw2 = upconv.backward(L, nhwc(gy).clone(), x_lo, wph, need_dx=False)
    assert gx2 is None and torch.allclose(dw2, dw5, atol=1e-5)
