"""First GPU contact for the experimental phase form of upsample + 5x5 conv (DESIGN.md s7 "next"; run on a B200):

    MUNIT_UPCONV_PHASE=2 python tools/upconv_check.py

For the two decoder layers at batch 8 (and 16, the batched gen_update calls) it checks forward / input gradient /
weight gradient of ops.UpConvPhaseFn against torch's direct formulation and times forward and backward against the
current path (norm-apply-written upsampled buffer + halo-resident 5x5 tap-GEMM, ops.ConvFn)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import kernels as K  # noqa: E402
from munit_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


def main():
    assert int(ops.UPCONV_PHASE) >= 2, "run with MUNIT_UPCONV_PHASE=2"
    for n, h, cin, cout in ((8, 64, 256, 128), (8, 128, 128, 64), (16, 64, 256, 128), (16, 128, 128, 64)):
        torch.manual_seed(0)
        x = torch.randn(n, cin, h, h, device="cuda").to(torch.bfloat16)
        wt = (torch.randn(cout, cin, 5, 5, device="cuda") * (2.0 / (25 * cin)) ** 0.5).to(torch.bfloat16).float()
        bias = torch.randn(cout, device="cuda") * 0.1
        # ---- reference (fp32, cuDNN) -- a checker only
        xr, wr = x.float().requires_grad_(True), wt.clone().requires_grad_(True)
        ref = F.conv2d(F.pad(F.interpolate(xr, scale_factor=2, mode="nearest"), (2,) * 4, mode="reflect"), wr, bias)
        gy = torch.randn_like(ref).to(torch.bfloat16)
        ref.backward(gy.float())
        # ---- phase form
        layer = ops.ConvLayer(cin, cout, 5, 1, 2)
        wg = wt.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        x_lo = F.pad(x.float(), (1,) * 4, mode="reflect").permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        x_lo.requires_grad_(True)
        y = ops.UpConvPhaseFn.apply(x_lo, wg, bias, layer)
        g_nhwc = gy.permute(0, 2, 3, 1).contiguous()
        y.backward(g_nhwc.clone())
        print("n=%d %dx%d %d->%d  fwd %.2e  dx %.2e  dw %.2e (rel-L2 vs fp32)" % (
            n, h, h, cin, cout, rel(y.permute(0, 3, 1, 2), ref), rel(x_lo.grad[:, 1:-1, 1:-1].permute(0, 3, 1, 2), xr.grad),
            rel(wg.grad, wr.grad)))
        # ---- timing: phase form vs current path (the upsampled, reflect-padded buffer is what norm_apply writes today)
        up = F.pad(F.interpolate(x.float(), scale_factor=2, mode="nearest"), (2,) * 4, mode="reflect")
        up = up.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).requires_grad_(True)
        layer2 = ops.ConvLayer(cin, cout, 5, 1, 2)
        wg2 = wt.contiguous(memory_format=torch.channels_last).requires_grad_(True)

        def cur_fwd():
            with torch.no_grad():
                return ops.ConvFn.apply(up, wg2, bias, layer2, "none", 0, 2)

        def ph_fwd():
            with torch.no_grad():
                return ops.upconv_phase_forward(x_lo.detach(), wg, bias, layer)

        def cur_fb():
            out = ops.ConvFn.apply(up, wg2, bias, layer2, "none", 0, 2)
            out.backward(g_nhwc)

        def ph_fb():
            out = ops.UpConvPhaseFn.apply(x_lo, wg, bias, layer)
            out.backward(g_nhwc.clone())

        before = K._lib.launches if hasattr(K, "_lib") else 0
        print("    forward  current %.1f us   phase form %.1f us" % (timeit(cur_fwd), timeit(ph_fwd)))
        print("    fwd+bwd  current %.1f us   phase form %.1f us (glue still tensor arithmetic)" % (
            timeit(cur_fb, 10), timeit(ph_fb, 10)))
        del before


if __name__ == "__main__":
    main()
