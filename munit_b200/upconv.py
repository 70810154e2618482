"""EXPERIMENTAL (MUNIT_UPCONV_PHASE, off by default): nn.Upsample(2) -> ReflectionPad2d(2) -> Conv2d(5x5)
(networks.py:534-545) in phase form -- 9 MACs per output instead of 25 (geometry.plan_upconv_phases and the comment
above it; the algebra is pinned in tests/test_upconv_math_cpu.py).

This module holds the orchestration of the forward and the backward pass around the tap-GEMM / wgrad launches.  It
is written against a small launcher interface so that the SAME code runs
  * on the GPU through kernels.tapgemm / kernels.wgrad (bf16 operands, `GpuLauncher`), and
  * on the CPU through the descriptor emulation of tests/emulate.py (fp32), where tests/test_upconv_cpu.py checks
    the whole backward -- ring strips, bands, corners, replicate-halo fold, phase-gradient fold -- against autograd
    of the direct formulation.
The glue between the launches (cutting the ring strips out of dY, adding the bands, folding the halo, folding the
phase gradients back to the 5x5 parameter) is plain tensor arithmetic here: a correctness scaffold, to be replaced
by fused kernels once the path has been validated and measured on a B200 (it has not run on one yet).
"""
from __future__ import annotations

import functools

import torch

from . import geometry as G


_IDX_DEV = {}  # (id of the cached host map, device) -> device copy


@functools.lru_cache(maxsize=None)
def _dgrad_map(co, c, ck):
    return G.upconv_dgrad_index_map(co, c, c, ck)


@functools.lru_cache(maxsize=None)
def _ring_map(co, c, ck, side):
    return G.upconv_ring_dgrad_index_map(co, c, c, ck, side)


class GpuLauncher:
    """kernels.tapgemm / kernels.wgrad on bf16 operands."""

    dtype = torch.bfloat16

    def tapgemm(self, plan, a, b, out, bias=None):
        from . import kernels as K

        K.tapgemm(plan, a, b, out, bias, "none", ksplit=1)

    def wgrad(self, plan, dy, x, dw):
        from . import kernels as K

        K.wgrad(plan, dy, x, dw)

    def gather(self, src_flat, idx, rows):
        from . import kernels as K

        key = (id(idx), str(src_flat.device))
        dev_idx = _IDX_DEV.get(key)
        if dev_idx is None:  # the host maps are lru-cached objects: one upload per (shape, device)
            dev_idx = _IDX_DEV[key] = idx.to(src_flat.device)
        dst = torch.empty(rows, idx.numel() // rows, dtype=torch.bfloat16, device=src_flat.device)
        return K.gather_cast(src_flat.contiguous(), dev_idx, dst)


def _tap_matrix(dtype, device):
    return torch.tensor([G.UP_ROW_TYPES[t] for t in range(4)], dtype=dtype, device=device)  # [4, 3, 5]


def forward(L, x_lo, wph_mat, bias, co_rows):
    """x_lo [n, h+2, w+2, c] with a REPLICATE halo of 1; wph_mat [co_rows, 16*9*c] -> raw conv output
    [n, 2h, 2w, co_rows]."""
    n, hp, wp, c = x_lo.shape
    h, w = hp - 2, wp - 2
    out = torch.empty(n, 2 * h, 2 * w, co_rows, dtype=x_lo.dtype, device=x_lo.device)
    geom = (4 * h * w * co_rows, 2 * w * co_rows, co_rows, 0, 0)
    for p in G.plan_upconv_phases(n, h, w, c, co_rows, geom):
        L.tapgemm(p, x_lo, wph_mat, out, bias)
    return out


def backward(L, gy, x_lo, wph32, need_dx=True, need_dw=True):
    """gy [n, 2h, 2w, co] gradient of the raw conv output (MODIFIED in place: its outermost ring is zeroed);
    x_lo [n, h+2, w+2, c] the replicate-padded low-res input; wph32 [co, 16, 3, 3, c] fp32 tap sums.
    Returns (gx [n, h+2, w+2, c] -- gradient for the producer of x_lo: interior filled, halo zero -- or None,
             dw5 [co, 5, 5, c] fp32 (channels_last order of the 5x5 parameter) or None)."""
    n, h2, w2, co = gy.shape
    h, w = h2 // 2, w2 // 2
    c = x_lo.shape[3]
    dev, dt = gy.device, gy.dtype
    assert co % 64 == 0 and c % 64 == 0, "phase-form backward needs channel counts that are multiples of 64"
    # ---- cut the ring out of dY: four strips (corners zeroed) and the four corner pixels, then zero the ring
    corners = torch.stack([gy[:, 0, 0], gy[:, 0, -1], gy[:, -1, 0], gy[:, -1, -1]], 1).float()   # [n, 4, co]
    strips = [gy[:, 0].clone(), gy[:, -1].clone(), gy[:, :, 0].clone(), gy[:, :, -1].clone()]   # [n, 2L, co]
    for s in strips:
        s[:, 0] = 0
        s[:, -1] = 0
    gy[:, 0] = 0
    gy[:, -1] = 0
    gy[:, :, 0] = 0
    gy[:, :, -1] = 0
    wflat = wph32.reshape(-1)
    corner_geo = ((2, 2, 0, 0), (2, 3, 0, w - 1), (3, 2, h - 1, 0), (3, 3, h - 1, w - 1))  # (rt, ct, row0, col0)
    gx = None
    if need_dx:
        ck = max(64, co)
        dxr = torch.empty(n, h + 2, w + 2, c, dtype=dt, device=dev)
        wd = L.gather(wflat, _dgrad_map(co, c, ck), c)
        L.tapgemm(G.plan_upconv_dgrad_interior(n, h, w, c, co), gy, wd, dxr)
        acc = dxr.float()
        for side, s in enumerate(strips):
            ln = (w if side < 2 else h) + 2
            band = torch.empty((n, 3, ln, c) if side < 2 else (n, ln, 3, c), dtype=dt, device=dev)
            wr = L.gather(wflat, _ring_map(co, c, ck, side), c)
            L.tapgemm(G.plan_upconv_dgrad_ring(n, h, w, c, co, side), s.contiguous(), wr, band)
            if side == 0:
                acc[:, 0:3] += band.float()
            elif side == 1:
                acc[:, h - 1:h + 2] += band.float()
            elif side == 2:
                acc[:, :, 0:3] += band.float()
            else:
                acc[:, :, w - 1:w + 2] += band.float()
        for k, (rt, ct, r0, c0) in enumerate(corner_geo):   # 4 pixels per image: elementwise, no GEMM
            wk = wph32[:, 4 * rt + ct]                                                   # [co, 3, 3, c]
            acc[:, r0:r0 + 3, c0:c0 + 3] += (corners[:, k, :, None, None, None] * wk[None]).sum(1)
        # fold the replicate halo onto the edge pixels (clamp adjoint); the producer gets a zero halo
        core = acc[:, 1:-1, 1:-1].clone()
        core[:, 0] += acc[:, 0, 1:-1]
        core[:, -1] += acc[:, -1, 1:-1]
        core[:, :, 0] += acc[:, 1:-1, 0]
        core[:, :, -1] += acc[:, 1:-1, -1]
        core[:, 0, 0] += acc[:, 0, 0]
        core[:, 0, -1] += acc[:, 0, -1]
        core[:, -1, 0] += acc[:, -1, 0]
        core[:, -1, -1] += acc[:, -1, -1]
        gx = torch.zeros(n, h + 2, w + 2, c, dtype=dt, device=dev)
        gx[:, 1:-1, 1:-1] = core.to(dt)
    dw5 = None
    if need_dw:
        scratch = torch.zeros(co * 16 * 9 * c, dtype=torch.float32, device=dev)
        gflat, xflat = gy.reshape(-1), x_lo.reshape(-1)
        for py in (0, 1):
            for px in (0, 1):
                L.wgrad(G.plan_upconv_wgrad_interior(n, h, w, c, co, py, px), gflat[(py * 2 * w + px) * co:], xflat,
                        scratch[(4 * py + px) * 9 * c:])
        for side, s in enumerate(strips):
            sf = s.contiguous().reshape(-1)
            t = 2 if side in (0, 2) else 3
            for p in (0, 1):
                typ = (4 * t + p) if side < 2 else (4 * p + t)
                L.wgrad(G.plan_upconv_wgrad_ring(n, h, w, c, co, side, p), sf[p * co:], xflat, scratch[typ * 9 * c:])
        sc = scratch.view(co, 16, 3, 3, c)
        xf = x_lo.float()
        for k, (rt, ct, r0, c0) in enumerate(corner_geo):
            patch = xf[:, r0:r0 + 3, c0:c0 + 3]                                         # [n, 3, 3, c]
            sc[:, 4 * rt + ct] += (corners[:, k, :, None, None, None] * patch[:, None]).sum(0)
        a = _tap_matrix(torch.float32, dev)
        # dW5[o, ky, kx, i] = sum_{r, c, dy, dx} A_r[dy][ky] * A_c[dx][kx] * dWph[o, r, c, dy, dx, i]
        dw5 = torch.einsum("rak,cbl,orcabi->okli", a, a, sc.view(co, 4, 4, 3, 3, c)).contiguous()
    return gx, dw5
