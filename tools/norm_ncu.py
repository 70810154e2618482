"""Profiling aid: launches each normalisation kernel (statistics, apply, backward reduce, backward apply) once inside
the profiler range on two generator shapes -- the res-block IN (B=8, 64x64x256, fp16 pre-norm tensor) and the decoder
LayerNorm with the 2x up-sampling producer (B=8, 128x128x128 -> 256x256) -- after a warm-up pass, for

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/norm python tools/norm_ncu.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import kernels as K  # noqa: E402


def run(mode, n, h, w, c, out_pad, up, relu):
    g = torch.Generator(device="cuda").manual_seed(0)
    y = (torch.randn(n, h, w, c, device="cuda", generator=g) * 1.3 + 0.2).to(torch.float16)
    p_w = p_b = None
    if mode == "ln":
        p_w, p_b = torch.rand(c, device="cuda"), torch.randn(c, device="cuda")
    out, coef = K.norm_fwd(y, mode, p_w, p_b, 0, 1e-5, relu, None, 1, out_pad, up)
    g_out = torch.randn(out.shape, device="cuda", generator=g).to(torch.bfloat16)
    g_w = g_b = None
    if mode == "ln":
        g_w, g_b = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    K.norm_bwd(g_out, out_pad, up, y, coef, relu, mode, p_w, 0, g_w, g_b, 0, False, 1)
    torch.cuda.synchronize()


if __name__ == "__main__":
    shapes = [("in", 8, 64, 64, 256, 1, 1, True), ("ln", 8, 128, 128, 128, 2, 2, True)]
    for s in shapes:
        run(*s)
    torch.cuda.profiler.start()
    for s in shapes:
        run(*s)
    torch.cuda.profiler.stop()
    print("ok")
