"""Profiling aid: a few launches of the res-block 3x3 256->256 convolution (B = 8, 64x64; forward and dgrad) inside the
profiler range, rotating over buffers larger than L2 as in the step, for

    [MUNIT_PAIR=1] ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/conv python tools/conv_ncu.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import geometry as G, kernels as K  # noqa: E402

n, h, w, c = 8, 64, 64, 256
hp, wp = h + 2, w + 2
fwd = G.plan_fwd(n, hp, wp, c, 3, 3, 1, 1, c, (h * w * c, w * c, c, 0, 0))
dg = G.plan_dgrad(n, hp, wp, c, 3, 3, 1, 1, c)
xs = [torch.randn(n, hp, wp, c, device="cuda").to(torch.bfloat16) for _ in range(6)]
wf = (torch.randn(c, 9 * c, device="cuda") * 0.02).to(torch.bfloat16)
wd = (torch.randn(c, dg.b_k, device="cuda") * 0.02).to(torch.bfloat16)
ys = [torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda") for _ in range(6)]
dxs = [torch.empty(n, hp, wp, c, dtype=torch.bfloat16, device="cuda") for _ in range(2)]
for i in range(6):
    K.tapgemm(fwd, xs[i], wf, ys[i])
K.tapgemm(dg, ys[0], wd, dxs[0])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(2):
    K.tapgemm(fwd, xs[i], wf, ys[i])
K.tapgemm(dg, ys[1], wd, dxs[1])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
