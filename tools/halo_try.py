import os, sys, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
from munit_b200 import geometry as G, kernels as K
from tests.gpu_util import bf16_round, nhwc, rel_l2, error_flag
for mode in (1, 3):  # 1: base_offset 0 (correct), 3: (addr>>7)&7 (rejected)
    for (n, h, w, cin, cout, k, pad) in [(2, 16, 16, 64, 64, 3, 1), (1, 24, 40, 128, 128, 5, 2)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = bf16_round(torch.randn(n, cin, h, w, device="cuda", generator=g))
        wt = bf16_round(torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.05)
        xp = F.pad(x, (pad,) * 4, mode="reflect")
        y = F.conv2d(xp, wt)
        plan = G.plan_fwd(n, h + 2 * pad, w + 2 * pad, cin, k, k, 1, 1, cout, (h * w * cout, w * cout, cout, 0, 0), halo=mode)
        out = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
        K.tapgemm(plan, nhwc(xp).to(torch.bfloat16), wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(torch.bfloat16), out)
        torch.cuda.synchronize()
        print("HALO mode", mode, (n, h, w, cin, cout, k), "rel_l2 %.5f" % rel_l2(out, nhwc(y)), "flag", error_flag(), flush=True)
