"""Times the one-launch weight-shadow refresh (ops.refresh_shadows -> munit_gather_cast_multi) of both optimiser arenas."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from munit_b200 import ops  # noqa: E402
from munit_b200.engine import StepRunner  # noqa: E402
from munit_b200.trainer import MUNIT_Trainer  # noqa: E402

cfg = bench.load_cfg()
t = MUNIT_Trainer(cfg).cuda()
r = StepRunner(t, cfg, 2, 64, use_graph=False)
xa, xb = bench.synthetic_images(2, 64, 1)
st = [torch.randn(2, t.style_dim, 1, 1) for _ in range(4)]
r.load_inputs(xa, xb, *st)
r.step()
r.step()
torch.cuda.synchronize()
for name, opt in (("dis", t.dis_opt), ("gen", t.gen_opt)):
    ps = opt._all_params()
    ops.refresh_shadows(opt, ps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.refresh_shadows(opt, ps)
    e1.record()
    torch.cuda.synchronize()
    _, layers, table, segs = opt._shadow_cache
    elems = sum(d.numel() for _, _, d in segs)
    ident = sum(d.numel() for _, i, d in segs if i is None)
    print("%s: %d layers, %d segments, %.1f M shadow elements (%.1f M identity), %.1f us per refresh" % (
        name, len(layers), len(segs), elems / 1e6, ident / 1e6, e0.elapsed_time(e1) * 100))
