"""Micro-benchmark of the bandwidth kernels (norm stats/apply/backward, act_bwd, adam) at config_256 shapes, B=8.
Reports achieved GB/s on the ALGORITHMIC bytes against MEASURED_PEAKS hbm_gbs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import kernels as K

SHAPES = {"in_256x64x64": (8, 64, 64, 256, "in", 1, 1), "in_64x256x256": (8, 256, 256, 64, "in", 1, 3),
          "ln_128x128x128_up": (8, 128, 128, 128, "ln", 2, 2), "adain_res_256x64x64": (8, 64, 64, 256, "adain", 1, 1)}

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

out = {}
for name, (n, h, w, c, mode, up, pad) in SHAPES.items():
    nb = 6  # rotate buffers: 6 x (>=16 MB) > L2 for the big ones
    ys = [torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16) for _ in range(nb)]
    res = torch.randn(n, h + 2, w + 2, c, device="cuda").to(torch.bfloat16) if "res" in name else None
    pw = torch.randn(n, c, device="cuda") if mode == "adain" else (torch.rand(c, device="cuda") if mode == "ln" else None)
    pb = torch.randn(n, c, device="cuda") if mode == "adain" else (torch.zeros(c, device="cuda") if mode == "ln" else None)
    ldw = c if mode == "adain" else 0
    el = n * h * w * c
    i = [0]
    def nxt():
        i[0] += 1
        return ys[i[0] % nb]
    stats, shift = K.norm_stats(ys[0])
    coef = K.norm_finalize(stats, shift, mode, pw, pb, ldw, h * w)
    outb = K.norm_apply(ys[0], coef[2], coef[3], True, res, 1, pad, up)
    gouts = [torch.randn_like(outb) for _ in range(2)]
    r = {}
    r["stats"] = (timeit(lambda: K.norm_stats(nxt())), 2 * el)
    r["finalize"] = (timeit(lambda: K.norm_finalize(stats, shift, mode, pw, pb, ldw, h * w)), 0)
    r["apply"] = (timeit(lambda: K.norm_apply(nxt(), coef[2], coef[3], True, res, 1, pad, up)), (2 + 2 * up * up + (2 if res is not None else 0)) * el)
    gw = torch.zeros(n, c, device="cuda") if mode == "adain" else (torch.zeros(c, device="cuda") if mode == "ln" else None)
    gb = torch.zeros_like(gw) if gw is not None else None
    r["bwd(3 kernels)"] = (timeit(lambda: K.norm_bwd(gouts[i[0] % 2], pad, up, nxt(), coef, True, mode, pw, ldw, gw, gb, c if mode == "adain" else 0, res is not None, 1)),
                           (2 * (2 * up * up + 2) + 2 + (2 if res is not None else 0)) * el)
    out[name] = {k: dict(us=round(v[0], 1), gbs=round(v[1] / v[0] / 1e3, 1) if v[1] else None) for k, v in r.items()}
    print(name, out[name], flush=True)
p = torch.randn(27_000_000, device="cuda"); g = torch.randn_like(p); m = torch.zeros_like(p); v = torch.zeros_like(p)
us = timeit(lambda: K.adam(p, g, m, v, None, None, 0, False, 1e-4, 0.5, 0.999, 1e-8, 1e-4, 3))
print("adam 27M", round(us, 1), "us", round(28 * p.numel() / us / 1e3, 1), "GB/s")
out["adam_27M"] = dict(us=round(us, 1), gbs=round(28 * p.numel() / us / 1e3, 1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_simt.json"), "w"), indent=1)
