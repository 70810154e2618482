"""In-graph micro-benchmark of the normalisation kernel sequences (forward: statistics -> finalize -> apply;
backward: reduce -> finalize -> apply) on the shapes of the MUNIT generator.  Each sequence is captured R times
into one CUDA graph and the graph replay is timed, so the numbers carry graph-replay launch behaviour rather than
Python / ctypes launch overhead.  Usage: python tools/bench_norm.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from munit_b200 import kernels as K  # noqa: E402

R = 20


def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / (iters * R)  # us per call of fn


def main():
    res = {}
    shapes = [("in_res", "in", 8, 64, 64, 256, 1, 1, True, False), ("adain_res", "adain", 8, 64, 64, 256, 1, 1, False, True),
              ("in_res_n16", "in", 16, 64, 64, 256, 1, 1, True, False), ("in_conv2", "in", 8, 128, 128, 128, 1, 1, True, False),
              ("in_conv1", "in", 8, 256, 256, 64, 1, 1, True, False), ("ln_up1", "ln", 8, 128, 128, 128, 2, 1, True, False),
              ("ln_up2", "ln", 8, 256, 256, 64, 3, 1, True, False), ("adain_up", "adain", 8, 64, 64, 256, 2, 2, False, True)]
    for name, mode, n, h, w, c, out_pad, up, relu, res_on in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        y = (torch.randn(n, h, w, c, device="cuda", generator=g) * 1.3 + 0.2).to(torch.float16)
        resid = torch.randn(n, h + 2, w + 2, c, device="cuda", generator=g).to(torch.bfloat16) if res_on else None
        p_w = p_b = None
        ldw = 0
        if mode == "adain":
            p_w, p_b, ldw = torch.randn(n, c, device="cuda"), torch.randn(n, c, device="cuda"), c
        elif mode == "ln":
            p_w, p_b = torch.rand(c, device="cuda"), torch.randn(c, device="cuda")
        out, coef = K.norm_fwd(y, mode, p_w, p_b, ldw, 1e-5, relu, resid, 1, out_pad, up)
        g_out = torch.randn(out.shape, device="cuda", generator=g).to(torch.bfloat16)
        g_w = g_b = None
        ldg = 0
        if mode == "adain":
            g_w, g_b, ldg = torch.empty(n, c, device="cuda"), torch.empty(n, c, device="cuda"), c
        elif mode == "ln":
            g_w, g_b = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        el = y.numel() * 2
        gel = g_out.numel() * 2
        r = {}
        stats, shift = K.norm_stats(y)
        r["stats"] = (graph_time(lambda: K.norm_stats(y)), el)
        r["finalize"] = (graph_time(lambda: K.norm_finalize(stats, shift, mode, p_w, p_b, ldw, h * w)), 0)
        r["apply"] = (graph_time(lambda: K.norm_apply(y, coef[2], coef[3], relu, resid, 1, out_pad, up)),
                      el + gel + (el if res_on else 0))
        r["fwd(3)"] = (graph_time(lambda: K.norm_fwd(y, mode, p_w, p_b, ldw, 1e-5, relu, resid, 1, out_pad, up)),
                       2 * el + gel + (el if res_on else 0))
        sums = K.norm_bwd_reduce(g_out, out_pad, up, y, coef, relu)
        kk = K.norm_bwd_finalize(sums, coef, mode, p_w, ldw, g_w, g_b, ldg, h * w)
        r["b_reduce"] = (graph_time(lambda: K.norm_bwd_reduce(g_out, out_pad, up, y, coef, relu)), el + gel)
        r["b_final"] = (graph_time(lambda: K.norm_bwd_finalize(sums, coef, mode, p_w, ldw, g_w, g_b, ldg, h * w)), 0)
        r["b_apply"] = (graph_time(lambda: K.norm_bwd_apply(g_out, out_pad, up, y, coef, relu, kk, res_on, 1)),
                        2 * el + gel + (el if res_on else 0))
        r["bwd(3)"] = (graph_time(lambda: K.norm_bwd(g_out, out_pad, up, y, coef, relu, mode, p_w, ldw, g_w, g_b, ldg,
                                                     res_on, 1)), 2 * (el + gel) + el + (el if res_on else 0))
        res[name] = {k: {"us": round(v[0], 2), "gbs": round(v[1] / v[0] / 1e3, 1) if v[1] else None} for k, v in r.items()}
        print(name, " ".join("%s=%.1f" % (k, v["us"]) for k, v in res[name].items()), flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
