"""Micro-benchmark of the tap-GEMM kernels on the config_256 layer shapes (B=8): CUDA-event timing, L2 flushed
between iterations by rotating over distinct buffers.  Usage: python tools/bench_conv.py [case ...]"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from munit_b200 import geometry as G, kernels as K

CASES = {
    # name: (n, h, w, cin, cout, k, s, pad)
    "res3x3": (8, 64, 64, 256, 256, 3, 1, 1),
    "dec2_5x5": (8, 128, 128, 256, 128, 5, 1, 2),
    "dec4_5x5": (8, 256, 256, 128, 64, 5, 1, 2),
    "down1_4x4s2": (8, 256, 256, 64, 128, 4, 2, 1),
    "down2_4x4s2": (8, 128, 128, 128, 256, 4, 2, 1),
    "dec5_7x7": (8, 256, 256, 64, 16, 7, 1, 3),
    "dis3_4x4s2": (8, 32, 32, 256, 512, 4, 2, 1),
    # kw-expanded first layer (7x7 RGB as a 7x1 GEMM over E) and the narrow-output layer (R space): (kh, kw) kernels
    "first_e_7x1": (8, 256, 256, 64, 64, (7, 1), 1, 3),
    "last_r_7x1": (8, 256, 256, 64, 32, (7, 1), 1, 3),
}

def run_rect(name, iters=20, nbuf=4):
    n, h, w, cin, cout, (kh, kw), s, pad = CASES[name]
    hp, wp = h + 2 * pad, w  # vertical halo only
    ho, wo = G.conv_out(hp, kh, s), G.conv_out(wp, kw, 1)
    flops = 2.0 * n * ho * wo * cout * kh * kw * cin
    xs = [torch.randn(n, hp, wp, cin, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    wf = (torch.randn(cout, kh * kw * cin, device="cuda") * 0.05).to(torch.bfloat16)
    ys = [torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
    halo = int(os.environ.get("HALO", 0))
    fwd = G.plan_fwd(n, hp, wp, cin, kh, kw, s, 1, cout, (ho * wo * cout, wo * cout, cout, 0, 0), halo=halo)
    dg = G.plan_dgrad(n, hp, wp, cin, kh, kw, s, 1, cout, halo=halo if cout >= 64 else 0)
    print("halo fwd/dgrad:", fwd.halo, dg.halo, "tile", (fwd.tw, fwd.th, fwd.tn))
    wd = (torch.randn(cin, dg.b_k, device="cuda") * 0.05).to(torch.bfloat16)
    dxs = [torch.empty(n, hp, wp, cin, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
    wg = G.plan_wgrad(n, hp, wp, cin, kh, kw, s, 1, cout, cout, kh * kw * cin, cin, 1)
    dw = torch.zeros(cout, kh, kw, cin, device="cuda")
    stages = int(os.environ.get("STAGES", 0))
    res = {}
    for label, fn in (("fwd", lambda i: K.tapgemm(fwd, xs[i % nbuf], wf, ys[i % nbuf], stages=stages)),
                      ("dgrad", lambda i: K.tapgemm(dg, ys[i % nbuf], wd, dxs[i % nbuf], stages=stages)),
                      ("wgrad", lambda i: K.wgrad(wg, ys[i % nbuf], xs[i % nbuf], dw))):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[label] = dict(us=ms * 1000, tflops=flops / ms / 1e9)
    gb = (xs[0].numel() + ys[0].numel()) * 2 / 1e9
    print(name, {k_: (round(v["us"], 1), round(v["tflops"], 1)) for k_, v in res.items()},
          "fwd GB/s %.0f" % (gb / (res["fwd"]["us"] * 1e-6)), flush=True)
    return res


def run(name, iters=20, nbuf=4):
    n, h, w, cin, cout, k, s, pad = CASES[name]
    if isinstance(k, tuple):
        return run_rect(name, iters, nbuf)
    hp, wp = h + 2 * pad, w + 2 * pad
    ho, wo = G.conv_out(hp, k, s), G.conv_out(wp, k, s)
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    xs = [torch.randn(n, hp, wp, cin, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    wf = (torch.randn(cout, k * k * cin, device="cuda") * 0.05).to(torch.bfloat16)
    ys = [torch.empty(n, ho, wo, cout, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
    halo = int(os.environ.get("HALO", 0))
    fwd = G.plan_fwd(n, hp, wp, cin, k, k, s, s, cout, (ho * wo * cout, wo * cout, cout, 0, 0), halo=halo)
    if os.environ.get("BN"):
        fwd.bn = int(os.environ["BN"])
    stages = int(os.environ.get("STAGES", 0))
    dg = G.plan_dgrad(n, hp, wp, cin, k, k, s, s, cout, halo=halo)
    ck = max(64, cout)
    wd = (torch.randn(cin, dg.b_k, device="cuda") * 0.05).to(torch.bfloat16)
    dxs = [torch.empty(n, hp, wp, cin, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
    wg = G.plan_wgrad(n, hp, wp, cin, k, k, s, s, cout, cout, k * k * cin, cin, 1)
    if int(os.environ.get("WBN", 0)):
        wg.bn = int(os.environ["WBN"])
    if os.environ.get("NOSWAP"):
        wg = G.plan_wgrad(n, hp, wp, cin, k, k, s, s, cout, cout, k * k * cin, cin, 1, swap=False)
    ksplit, wstages = int(os.environ.get("KSPLIT", 0)), int(os.environ.get("WSTAGES", 0))
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    res = {}
    for label, fn in (("fwd", lambda i: K.tapgemm(fwd, xs[i % nbuf], wf, ys[i % nbuf], stages=stages)),
                      ("dgrad", lambda i: K.tapgemm(dg, ys[i % nbuf], wd, dxs[i % nbuf])),
                      ("wgrad", lambda i: K.wgrad(wg, ys[i % nbuf], xs[i % nbuf], dw, ksplit=ksplit, stages=wstages))):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[label] = dict(us=ms * 1000, tflops=flops / ms / 1e9)
    print(name, {k_: (round(v["us"], 1), round(v["tflops"], 1)) for k_, v in res.items()}, flush=True)
    return res

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    out = {nm: run(nm) for nm in names}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_conv.json"), "w"), indent=1)
