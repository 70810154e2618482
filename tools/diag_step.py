"""Diagnostic (run on the GPU box): one dis_update+gen_update of munit_b200 vs the oracle in fp32 mode and
in storage-aware (bf16 STE) mode; dumps forward-stage and per-parameter gradient errors."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import munit_oracle as O  # noqa: E402
from tests.gpu_util import rel_l2  # noqa: E402
from tests.test_trainer_gpu import _build, _images  # noqa: E402

hw, b = int(os.environ.get("HW", 64)), int(os.environ.get("B", 2))
cfg = O.config_256_core(gen_state=int(os.environ.get("GS", 1)), guided=int(os.environ.get("GUIDED", 1)))
seeds = dict(gen=21, gen_b=22, dis_a=23, dis_b=24)
out = {}
for mode in ("fp32", "bf16ste"):
    O.QUANT = mode == "bf16ste"
    t, orc = _build(cfg, seeds)
    x_a, x_b = _images(1234, b, hw)
    xa, xb = x_a.cuda(), x_b.cuda()
    # forward stages
    with torch.no_grad():
        ca, sa = t.gen.encode(xa, 1) if cfg["gen_state"] == 1 else t.gen_a.encode(xa)
        oca, osa = orc.enc_a(x_a)
        xab = t._dec("b", t._enc("a", xa)[0], t._enc("b", xb)[1])
        oxab = orc.dec_b(oca, orc.enc_b(x_b)[1])
    fw = dict(c_a=rel_l2(ca.cpu(), oca), s_a=rel_l2(sa.cpu(), osa), x_ab=rel_l2(xab.cpu(), oxab))
    torch.manual_seed(99)
    rng = torch.get_rng_state()
    t.iterations = 0
    t.dis_update(xa, xb, cfg)
    dis_g = {f"a/{n}": p.grad.detach().cpu().clone() for n, p in t.dis_a.named_parameters()}
    t.gen_update(xa, xb, cfg)
    torch.set_rng_state(rng)
    orc.dis_update(x_a, x_b)
    orc.gen_update(x_a, x_b)
    losses = {k: (float(getattr(t, k)), orc.losses[k]) for k in orc.losses}
    gens = {"": t.gen} if cfg["gen_state"] == 1 else {"a": t.gen_a, "b": t.gen_b}
    ge = {}
    for gn, g in gens.items():
        for n, p in g.named_parameters():
            key = f"{gn}/{n}"
            if key in orc.gen_grads:
                ge[key] = (rel_l2(p.grad.cpu(), orc.gen_grads[key]), float(orc.gen_grads[key].norm()))
    de = {k: (rel_l2(v, orc.dis_grads[k]), float(orc.dis_grads[k].norm())) for k, v in dis_g.items()}
    out[mode] = dict(forward=fw, losses=losses, gen_grad=ge, dis_grad=de)
    print("==== mode", mode)
    print("forward rel-L2:", {k: round(v, 5) for k, v in fw.items()})
    print("losses (ours, oracle):", {k: (round(a, 5), round(b_, 5)) for k, (a, b_) in losses.items()})
    vals = sorted(v[0] for v in ge.values())
    print("gen grad rel-L2: median %.4f  p90 %.4f  max %.4f" % (vals[len(vals) // 2], vals[int(len(vals) * .9)], vals[-1]))
    for k, v in ge.items():
        print("  %-50s %.4f  |g|=%.3e" % (k, v[0], v[1]))
    vals = sorted(v[0] for v in de.values())
    print("dis grad rel-L2: median %.4f max %.4f" % (vals[len(vals) // 2], vals[-1]))
    for k, v in de.items():
        print("  %-50s %.4f  |g|=%.3e" % (k, v[0], v[1]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_step.json"), "w"), indent=1)
