"""Diagnostic: first-update generator gradients with / without trainer.reuse_forward, per parameter."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import munit_oracle as O
from munit_b200.trainer import MUNIT_Trainer

def run(reuse, gan_w=0):
    cfg = O.config_256_core(gan_w=gan_w)
    g = torch.Generator().manual_seed(9)
    xa = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    xb = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    torch.manual_seed(3)
    t = MUNIT_Trainer(cfg).cuda()
    t.reuse_forward = reuse
    torch.manual_seed(11)
    t.iterations = 0
    t.dis_update(xa, xb, cfg)
    t.gen_update(xa, xb, cfg)
    torch.cuda.synchronize()
    global LAST
    LAST = {k: v.clone() for k, v in t._last.items()}
    return {n: p.grad.clone() for n, p in t.gen.named_parameters()}, {k: float(getattr(t, k)) for k in ("loss_gen_total", "loss_gen_recon_x_a", "loss_gen_recon_c_a", "loss_gen_recon_s_a", "loss_gen_cycrecon_x_a")}

a, la = run(False); xa_ = LAST
b, lb = run(False); xb_ = LAST
c, lc = run(True); xc_ = LAST
print("x_ab bitwise equal noreuse/noreuse:", torch.equal(xa_["x_ab"], xb_["x_ab"]), " noreuse/reuse:", torch.equal(xa_["x_ab"], xc_["x_ab"]),
      float((xa_["x_ab"] - xc_["x_ab"]).abs().max()))
print("losses run1", la); print("losses run2", lb)
worst2 = sorted(((float(torch.dot(a[n].flatten(), b[n].flatten()) / (a[n].norm() * b[n].norm() + 1e-30)), n) for n in a if a[n].norm() > 0))[:8]
print("noreuse vs noreuse worst:", worst2)
best2 = sorted(((float(torch.dot(a[n].flatten(), b[n].flatten()) / (a[n].norm() * b[n].norm() + 1e-30)), n) for n in a if a[n].norm() > 0))[-8:]
print("noreuse vs noreuse best:", best2)
cos = lambda u, v: float(torch.dot(u.flatten(), v.flatten()) / (u.norm() * v.norm() + 1e-30))
print("losses", la, lc)
fa = torch.cat([v.flatten() for v in a.values()]); fb = torch.cat([v.flatten() for v in b.values()]); fc = torch.cat([v.flatten() for v in c.values()])
print("noreuse vs noreuse", cos(fa, fb), " noreuse vs reuse", cos(fa, fc))
worst = sorted(((cos(a[n], c[n]), n, float(a[n].norm()), float(c[n].norm())) for n in a if a[n].norm() > 0))[:6]
for w in worst:
    print("%.5f %-50s |g| %.4e %.4e" % w)
