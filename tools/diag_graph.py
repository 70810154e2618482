"""Diagnostic: which part of the step breaks CUDA graph capture."""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from munit_b200.engine import StepRunner
from munit_b200.trainer import MUNIT_Trainer
from munit_b200 import kernels as K

cfg = bench.load_cfg()
torch.manual_seed(0)
t = MUNIT_Trainer(cfg).cuda()
r = StepRunner(t, cfg, 2, 64, use_graph=False)
xa, xb = bench.synthetic_images(2, 64, 1)
r.x_a.copy_(xa); r.x_b.copy_(xb)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        r._prepare_host_state(); r._eager_step(); r._advance()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
print("eager ok", float(t.loss_gen_total))

def try_capture(name, fn, mode="global"):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK  ", name, mode)
        return True
    except Exception as e:
        print("CAPTURE FAIL", name, mode, str(e).splitlines()[0][:150])
        torch.cuda.synchronize()
        return False

y = torch.randn(2, 16, 16, 64, device="cuda").to(torch.bfloat16)
try_capture("norm_stats", lambda: K.norm_stats(y))
try_capture("fill", lambda: K.fill(torch.empty(100, device="cuda"), 1.0))
def fwd_only():
    with torch.no_grad():
        c, s_ = t._enc("a", r.x_a)
        t._dec("b", c, s_)
try_capture("gen forward no_grad", fwd_only)
def dis_fwd():
    l = t.dis_a.calc_dis_loss(r.x_a, r.x_b)
    return l
try_capture("dis forward (autograd on)", dis_fwd)
def dis_fb():
    t.dis_opt.zero_grad()
    l = t.dis_a.calc_dis_loss(r.x_a, r.x_b)
    l.backward()
for m in ("global", "thread_local", "relaxed"):
    try_capture("dis fwd+bwd", dis_fb, m)
try_capture("upload_hyper", lambda: t.dis_opt.upload_hyper(t.dis_opt.step_count + 1))
try_capture("dis_opt_step", lambda: t.dis_opt_step())
for m in ("global", "thread_local", "relaxed"):
    try_capture("_seg_dis", r._seg_dis, m)
    try_capture("_seg_mid", r._seg_mid, m)
