"""Diagnostic: find the first op that invalidates stream capture (wraps every kernels.* wrapper)."""
import ctypes, os, sys, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from munit_b200.engine import StepRunner
from munit_b200.trainer import MUNIT_Trainer
from munit_b200 import kernels as K

cudart = ctypes.CDLL("libcudart.so.12")
def cap_status():
    st = ctypes.c_int(-1)
    rc = cudart.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value

first_bad = []
def wrap(name, fn):
    def inner(*a, **k):
        before = cap_status()
        out = fn(*a, **k)
        after = cap_status()
        if (after != before or after[0] != 0 or after[1] == 2) and not first_bad:
            first_bad.append((name, before, after, threading.current_thread().name, torch.cuda.current_stream().cuda_stream))
            print("FIRST BAD OP:", first_bad[0], flush=True)
        return out
    return inner
for n in dir(K):
    f = getattr(K, n)
    if callable(f) and not n.startswith("_") and getattr(f, "__module__", "") == K.__name__:
        setattr(K, n, wrap(n, f))

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
print("eager ok; main thread", threading.current_thread().name)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        print("capture stream", torch.cuda.current_stream().cuda_stream, cap_status())
        t.dis_opt.zero_grad()
        l = t.dis_a.calc_dis_loss(r.x_a, r.x_b)
        print("after fwd", cap_status())
        l.backward()
        print("after bwd", cap_status())
except Exception as e:
    print("FAIL", str(e).splitlines()[0])
import traceback
torch.cuda.synchronize()
print("---- retry with autograd multithreading disabled")
first_bad.clear()
g2 = torch.cuda.CUDAGraph()
try:
    with torch.autograd.set_multithreading_enabled(False):
        with torch.cuda.graph(g2):
            t.dis_opt.zero_grad()
            l = t.dis_a.calc_dis_loss(r.x_a, r.x_b)
            l.backward()
            print("after bwd (single thread)", cap_status())
    g2.replay(); torch.cuda.synchronize(); print("single-thread capture OK")
except Exception as e:
    traceback.print_exc()
print("---- native-only backward in capture")
w = torch.randn(64, 64, device="cuda", requires_grad=True)
g3 = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g3):
        (w @ w).sum().backward()
    print("native capture OK")
except Exception as e:
    print("native FAIL", str(e).splitlines()[0])
