"""Diagnostic: content-encoder activations layer by layer, munit_b200 vs storage-aware oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import munit_oracle as O  # noqa: E402
from tests.gpu_util import rel_l2, nchw  # noqa: E402
from munit_b200.networks import AdaINGen_double, Conv2dBlock  # noqa: E402

O.QUANT = True
cfg = O.config_256_core()
gsd = O.init_state_dict(O.gen_spec(cfg["gen"], 3, True), 21, "kaiming")
gen = AdaINGen_double(3, cfg["gen"])
gen.load_state_dict(gsd)
gen = gen.cuda()
g = torch.Generator().manual_seed(1234)
x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1

acts = []
orig = Conv2dBlock.forward_act


def hooked(self, x, out_pad=0, upsample=1, residual=None, frozen=False):
    a = orig(self, x, out_pad, upsample, residual, frozen)
    p = a.pad
    t = a.t[:, p:a.t.shape[1] - p, p:a.t.shape[2] - p] if p else a.t
    acts.append(nchw(t.float()).cpu())
    return a


Conv2dBlock.forward_act = hooked
with torch.no_grad():
    gen.enc1_content.forward_act(x.cuda(), 1)
    p = "enc1_content."
    refs = []
    y = O.conv_block(gsd, f"{p}model.0.", x, 1, 3, "in", "relu", image=True); refs.append(y)
    for i in (1, 2):
        y = O.conv_block(gsd, f"{p}model.{i}.", y, 2, 1, "in", "relu"); refs.append(y)
    xx = y
    for r in range(4):
        y = O.conv_block(gsd, f"{p}model.3.model.{r}.model.0.", xx, 1, 1, "in", "relu"); refs.append(y)
        xx = O.conv_block(gsd, f"{p}model.3.model.{r}.model.1.", y, 1, 1, "in", "none", residual=xx); refs.append(xx)
for i, (a, r) in enumerate(zip(acts, refs)):
    d = (a - r).abs()
    print(i, tuple(a.shape), "rel_l2 %.5f  frac_diff %.5f  max %.4f" % (rel_l2(a, r), float((d > 0).float().mean()), float(d.max())))
# re-run each of our layers on the ORACLE's input to see per-layer fresh error
