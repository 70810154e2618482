"""Load the *unmodified* reference (cc-ai/MUNIT) on CPU -- TEST INFRASTRUCTURE.

Only usable where /root/reference exists (the build container); nothing on the GPU box
may call this.  Two runtime shims, no source edits:
  1. scripts/extraadam.py has no imports -> exec it in a namespace pre-seeded with
     torch / math / Optimizer and register it as sys.modules["extraadam"];
  2. trainer.py hard-codes .cuda() (trainer.py:94-95,366-367,1146-1147) -> identity on CPU;
  3. domainClassifier passes stride=True to BasicBlock (utils.py:1375,1377), which torch 0.4.1 accepted as the
     integer 1 and torch 2.x's conv2d rejects as a tuple of bools -> utils.conv3x3 / conv1x1 are wrapped so the
     stride reaches nn.Conv2d as int(stride) (same value, bool is an int subclass).
"""
from __future__ import annotations

import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MUNIT_REFERENCE", "/root/reference")
REF_SCRIPTS = os.path.join(REF_ROOT, "scripts")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SCRIPTS, "networks.py"))


_loaded = {}


def load():
    """Returns (networks, trainer) reference modules."""
    if _loaded:
        return _loaded["networks"], _loaded["trainer"]
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    if REF_SCRIPTS not in sys.path:
        sys.path.insert(0, REF_SCRIPTS)
    # shim 1: extraadam imports
    src = open(os.path.join(REF_SCRIPTS, "extraadam.py")).read()
    mod = types.ModuleType("extraadam")
    mod.__dict__.update(torch=torch, math=math, Optimizer=torch.optim.Optimizer)
    exec(compile(src, os.path.join(REF_SCRIPTS, "extraadam.py"), "exec"), mod.__dict__)
    sys.modules["extraadam"] = mod
    # shim 2: .cuda() -> identity
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import networks  # noqa
    import trainer  # noqa
    import utils as ref_utils  # noqa
    # shim 3: stride=True -> 1
    _c3, _c1 = ref_utils.conv3x3, ref_utils.conv1x1
    ref_utils.conv3x3 = lambda i, o, stride=1, groups=1, dilation=1: _c3(i, o, int(stride), groups, dilation)
    ref_utils.conv1x1 = lambda i, o, stride=1: _c1(i, o, int(stride))
    _loaded.update(networks=networks, trainer=trainer, extraadam=mod)
    return networks, trainer


def make_trainer(cfg: dict):
    """Reference MUNIT_Trainer on CPU with `cfg` (see munit_oracle.config_256_core)."""
    _, trainer = load()
    t = trainer.MUNIT_Trainer(cfg)
    t.iterations = 0
    return t
