"""The C-ABI library loads without a GPU and exports every symbol include/munit_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "munit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(munit_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"munit_status"}
    assert len(declared) > 30
    lib = ctypes.CDLL(os.path.join(ROOT, "munit_b200", "libmunit_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_matches_header():
    from munit_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "munit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(munit_[a-z0-9_]+)\s*\(", hdr))
    assert set(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ declared
    assert _lib.lib.munit_version() == 100


def test_no_fallback_without_gpu():
    import pytest
    import torch

    from munit_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MunitError):
        _lib.check(_lib.lib.munit_init(), "munit_init")
