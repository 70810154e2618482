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


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference
    arm may import it.  The package and the drop-in scripts must not (a product path that routes through the CPU
    restatement would void every parity claim)."""
    import glob

    offenders = []
    files = glob.glob(os.path.join(ROOT, "munit_b200", "**", "*.py"), recursive=True)
    files += glob.glob(os.path.join(ROOT, "scripts", "*.py"))
    for path in files:
        src = open(path).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "munit_oracle" in src or "ref_loader" in src:
            offenders.append(os.path.relpath(path, ROOT))
    assert not offenders, offenders
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py: every oracle import sits inside the CPU-baseline / reference-arm function
    for m in re.finditer(r"^(\s*)from oracle import", bench, flags=re.M):
        assert len(m.group(1)) >= 4, "oracle import at module level of bench.py"
    # ... and only inside the CPU-baseline helpers (_cpu_trainer / cpu_reference_infer), never in the B200 arm
    for m in re.finditer(r"^    from oracle import", bench, flags=re.M):
        owner = re.findall(r"^def (\w+)\(", bench[: m.start()], flags=re.M)[-1]
        assert owner in ("_cpu_trainer", "cpu_reference_infer"), f"oracle import inside {owner}()"
