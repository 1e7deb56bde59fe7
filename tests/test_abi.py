"""The C-ABI library loads and exports exactly the symbols include/munk_b200.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "munk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(munk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from munk_b200 import build
    so = build.build()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    from munk_b200._lib import SIGNATURES
    assert sorted(SIGNATURES) == names                       # ctypes table mirrors the header
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (munk_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == names                                 # nothing undeclared leaks out


def test_error_reporting_without_gpu():
    from munk_b200._lib import lib
    assert lib.munk_version() == 100
    assert lib.munk_linv_bytes(129) == 2 * 128 * 128 * 8
    assert lib.munk_dgemm(None, 0, 0, 1, 1, 1, 1.0, None, 2, None, 2, 0.0, None, 2, 0, None) == -1


def test_product_has_no_oracle_or_cpu_fallback():
    """Nothing under munk_b200/ may import the oracle (tier rule 3)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "munk_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
