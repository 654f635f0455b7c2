"""The C-ABI shared library loads and exports exactly what include/ma_b200.h declares (no compute)."""
import os
import re
import subprocess

import pytest

from mongeampere_b200 import build as mb
from mongeampere_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return mb.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(ma_[a-z0-9_]+)\s*\(", src))


def test_header_and_binding_agree():
    assert header_symbols() == set(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = header_symbols() - exported
    assert not missing, missing


def test_library_is_sm100a_only(libpath):
    out = subprocess.check_output(["cuobjdump", "--list-elf", libpath], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_loads_with_ctypes_and_reports_version(libpath):
    L = capi.load_library()
    assert L.ma_abi_version() == 3


def test_no_cpu_fallback_without_device(libpath):
    """Without a CUDA device ma_create must fail loudly with MA_CUDA_ERROR (never compute on the CPU)."""
    import ctypes as C
    L = capi.load_library()
    h = C.c_void_p()
    rc = L.ma_create(C.byref(h), 0)
    if rc == capi.MA_OK:  # a GPU is present (GPU box): nothing to check here
        L.ma_destroy(h)
        pytest.skip("CUDA device present")
    assert rc == capi.MA_CUDA_ERROR
    assert b"no CPU fallback" in L.ma_last_error(h)
    L.ma_destroy(h)
    with pytest.raises(capi.MAError):
        capi.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mongeampere_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)
