"""The C-ABI shared library loads without a GPU and exports every symbol include/b200xgb.h declares (no compute here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200xgb.h")
LIB = os.path.join(ROOT, "sagemaker-xgboost-container_b200", "lib", "libb200xgb.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(LIB)


def declared_symbols():
    txt = open(HEADER).read()
    return re.findall(r"^XGB_DLL\s+[\w\s\*]+?\b(XG\w+)\s*\(", txt, flags=re.M)


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("XGDMatrixCreateFromMat", "XGBoosterUpdateOneIter", "XGBoosterPredictFromDMatrix", "XGBoosterSaveModel", "XGBoosterEvalOneIter",
                 "XGCommunicatorInit", "XGBGetLastError"):
        assert must in syms
    assert len(syms) >= 45


def test_every_declared_symbol_is_exported(lib):
    nm = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", nm))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    for s in declared_symbols():
        assert hasattr(lib, s)


def test_version_and_build_info_need_no_gpu(lib):
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.XGBoostVersion(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    assert (a.value, b.value, c.value) == (3, 0, 5)
    out = ctypes.c_char_p()
    assert lib.XGBuildInfo(ctypes.byref(out)) == 0
    assert b"sm_100a" in out.value and b'"CPU_FALLBACK":false' in out.value


def test_no_cpu_fallback_fails_loudly_without_gpu(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    data = (ctypes.c_float * 4)(1, 2, 3, 4)
    h = ctypes.c_void_p()
    rc = lib.XGDMatrixCreateFromMat(data, ctypes.c_uint64(2), ctypes.c_uint64(2), ctypes.c_float(float("nan")), ctypes.byref(h))
    assert rc == -1
    lib.XGBGetLastError.restype = ctypes.c_char_p
    msg = lib.XGBGetLastError().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sagemaker-xgboost-container_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "libgbt_oracle" not in src, os.path.join(dirpath, f)
