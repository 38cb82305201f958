"""The C-ABI library loads on a CPU-only box and exports every symbol include/bp5_b200.h
declares; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bp5_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp5_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import dealceed_b200 as dc
    L = ctypes.CDLL(dc.bindings.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    bound = {name for name, _, _ in dc.bindings.ABI}
    assert set(syms) == bound, (set(syms) ^ bound)


def test_no_cpu_fallback():
    import torch
    import dealceed_b200 as dc
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dc.Bp5Error) as e:
        dc.Context(0)
    assert e.value.code == dc.bindings.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    bad = []
    pkg = os.path.join(ROOT, "deal-and-ceed-on-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle|/root/reference", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_problem_struct_layout_matches_header():
    import dealceed_b200 as dc
    # int32 x4, int32 x3, (pad), double x3, double x3, int32, (pad), double, int32 x3, int32 x3, int32 x8
    assert ctypes.sizeof(dc.bindings.Problem) == 16 + 12 + 4 + 24 + 24 + 4 + 4 + 8 + 12 + 12 + 32
