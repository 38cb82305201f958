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


def test_header_is_plain_c_and_struct_sizes_match_the_ctypes_mirror(tmp_path):
    """include/bp5_b200.h must compile as C99 (it is the FFI surface: no C++, no CUDA, no torch types), and the
    structs that cross the boundary by value/pointer must have the layout the ctypes bindings assume"""
    import subprocess
    import dealceed_b200 as dc
    src = tmp_path / "abi_probe.c"
    src.write_text('#include <stdio.h>\n#include "bp5_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(bp5_problem_t), sizeof(bp5_peer_info_t), '
                   'sizeof(bp5_matrix_free_data_t)); return 0; }\n')
    exe = tmp_path / "abi_probe"
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == ctypes.sizeof(dc.bindings.Problem)
    assert sizes[1] == ctypes.sizeof(dc.bindings.PeerInfo)
    assert sizes[2] == 5 * 8 + 4 * 4 + 3 * 81 * 8


def test_facade_headers_compile_without_cuda(tmp_path):
    """the host facade is header-only C++17 and must not need nvcc or the CUDA headers"""
    import subprocess
    src = tmp_path / "facade_probe.cc"
    src.write_text('#include "dealii_b200/dealii_b200.h"\nint main() { dealii::Triangulation<3> t; '
                   'dealii::GridGenerator::hyper_cube(t); t.refine_global(2); return t.n_global_active_cells() == 64 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
