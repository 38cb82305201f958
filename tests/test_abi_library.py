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
    # int32 x4, int32 x3, (pad), double x3, double x3, int32, (pad), double, int32 x3, int32 x3,
    # cell_order + refine_lo[3] + refine_hi[3] + reserved[1] = int32 x8
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
    assert sizes[2] == 5 * 8 + 4 * 4 + 3 * 81 * 8 + 2 * 81 * 8     # + hanging_interpolation[2][81]


def test_facade_headers_compile_without_cuda(tmp_path):
    """the host facade is header-only C++17 and must not need nvcc or the CUDA headers"""
    import subprocess
    src = tmp_path / "facade_probe.cc"
    src.write_text('#include "dealii_b200/dealii_b200.h"\nint main() { dealii::Triangulation<3> t; '
                   'dealii::GridGenerator::hyper_cube(t); t.refine_global(2); return t.n_global_active_cells() == 64 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_cpp_drivers_fail_loudly_without_a_gpu():
    """the C++ drivers on the facade (single block and forked ranks) exit 1 with the library's message on a box
    without a GPU -- no CPU fallback, and no rank left hanging in a barrier"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    for cmd in (["bp5_step64", "--cycle-min", "7", "--cycle-max", "7"],
                ["bp5_step64_multi", "--ranks", "2", "--cycle-min", "7", "--cycle-max", "7"]):
        out = subprocess.run([os.path.join(ROOT, "build", "examples", cmd[0])] + cmd[1:], capture_output=True, text=True,
                             timeout=60)
        assert out.returncode == 1, (cmd, out.returncode)
        assert "no CPU fallback" in out.stderr, out.stderr[-500:]


def test_facade_process_grid_matches_the_python_layer(tmp_path):
    """the block grid a C++ host gets from dealii::b200::process_grid is the one distributed.process_grid gives the
    Python layer (2x1x1, 2x2x1, 2x2x2, ...): both sides of a mixed deployment must agree on who is whose neighbour"""
    import subprocess
    from dealceed_b200.distributed import process_grid
    src = tmp_path / "grid_probe.cc"
    src.write_text('#include <cstdio>\n#include "dealii_b200/dealii_b200.h"\nint main() { for (int w = 1; w <= 24; ++w) { '
                   'auto g = dealii::b200::process_grid(w); std::printf("%d %d %d\\n", g[0], g[1], g[2]); } return 0; }\n')
    exe = tmp_path / "grid_probe"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "deal-and-ceed-on-gpu_b200"), "-lbp5b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "deal-and-ceed-on-gpu_b200")])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    for w in range(1, 25):
        assert tuple(int(v) for v in lines[w - 1].split()) == tuple(process_grid(w)), w


def test_generic_functor_path_rejects_a_partitioned_triangulation(tmp_path):
    """CUDAWrappers::MatrixFree (the reference's user-functor interface) drives one block: asked to reinit on a
    triangulation split over two ranks it throws instead of silently building the whole mesh on every rank (or
    silently ignoring overlap_communication_computation, bp5/step-64.cu:241).  The check runs before any device call,
    so the probe needs no GPU."""
    import subprocess
    src = tmp_path / "mf_probe.cu"
    src.write_text(r'''
#include <cstdio>
#include <cstring>
#include "dealii_b200/dealii_b200.h"
#include "dealii_b200/cuda_matrix_free.cuh"
using namespace dealii;
struct TwoRanks : b200::Communicator {
  int rank() const override { return 0; }
  int size() const override { return 2; }
  void allgather(const void *s, void *r, std::size_t n) override { std::memcpy(r, s, n); std::memcpy((char *)r + n, s, n); }
  void barrier() override {}
};
int main() {
  TwoRanks comm;
  Triangulation<3> tria(&comm);
  Point<3> p2;
  for (int d = 0; d < 3; ++d) p2[d] = 4.;
  GridGenerator::subdivided_hyper_rectangle(tria, std::vector<unsigned int>(3, 4), Point<3>(), p2);
  FE_Q<3> fe(2);
  DoFHandler<3> dof_handler(tria);
  dof_handler.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  CUDAWrappers::MatrixFree<3, double> mf;
  CUDAWrappers::MatrixFree<3, double>::AdditionalData ad;
  ad.overlap_communication_computation = true;
  try {
    mf.reinit(MappingQGeneric<3>(2), dof_handler, constraints, QGauss<1>(3), ad);
  } catch (const ExcMessage &e) {
    std::printf("%s\n", e.what());
    return std::strstr(e.what(), "partitioned mesh") ? 0 : 2;
  }
  return 1;
}
''')
    exe = tmp_path / "mf_probe"
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O0",
                           "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.join(ROOT, "deal-and-ceed-on-gpu_b200"), "-lbp5b200",
                           "-Xlinker", "-rpath", "-Xlinker", os.path.join(ROOT, "deal-and-ceed-on-gpu_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr[-500:])
