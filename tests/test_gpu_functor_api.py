"""-m gpu: the DEVICE-side drop-in boundary.  examples/bp5_functors.cu holds the reference's operators
written as device functors on CUDAWrappers::MatrixFree / FEEvaluationGL (bp5/step-64.cu:60-276,
step-64/step-64.cu:69-322, bp5/fe_evaluation_gl.h:31-98) and compiled with nvcc against
include/dealii_b200/cuda_matrix_free.cuh.  The program itself compares the user-written operators with the
library's tuned kernel (<= 1e-12); here its printed norms and iteration counts are checked against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CELLS, EPS = (3, 2, 2), 0.1


def _run(degree, quad, dump_dir=None):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    cmd = [os.path.join(ROOT, "build", "examples", "bp5_functors"), str(degree), quad, *[str(c) for c in CELLS], str(EPS)]
    if dump_dir is not None:
        cmd += ["1.0", os.path.join(str(dump_dir), "v_")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.strip().endswith("OK"), out.stdout[-3000:]
    vals = {}
    for line in out.stdout.splitlines():
        m = re.match(r"^(\w+) (\S+)$", line.strip())
        if m:
            vals[m.group(1)] = float(m.group(2))
    return vals


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("degree", range(1, 9))
@pytest.mark.parametrize("quad", ["gauss", "gll"])
def test_user_functors_match_library_and_oracle(degree, quad, tmp_path):
    import oracle as O
    v = _run(degree, quad, tmp_path)
    # use_coloring mode: eight colour passes with plain stores, bitwise reproducible
    assert v["bp5_colored_bitwise_reproducible"] == 1
    if quad == "gauss":
        assert v["helmholtz_colored_bitwise_reproducible"] == 1
    q = O.GLL if quad == "gll" else O.GAUSS
    m = O.OracleMesh(degree, CELLS, quad=q, deform=1, eps=EPS)
    assert v["n_dofs"] == m.n_dofs
    b = m.rhs()
    Ab = m.vmult(b)
    AAb = m.vmult(Ab)
    assert v["bp5_norm_b"] == pytest.approx(np.linalg.norm(b), rel=1e-12)
    assert v["bp5_norm_Ab"] == pytest.approx(np.linalg.norm(Ab), rel=1e-12)       # fp64 operator: 1e-12
    assert v["bp5_norm_AAb"] == pytest.approx(np.linalg.norm(AAb), rel=1e-12)
    x, its, res, hist, ok = m.cg(b, variant=1, control=0, tol=1e-6 * np.linalg.norm(b), max_its=200)
    assert abs(v["bp5_merged_its_user"] - its) <= 1                               # same count +-1
    assert abs(v["bp5_standard_its_user"] - its) <= 1
    assert v["bp5_norm_x"] == pytest.approx(np.linalg.norm(x), rel=1e-7)
    # the vectors themselves, not only their norms
    assert _rel(np.fromfile(tmp_path / "v_bp5_Ab.f64"), Ab) <= 1e-12
    assert _rel(np.fromfile(tmp_path / "v_bp5_x.f64"), x) <= 1e-7
    if quad == "gauss":
        Hb = m.vmult(b, kind=O.HELMHOLTZ)
        assert _rel(np.fromfile(tmp_path / "v_helmholtz_Ab.f64"), Hb) <= 1e-12
        assert v["helmholtz_norm_Ab"] == pytest.approx(np.linalg.norm(Hb), rel=1e-12)
        xh, its_h, _, _, _ = m.cg(b, kind=O.HELMHOLTZ, variant=1, control=1, tol=1e-12 * np.linalg.norm(b),
                                  max_its=m.n_dofs)
        assert abs(v["helmholtz_merged_its_user"] - its_h) <= 1
        assert abs(v["helmholtz_standard_its_library"] - its_h) <= 1
        assert v["helmholtz_norm_x"] == pytest.approx(np.linalg.norm(xh), rel=1e-7)


def test_step64_tutorial_through_user_functors():
    """the reference's own step-64 run (step-64/step-64.cu:610-631: Q3, unit cube refined once, 343 DoFs;
    upstream tutorial output: 27 iterations) through the user-written LocalHelmholtzOperator"""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "build", "examples", "bp5_functors"), "3", "gauss", "2", "2", "2", "0",
                          "0.5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "n_dofs 343" in out.stdout
    its = int(re.search(r"helmholtz_merged_its_user (\d+)", out.stdout).group(1))
    assert abs(its - 27) <= 1
