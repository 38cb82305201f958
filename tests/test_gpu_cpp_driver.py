"""-m gpu: the host-only C++ drivers written against the dealii_b200 facade (examples/), i.e. the
reference's own experiment through the C ABI: BP5 ladder cycles 7-8 at degree 5 (bp5/step-64.cu:724-730)
and the step-64 Helmholtz tutorial run."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)


def test_bp5_driver_reproduces_ladder_fixture():
    _build()
    gold = json.load(open(os.path.join(GOLD, "bp5_ladder_p5.json")))
    for quad in ("gauss", "gll"):
        out = subprocess.run([os.path.join(ROOT, "build", "examples", "bp5_step64"), "--degree", "5", "--cycle-min", "7",
                              "--cycle-max", "8", "--repetitions", "2", "--quadrature", quad],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        text = out.stdout
        for tag in ("pcg-standard", "pcg-merged", "vmult"):
            assert len(re.findall(rf"^{tag} \d+ [0-9.e+]+$", text, flags=re.M)) == 2, text[-1500:]
        blocks = text.split("Cycle ")[1:]
        for blk, cyc in zip(blocks, (7, 8)):
            g = gold[f"cycle{cyc}_{quad}"]
            assert f"Number of degrees of freedom: {g['n_dofs']}" in blk
            solved = re.findall(r"Solved in (\d+) iterations with time \S+ and DoFs/s \S+ norm (\S+)", blk)
            assert len(solved) == 4            # 2 repetitions x (standard, merged)
            for its, norm in solved:
                assert abs(int(its) - g["its_merged"]) <= 1
                assert float(norm) == pytest.approx(g["x_l2"], rel=1e-5)


def test_bp5_driver_on_a_locally_refined_mesh_matches_the_hanging_oracle():
    """the reference's driver with the corner octant of ladder cycle 7 (3 x 2 x 2 cells) refined once more: DoF count,
    iteration count of both solvers and the two printed norms against oracle/hanging_oracle.py"""
    import numpy as np
    import oracle as O
    from hanging_oracle import HangingMesh
    _build()
    p = 3
    hm = HangingMesh(p, (3, 2, 2), (0, 0, 0), (1, 1, 1), quad=O.GAUSS, upper=(3., 2., 2.))
    b = hm.rhs()
    x, its, _ = hm.cg(b, tol=1e-6 * np.linalg.norm(b), max_its=200)
    out = subprocess.run([os.path.join(ROOT, "build", "examples", "bp5_step64"), "--degree", str(p), "--cycle-min", "7",
                          "--cycle-max", "7", "--repetitions", "1", "--refine-corner", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    text = out.stdout
    assert f"Number of active cells:       {hm.n_cells}" in text and f"Number of degrees of freedom: {hm.n_dofs}" in text
    solved = re.findall(r"Solved in (\d+) iterations with time \S+ and DoFs/s \S+ norm (\S+)", text)
    assert len(solved) == 2
    for it, norm in solved:
        assert abs(int(it) - its) <= 1
        assert float(norm) == pytest.approx(np.linalg.norm(x), rel=1e-5)
    # after the vmult block the driver's solution vector still holds the merged solve's x
    assert float(re.search(r"solution norm: (\S+)", text).group(1)) == pytest.approx(hm.l2_norm(x), rel=1e-4)


def test_step64_driver_reproduces_tutorial_iterations():
    _build()
    out = subprocess.run([os.path.join(ROOT, "build", "examples", "step64_helmholtz"), "2"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    its = [int(v) for v in re.findall(r"Solved in (\d+) iterations", out.stdout)]
    assert len(its) == 4                        # (SolverCG, merged) x 2 cycles
    for got, ref in zip(its, (27, 60, 27, 60)):  # deal.II step-64 tutorial: 343 DoFs -> 27, 2197 -> 60
        assert abs(got - ref) <= 1
    norms = re.findall(r"solution norm: (\S+)", out.stdout)    # tutorial output: 0.0205439, 0.0205269
    assert [f"{float(v):.6g}" for v in norms] == ["0.0205439", "0.0205269", "0.0205439", "0.0205269"]


def test_bp5_multi_rank_driver_reproduces_ladder_fixture():
    """host code in C++ only on a PARTITIONED mesh (examples/bp5_step64_multi.cc): 2 forked ranks, partitioned
    operator, owned + ghost vectors, peer-memory halo and all-rank sums behind the facade; same ladder fixture as the
    single-block driver, and cell_loop bracketed by update_ghost_values / compress(add) must equal vmult.
    On a 1-GPU box both ranks share cuda:0 (CUDA IPC works across processes on one device)."""
    import torch
    _build()
    gold = json.load(open(os.path.join(GOLD, "bp5_ladder_p5.json")))
    devices = max(1, min(2, torch.cuda.device_count()))
    for quad in ("gauss", "gll"):
        out = subprocess.run([os.path.join(ROOT, "build", "examples", "bp5_step64_multi"), "--ranks", "2", "--devices",
                              str(devices), "--degree", "5", "--cycle-min", "7", "--cycle-max", "8", "--repetitions", "1",
                              "--quadrature", quad], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-2000:]
        blocks = out.stdout.split("Cycle ")[1:]
        assert len(blocks) == 2
        for blk, cyc in zip(blocks, (7, 8)):
            g = gold[f"cycle{cyc}_{quad}"]
            assert f"Number of degrees of freedom: {g['n_dofs']}" in blk
            solved = re.findall(r"Solved in (\d+) iterations with time \S+ and DoFs/s \S+ norm (\S+)", blk)
            assert len(solved) == 2            # standard, merged
            for its, norm in solved:
                assert abs(int(its) - g["its_merged"]) <= 1
                assert float(norm) == pytest.approx(g["x_l2"], rel=1e-5)
            ghost = re.findall(r"ghost semantics: .* = (\S+)", blk)
            assert len(ghost) == 1 and float(ghost[0]) <= 1e-12
