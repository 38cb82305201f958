"""Pins the oracle (CPU).  The reference holds no golden vectors for this path
(SURVEY.md section 4); the external anchor is the published output of the upstream deal.II
step-64 tutorial, of which step-64/step-64.cu is a modified copy."""
import json
import os

import numpy as np
import pytest

import oracle as O
from conftest import ladder

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# deal.II step-64 tutorial, Q3 Helmholtz a(x)=10/(0.05+2|x|^2) on the unit cube, SolverCG,
# tol 1e-12*|b| (step-64/step-64.cu:513-514): refinement -> (DoFs, iterations, solution norm)
STEP64 = {1: (343, 27, 0.0205439), 2: (2197, 60, 0.0205269), 3: (15625, 114, 0.0205261)}


@pytest.mark.parametrize("refine", [1, 2, 3])
@pytest.mark.parametrize("variant", [0, 1])
def test_step64_tutorial_known_answers(refine, variant):
    n_dofs, its_ref, norm_ref = STEP64[refine]
    c = 2 ** refine
    m = O.OracleMesh(3, (c, c, c), quad=O.GAUSS, upper=(1., 1., 1.))
    assert m.n_dofs == n_dofs
    b = m.rhs()
    x, its, res, hist, ok = m.cg(b, kind=O.HELMHOLTZ, variant=variant, control=1, tol=1e-12 * np.linalg.norm(b),
                                 max_its=m.n_dofs)
    assert ok
    assert its == its_ref
    assert f"{m.l2_norm(x):.6g}" == f"{norm_ref:.6g}"      # every printed digit


def test_bp5_ladder_matches_committed_fixture():
    gold = json.load(open(os.path.join(GOLD, "bp5_ladder_p5.json")))
    for cyc in (7, 8):
        cells, upper = ladder(cyc)
        for quad, qn in ((O.GAUSS, "gauss"), (O.GLL, "gll")):
            g = gold[f"cycle{cyc}_{qn}"]
            m = O.OracleMesh(5, cells, quad=quad, upper=upper)
            b = m.rhs()
            x, its, res, hist, ok = m.cg(b, variant=1, control=0, tol=1e-6 * np.linalg.norm(b), max_its=200)
            assert m.n_dofs == g["n_dofs"] and its == g["its_merged"]
            assert np.linalg.norm(x) == pytest.approx(g["x_l2"], rel=1e-10)
            assert np.linalg.norm(b) == pytest.approx(g["b_l2"], rel=1e-12)


def test_bp5_ladder_survey_probe_values():
    # SURVEY.md 8(c) anchor 2 (independent numpy probe, QGauss): cycle -> (its, |x|_2, |b|_2)
    probe = {7: (37, 4.581847, 0.3771922), 8: (42, 1.452966, 0.05448710), 12: (56, 2.223825, 0.01368822)}
    gold = json.load(open(os.path.join(GOLD, "bp5_ladder_p5.json")))
    for cyc, (its, xn, bn) in probe.items():
        g = gold[f"cycle{cyc}_gauss"]
        assert g["its_merged"] == its and g["its_standard"] == its
        assert f"{g['x_l2']:.7g}" == f"{xn:.7g}" and f"{g['b_l2']:.7g}" == f"{bn:.7g}"


def test_config1_survey_probe_values():
    # |b| = 8.37489887e-4, |x_200| = 36.1851795, rel. residual 2.25e-3 at the cap
    g = json.load(open(os.path.join(GOLD, "config1_p4_32.json")))
    assert g["n_dofs"] == 2146689 and g["its"] == 200
    assert f"{g['b_l2']:.9g}" == "0.000837489887"
    assert f"{g['x_l2']:.9g}" == "36.1851795"
    assert f"{g['rel_res']:.3g}" == "0.00225"


def test_merged_cg_variants():
    """Correct-parity merged CG == textbook CG; the as-shipped x update (solver.h:425) is wrong
    for it >= 4 but leaves the residual history and iteration count unchanged (SURVEY finding 4)."""
    m = O.OracleMesh(3, (3, 2, 2), quad=O.GAUSS)
    b = m.rhs(); tol = 1e-6 * np.linalg.norm(b)
    x0, it0, r0, h0, _ = m.cg(b, variant=0, tol=tol)
    x1, it1, r1, h1, _ = m.cg(b, variant=1, tol=tol)
    x2, it2, r2, h2, _ = m.cg(b, variant=2, tol=tol)
    assert it0 == it1 == it2
    np.testing.assert_allclose(h1, h0, rtol=1e-9)
    np.testing.assert_allclose(h2, h1, rtol=1e-13)
    assert np.linalg.norm(x1 - x0) <= 1e-9 * np.linalg.norm(x0)
    assert np.linalg.norm(x2 - x0) > 0.1 * np.linalg.norm(x0)
    # diagonal preconditioner slot
    diag = 1.0 / (1.0 + np.arange(m.n_dofs) % 3)
    xa, ita, *_ = m.cg(b, variant=0, tol=tol, diag=diag)
    xb, itb, *_ = m.cg(b, variant=1, tol=tol, diag=diag)
    assert abs(ita - itb) <= 1 and np.linalg.norm(xa - xb) <= 1e-6 * np.linalg.norm(xa)


def test_stopping_rules():
    m = O.OracleMesh(2, (2, 2, 2), quad=O.GAUSS)
    b = m.rhs()
    # IterationNumberControl: success at the cap; SolverControl: failure at the cap
    *_, ok = m.cg(b, variant=1, control=0, tol=0.0, max_its=3)
    assert ok
    x, its, res, hist, ok = m.cg(b, variant=1, control=1, tol=0.0, max_its=3)
    assert not ok and its == 3
    # zero rhs: converged at step 0
    x, its, res, hist, ok = m.cg(np.zeros(m.n_dofs), variant=1, control=1, tol=0.0, max_its=5)
    assert ok and its == 0
