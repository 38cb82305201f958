"""-m gpu: the manufactured-solution anchor through the CUDA path (C ABI): RHS imported from the host
(bp5/step-64.cu:415-417), CG on the device, L2 error ~ h^(p+1) for p = 1..6, both quadratures, affine and
deformed meshes -- and the same error as the oracle's solve to 6 digits."""
import numpy as np
import pytest

import oracle as O
from test_manufactured_solution import check_orders, observed_orders, oracle_solve

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", range(1, 7))
@pytest.mark.parametrize("quad", [0, 1])
@pytest.mark.parametrize("deform", [0, 1])
@pytest.mark.parametrize("solver", ["merged", "standard"])
def test_cuda_l2_error_converges_with_order_p_plus_1(gpu_ctx, p, quad, deform, solver):
    import dealceed_b200 as dc
    if solver == "standard" and (quad == 0 or deform == 0) and p not in (2, 5):
        pytest.skip("standard CG: a subset of the matrix is enough")

    def cuda_solve(p, n, quad, deform, b):
        op = dc.PoissonOperator(gpu_ctx, dc.make_problem(p, (n, n, n), quadrature=quad, upper=(1., 1., 1.),
                                                         deformation=deform, eps=0.1))
        bv, xv = op.initialize_dof_vector(), op.initialize_dof_vector()
        bv.import_host(b)
        op.do_zero_out = False
        control = dc.SolverControl(5000, 1e-13 * np.linalg.norm(b))
        (dc.SolverCGFullMerge if solver == "merged" else dc.SolverCG)(control).solve(op, xv, bv, history=False)
        x = xv.to_host()
        bv.close(); xv.close(); op.close()
        return x

    errs, orders = observed_orders(p, quad, deform, cuda_solve)
    check_orders(p, deform, errs, orders)
    errs_o, _ = observed_orders(p, quad, deform, oracle_solve)
    np.testing.assert_allclose(errs, errs_o, rtol=1e-6)
