import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def ladder(cycle):
    """Mesh ladder of PoissonProblem::run (bp5/step-64.cu:633-663): returns (cells, upper)."""
    n_refine, rem = cycle // 6, cycle % 6
    sub = [1, 1, 1]
    if rem == 1 and cycle > 1:
        sub = [3, 2, 2]; n_refine -= 1
    if rem == 2:
        sub[0] = 2
    elif rem == 3:
        sub[0] = 3
    elif rem == 4:
        sub[0] = sub[1] = 2
    elif rem == 5:
        sub[0] = 3; sub[1] = 2
    return tuple(s * 2 ** n_refine for s in sub), tuple(float(s) for s in sub)


@pytest.fixture(scope="session")
def gpu_ctx():
    import dealceed_b200 as dc
    ctx = dc.Context(0)
    yield ctx
    ctx.close()
