import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_formulation():
    with np.load(os.path.join(GOLDEN, "formulation.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_prbs():
    with np.load(os.path.join(GOLDEN, "prbs.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cstrs_problem():
    from industrial_nnmpc_2021_b200.plants import get_cstrs_problem
    return get_cstrs_problem()


@pytest.fixture(scope="session")
def cdu_small_problem():
    """Reduced model of the CDU family (24 states, 4 inputs, 8 outputs, N=10)."""
    from industrial_nnmpc_2021_b200.plants import get_cdu_problem
    return get_cdu_problem(Nx=24, Nu=4, Ny=8, N=10)


@pytest.fixture(scope="session")
def built_lib():
    from industrial_nnmpc_2021_b200 import build
    return build.build()
