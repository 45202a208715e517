import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the CPU checkers once per session."""
    import __graft_entry__ as g
    g.build_libfsb()
    g.build_oracle()
    return True


@pytest.fixture(scope="session")
def port(built):
    import oracle_lib
    return oracle_lib.OracleLib("fso")


@pytest.fixture(scope="session")
def ref(built):
    import oracle_lib
    if not oracle_lib.available("fsr"):
        pytest.skip("oracle/_ref/libfsref.so not built (needs /root/reference)")
    return oracle_lib.OracleLib("fsr")


@pytest.fixture(scope="session")
def checkers(built):
    """Every CPU checker that exists: the C restatement always, the compiled reference if present."""
    import oracle_lib
    libs = [oracle_lib.OracleLib("fso")]
    if oracle_lib.available("fsr"):
        libs.append(oracle_lib.OracleLib("fsr"))
    return libs


@pytest.fixture(scope="session")
def capi(built):
    from fluid_simulation_b200 import capi as m
    return m
