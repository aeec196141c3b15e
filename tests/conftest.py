import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref16():
    from oracle import oracle as o
    if not o.have_ref("int16"):
        pytest.skip("oracle/_ref/libref_int16.so not built (needs /root/reference)")
    return o.Ref("int16")


@pytest.fixture(scope="session")
def ref32():
    from oracle import oracle as o
    if not o.have_ref("fp32"):
        pytest.skip("oracle/_ref/libref_fp32.so not built (needs /root/reference)")
    return o.Ref("fp32")


@pytest.fixture(scope="session")
def accel16():
    from yolo2_b200.accel import Accelerator
    a = Accelerator(0, "int16")
    yield a
    a.close()


@pytest.fixture(scope="session")
def accel32():
    from yolo2_b200.accel import Accelerator
    a = Accelerator(0, "fp32")
    yield a
    a.close()
