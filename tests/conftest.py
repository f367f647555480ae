import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def xgb():
    """The product package (CUDA backend)."""
    import sagemaker_xgboost_container_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def oracle():
    from oracle import gbt_oracle
    gbt_oracle.build()
    return gbt_oracle
