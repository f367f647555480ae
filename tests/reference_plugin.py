"""pytest plugin used to run the REFERENCE's own unit tests (read-only under /root/reference/test/unit) against this
package bound as `xgboost`, on CPU through the oracle-backed engine:

    PYTHONPATH=tests:.:/root/reference:/root/reference/src python -m pytest -p reference_plugin /root/reference/test/unit/test_checkpointing.py

Nothing of the reference is copied or modified; absent third-party modules are stubbed (tests/reference_stubs.py)."""
import sys
import unittest.mock


def pytest_configure(config):
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    import reference_stubs
    backend._BACKEND = OracleBackend(error_cls=xgb.XGBoostError)
    sys.modules.setdefault("mock", unittest.mock)
    reference_stubs.install(xgb)
