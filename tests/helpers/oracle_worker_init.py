"""Picked up by the GPU worker processes of multi_gpu.run_training_with_dask in the CPU tests: selects the oracle-backed test
engine (the product itself only ever creates the CUDA backend) and binds the package as `xgboost` with the reference stubs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def use_oracle_engine():
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    import reference_stubs
    backend._BACKEND = OracleBackend(error_cls=xgb.XGBoostError)
    if reference_stubs.reference_available():
        reference_stubs.install(xgb)


def bind_package():
    """GPU runs: only the `xgboost` alias (+ the container sources when they are mounted)"""
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sagemaker_xgboost_container_b200 as xgb
    import reference_stubs
    if reference_stubs.reference_available():
        reference_stubs.install(xgb)
    else:
        xgb.install_as_xgboost()
