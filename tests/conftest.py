import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """Build the CPU oracle and the host harness once per session (test infrastructure)."""
    import oracle_lib
    oracle_lib.build()
    import hostsim_lib
    hostsim_lib.build()
    yield


def pytest_sessionfinish(session, exitstatus):
    """Parity margins observed in this run (tests/parity_util.py record()) -> JSON."""
    import json
    try:
        import parity_util as P
    except Exception:
        return
    if not P.REPORT:
        return
    path = os.environ.get("MSOC_PARITY_REPORT")
    if path is None:
        if not _has_cuda():
            return
        path = os.path.join(ROOT, "gpurun_out", "parity_report.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    bands = {"rtol": P.RTOL, "atol": P.ATOL, "max_over_fraction": P.MAX_OVER_FRACTION, "max_ratio": P.MAX_RATIO}
    with open(path, "w") as f:
        json.dump({"bands": bands, "checks": P.REPORT}, f, indent=1, sort_keys=True)
