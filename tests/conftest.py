import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)

    return load


def log_err(name: str, **values) -> None:
    """Append measured errors of a parity assertion to gpurun_out/test_errors.jsonl (read back after a GPU run to keep the asserted
    tolerances at ~3x what is measured).  Never raises."""
    import json

    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "test_errors.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: float(v) for k, v in values.items()}}) + "\n")
    except Exception:
        pass
