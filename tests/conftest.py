"""pytest configuration: marker registration and shared helpers.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol
checks.  `-m gpu` runs on the B200 box: the parity tests proper, through the C-ABI.
"""
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


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


PARITY_REPORT = []


def pytest_sessionfinish(session, exitstatus):
    """Dump every measured parity error (tests call record_parity) next to the GPU run's other artefacts."""
    if not PARITY_REPORT:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(PARITY_REPORT, f, indent=1)
    except OSError:
        pass


def record_parity(what, linf, l2, tol):
    PARITY_REPORT.append(dict(what=what, rel_linf=linf, rel_l2=l2, tol=tol))
