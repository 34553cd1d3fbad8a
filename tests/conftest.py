"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`  : oracle vs golden vectors, host logic, C-ABI symbol export -- no GPU needed.
`-m gpu`        : the parity tests proper, driving the CUDA kernels THROUGH the C ABI on a B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # gpu-marked tests are never silently skipped into a fallback: without CUDA they are deselected
    # only through `-m "not gpu"`; if someone runs them on a CPU box they fail loudly.
    pass


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "sampler_golden.npz")
    z = np.load(path)
    cases = {}
    for k in z.files:
        name, G, dt, what = k.split("|")
        cases.setdefault((name, int(G[1:]), dt), {})[what] = z[k]
    return cases


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle
    oracle.build()
    yield
