import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()          # compiles oracle/liboracle.so (gcc, seconds); a checker, never the product
    return oracle_py


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def native_artifacts():
    """The product is native and built in-tree (libkaarme_gpu.so + the kaarme CLI).  Normally __graft_entry__.build()
    has produced them; build them here if a fresh checkout has not (nvcc cross-compiles without a GPU)."""
    import subprocess
    pkg = os.path.join(ROOT, "canonical-k-mer-hash-table_b200")
    if not (os.path.exists(os.path.join(pkg, "libkaarme_gpu.so")) and os.path.exists(os.path.join(pkg, "kaarme"))):
        subprocess.run(["make", "-s", "-C", pkg, "all"], check=True)
