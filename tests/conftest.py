import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_addoption(parser):
    parser.addoption("--pion-lib", default=None,
                     help="run the GPU tests against another build of libpion_b200.so (e.g. the -DPION_STRICT build)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    lib = config.getoption("--pion-lib")
    if lib:
        from pion_b200.capi import load_library
        load_library(str(Path(lib).resolve()))


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The C oracle is test infrastructure: build it on demand (gcc, <2 s)."""
    import subprocess
    lib = ROOT / "oracle" / "libpion_oracle.so"
    if not lib.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"], check=True)
    yield
