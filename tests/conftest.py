import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import parity_common
    return parity_common.load_package()


@pytest.fixture(scope="session")
def refdrv():
    import refdrv as r
    if not r.available():
        if os.path.isdir("/root/reference/Source"):
            import subprocess
            subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return r
