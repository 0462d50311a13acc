import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        import stratum_dsp_b200 as S

        return S.LIB_PATH.exists() and S.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a CPU box: they are skipped with an explicit reason.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device / extension not built")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build the oracle (gcc) and the CUDA library (nvcc cross-compiles without a GPU) once per session."""
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "-j8"], check=True)
    import stratum_dsp_b200 as S

    if not S.LIB_PATH.exists():
        S.build()
    yield
