import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    # build the oracle's native helpers when a compiler is present (test infrastructure only)
    if not os.path.exists(os.path.join(ROOT, "oracle", "c", "libgpdla_oracle.so")) or (
            os.path.exists("/root/reference/voigt.c")
            and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "voigt_ref.so"))):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False, capture_output=True)


@pytest.fixture(scope="session")
def synthetic_inputs():
    from gp_dla_detection_b200 import synthetic as syn
    model = syn.make_model()
    return dict(model=model, samples=syn.make_samples(10000), prior=syn.make_prior(),
                spectra=syn.make_spectra(model, 4, dla_fraction=0.5))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def boss_grid(n=1256, start=3.5563):
    return 10.0 ** (start + 1e-4 * np.arange(n))
