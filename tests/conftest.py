import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Build what is missing (in-tree): the product library, the synth library, the CLI, the oracle.
    need = [
        os.path.join(ROOT, "pollen_b200", "lib", "libflatgfa.so"),
        os.path.join(ROOT, "pollen_b200", "lib", "libfgfa_synth.so"),
        os.path.join(ROOT, "bin", "fgfa"),
        os.path.join(ROOT, "oracle", "libdepth_oracle.so"),
        os.path.join(ROOT, "build", "depth_example"),
    ]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-j8", "-C", ROOT, "all"], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(d, "manifest.json"), encoding="utf-8") as f:
        m = json.load(f)
    for c in m["cases"]:
        c["dir"] = d
    return m["cases"]


@pytest.fixture(scope="session")
def fgfa_bin():
    return os.path.join(ROOT, "bin", "fgfa")
