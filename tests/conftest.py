import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the tests exercise the built artefacts; build them if this checkout has not been built yet
    # (same recipe as __graft_entry__.build(): nvcc cross-compiles sm_100a without a GPU)
    if not os.path.exists(os.path.join(ROOT, "ccs_b200", "libccsgpu.so")) or \
            not os.path.exists(os.path.join(ROOT, "ccs_b200", "bin", "ccs")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ccs_b200", "csrc"), "-j8"])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
