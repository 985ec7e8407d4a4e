"""pytest configuration: the `gpu` marker and shared helpers."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    """Correlation-path fixtures (the convc1_* files hold the consumer layer of some of them)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                              for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith("convc1_"))


def convc1_names():
    return sorted(os.path.splitext(os.path.basename(p))[0][len("convc1_"):]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "convc1_*.npz")))


def load_convc1(name):
    """weight (64, L*(2r+1)), bias (64,), cor (T,B,64,H,W1) = relu(convc1(out[t])) of fixture `name`
    (reference core/update.py:70,77; tests/golden/make_golden.py)."""
    with np.load(os.path.join(GOLDEN_DIR, "convc1_" + name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        d = {k: z[k] for k in z.files}
    B, C, H, W1, W2, L, r, T = (int(v) for v in d["meta"])
    d["dims"] = dict(B=B, C=C, H=H, W1=W1, W2=W2, L=L, r=r, T=T)
    return d


def rel_max(a, b):
    """max|a-b| / max|b| -- the relative error north_star's 1e-3 tolerance is stated in."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def oracle():
    from oracle import corr_oracle
    corr_oracle.build()
    return corr_oracle
