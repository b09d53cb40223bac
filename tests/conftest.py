import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "rgba" not in d and "rgba_seed" in d:
        rgba = np.random.default_rng(int(d["rgba_seed"])).random(tuple(int(s) for s in d["rgba_shape"]),
                                                                 dtype=np.float32)
        if int(d.get("last_alpha_one", 0)):
            rgba[:, -1, 3] = 1.0
        d["rgba"] = rgba
    return d


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md section 8(c)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / (den if den > 0 else 1.0)


MPI_CASES = ["tiny_2mpi_3view", "tiny_2mpi_3view_acfalse", "alpha_one_planes", "sanity_all_alpha_one",
             "out_of_plane", "magnify_40_from_16", "minify_10_from_64", "nonsquare", "c1_small_64",
             "c2_small_4x32x64"]


@pytest.fixture(scope="session")
def golden():
    return load_golden
