import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_vec(n):
    """x[i] = sin(i*19 + 0.4) + cos(i*i*3) -- the reference's test vector (test_sparse.c:51)."""
    i = np.arange(n, dtype=np.int64)
    return np.sin(i * 19 + 0.4) + np.cos(i * i * 3)


def rhs_matrix(ncol, R):
    """X[c][k] = sin(7c + 17k + 0.3) -- the reference's bench pattern (bench_a_mul_b.c:149-152)."""
    return np.ascontiguousarray(np.sin(7.0 * np.arange(ncol)[:, None] + 17.0 * np.arange(R)[None, :] + 0.3))


# fp64 product tolerance: |delta| <= 1e-12 * max(1, sum |terms|).
# For the inputs used here |x| <= 2 and rows hold < 1e5 entries, so the bound below is the
# same statement in array form; callers pass `scale` = an upper bound on sum |terms| per element.
def assert_close(got, want, scale=1.0, rel=1e-12, what=""):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    tol = rel * np.maximum(1.0, scale)
    err = np.abs(got - want)
    bad = np.nonzero(err > tol)[0] if np.ndim(tol) else np.nonzero(err > tol)[0]
    assert bad.size == 0, f"{what}: {bad.size} elements off, worst {err.max():.3e} at {int(err.argmax())}"


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
