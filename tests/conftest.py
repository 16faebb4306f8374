import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    """tests/golden/<name>.npz -> {case: {key: array}} (or flat dict when keys have no '/')."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        if k.startswith("_") or "/" not in k:
            out[k] = z[k]
            continue
        case, key = k.split("/", 1)
        out.setdefault(case, {})[key] = z[k]
    return out


def bumpy_cloud(rng, n, extent=1.0):
    """Seeded surface-like cloud used by several parity tests."""
    xy = rng.uniform(-extent, extent, (n, 2))
    z = 0.15 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1]) + rng.normal(0, 0.01, n)
    pts = np.concatenate([xy, z[:, None]], 1)
    k = n // 4
    pts[:k] = rng.uniform(-extent, extent, (k, 3)) * [1, 1, 0.4]
    return pts.astype(np.float32)


def canonical_rows(inds, supports, queries, pad):
    """Sort every neighbour row by (d2 fp32 un-contracted, index): the tie-canonical form used to
    compare against nanoflann, whose std::sort leaves exact-d2 ties in unspecified order."""
    inds = np.asarray(inds).astype(np.int64)
    sp = np.concatenate([supports, np.full((1, 3), 1e18, np.float32)], 0).astype(np.float32)
    d = queries[:, None, :].astype(np.float32) - sp[np.minimum(inds, len(supports))]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
    d2 = (d2 + d[..., 2] * d[..., 2]).astype(np.float32)
    d2[inds >= pad] = np.inf
    order = np.lexsort((inds, d2), axis=1)
    return np.take_along_axis(inds, order, 1)


@pytest.fixture(scope="session")
def mvk():
    import mvkpconv_b200
    return mvkpconv_b200
